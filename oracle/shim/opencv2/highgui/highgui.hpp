#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
inline Mat imread(const std::string&, int = 1) { return Mat(64, 64, 16); }
inline bool imwrite(const std::string&, const Mat&, const std::vector<int>& = std::vector<int>()) { return true; }
inline void imshow(const std::string&, const Mat&) {}
inline int waitKey(int = 0) { return -1; }
inline void namedWindow(const std::string&, int = 0) {}
typedef void (*MouseCallback)(int, int, int, int, void*);
inline void setMouseCallback(const std::string&, MouseCallback, void* = 0) {}
}  // namespace cv
