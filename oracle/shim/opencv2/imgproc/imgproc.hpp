#pragma once
#include <opencv2/core/core.hpp>
namespace cv {
inline void resize(const Mat&, Mat&, Size, double = 0, double = 0, int = 1) {}
inline Mat getRotationMatrix2D(Point2f, double, double) { return Mat(2, 3, 0); }
inline void warpAffine(const Mat&, Mat&, const Mat&, Size, int = 1, int = 0, const Scalar& = Scalar()) {}
inline void copyMakeBorder(const Mat&, Mat&, int, int, int, int, int, const Scalar& = Scalar()) {}
inline void addWeighted(const Mat&, double, const Mat&, double, double, Mat&) {}
template <typename P> void circle(Mat, P, int, const Scalar&, int = 1, int = 8, int = 0) {}
template <typename P, typename Q> void line(Mat, P, Q, const Scalar&, int = 1, int = 8, int = 0) {}
template <typename P, typename Q> void rectangle(Mat, P, Q, const Scalar&, int = 1, int = 8, int = 0) {}
template <typename P, typename Q> void arrowedLine(Mat, P, Q, const Scalar&, int = 1, int = 8, int = 0, double = 0.1) {}
template <typename P> void putText(Mat, const std::string&, P, int, double, const Scalar&, int = 1, int = 8, bool = false) {}
}  // namespace cv
