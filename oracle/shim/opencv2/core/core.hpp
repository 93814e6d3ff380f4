// Stand-in for the OpenCV types the reference's step / physics code touches.  Point_ arithmetic and
// cv::norm follow OpenCV's definitions (core/types.hpp) because SimpleRace's numerics depend on
// them; every drawing / image function is inert (the render path is restated in xw_oracle.c and
// checked against the real cv2 in tests/).
#pragma once
#include <cmath>
#include <ostream>
#include <string>
#include <vector>
namespace cv {
template <typename T> struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T x_, T y_) : x(x_), y(y_) {}
    template <typename U> Point_(const Point_<U>& p) : x((T)p.x), y((T)p.y) {}
};
template <typename T> Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>((T)(a.x + b.x), (T)(a.y + b.y)); }
template <typename T> Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>((T)(a.x - b.x), (T)(a.y - b.y)); }
template <typename T> Point_<T>& operator+=(Point_<T>& a, const Point_<T>& b) { a.x += b.x; a.y += b.y; return a; }
template <typename T> Point_<T> operator*(const Point_<T>& a, int b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <typename T> Point_<T> operator*(int a, const Point_<T>& b) { return Point_<T>((T)(b.x * a), (T)(b.y * a)); }
template <typename T> Point_<T> operator*(const Point_<T>& a, float b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <typename T> Point_<T> operator*(float a, const Point_<T>& b) { return Point_<T>((T)(b.x * a), (T)(b.y * a)); }
template <typename T> Point_<T> operator*(const Point_<T>& a, double b) { return Point_<T>((T)(a.x * b), (T)(a.y * b)); }
template <typename T> Point_<T> operator*(double a, const Point_<T>& b) { return Point_<T>((T)(b.x * a), (T)(b.y * a)); }
template <typename T> double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
template <typename T> std::ostream& operator<<(std::ostream& os, const Point_<T>& p) { return os << "[" << p.x << ", " << p.y << "]"; }
typedef Point_<int> Point;
typedef Point_<float> Point2f;
struct Size { int width, height; Size() : width(0), height(0) {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x, y, width, height; Rect() : x(0), y(0), width(0), height(0) {} Rect(int a, int b, int c, int d) : x(a), y(b), width(c), height(d) {} };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; } };
struct Vec3b { unsigned char val[3]; unsigned char& operator[](int i) { return val[i]; } };
struct Mat {
    int rows, cols;
    Mat() : rows(0), cols(0) {}
    Mat(int r, int c, int) : rows(r), cols(c) {}
    Mat(int r, int c, int, const Scalar&) : rows(r), cols(c) {}
    static Mat zeros(int r, int c, int t) { return Mat(r, c, t); }
    Mat operator()(const Rect& r) const { return Mat(r.height, r.width, 0); }
    void copyTo(const Mat&) const {}
    Mat clone() const { return *this; }
    Size size() const { return Size(cols, rows); }
    bool empty() const { return rows == 0 || cols == 0; }
    template <typename T> T& at(int, int) { static T t; return t; }
    template <typename T> T& at(const Point&) { static T t; return t; }
};
struct LineIterator {
    int count;
    LineIterator(const Mat&, Point, Point, int = 8) : count(0) {}
    Point pos() const { return Point(); }
    unsigned char* operator*() { static unsigned char px[4]; return px; }
    LineIterator& operator++() { return *this; }
    LineIterator operator++(int) { return *this; }
};
enum { INTER_LINEAR = 1, BORDER_CONSTANT = 0, FONT_HERSHEY_SIMPLEX = 0, EVENT_LBUTTONDBLCLK = 7 };
}  // namespace cv
#define CV_8UC3 16
#define CV_AA 16
#define CV_IMWRITE_PNG_COMPRESSION 16
