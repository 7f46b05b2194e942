// TEST INFRASTRUCTURE ONLY.  Stand-in for <opencv2/core/core.hpp>: just enough of cv::Mat (a view of float / byte
// rows, optionally owning its buffer, ROI and reshape headers), cv::minMaxLoc and the point types for the reference's
// segmentation/pixel_distance.h and imagefilter/image_filter.cpp to compile unmodified into oracle/_ref.
#ifndef VSO_REF_SHIM_OPENCV_CORE_HPP_
#define VSO_REF_SHIM_OPENCV_CORE_HPP_
#include <stddef.h>
#include <stdint.h>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <type_traits>
#define CV_32F 5
#define CV_8U 0
#define CV_32S 4
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_32FC(n) CV_MAKETYPE(CV_32F, (n))
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_8UC(n) CV_MAKETYPE(CV_8U, (n))
typedef unsigned char uchar;
namespace cv {
struct Size {
  int width, height;
  Size(int w = 0, int h = 0) : width(w), height(h) {}
  bool operator==(const Size& o) const { return width == o.width && height == o.height; }
  bool operator!=(const Size& o) const { return !(*this == o); }
};
template <class S> S& operator<<(S& s, const Size& z) { return s << z.width << "x" << z.height; }
template <class T> struct Point_ {
  T x, y;
  Point_(T x_ = 0, T y_ = 0) : x(x_), y(y_) {}
  // cross-type conversion: to int it rounds to nearest even (cv::saturate_cast<int>(float) = cvRound), only used by
  // the reference's debug drawing
  template <class U> Point_(const Point_<U>& o) : x(conv(o.x)), y(conv(o.y)) {}
  template <class U> static T conv(U v) { return std::is_integral<T>::value && !std::is_integral<U>::value ? (T)std::nearbyint(v) : (T)v; }
  Point_ operator*(T f) const { return Point_(x * f, y * f); }
  Point_ operator+(const Point_& o) const { return Point_(x + o.x, y + o.y); }
  Point_ operator-(const Point_& o) const { return Point_(x - o.x, y - o.y); }
  Point_ operator-() const { return Point_(-x, -y); }
  Point_& operator+=(const Point_& o) { x += o.x; y += o.y; return *this; }
  Point_& operator-=(const Point_& o) { x -= o.x; y -= o.y; return *this; }
  Point_& operator*=(T f) { x *= f; y *= f; return *this; }
  bool operator==(const Point_& o) const { return x == o.x && y == o.y; }
  bool operator!=(const Point_& o) const { return !(*this == o); }
  T dot(const Point_& o) const { return x * o.x + y * o.y; }
  double cross(const Point_& o) const { return (double)x * o.y - (double)y * o.x; }
};
typedef Point_<int> Point;
typedef Point_<int> Point2i;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  double operator[](int i) const { return val[i]; }
  double& operator[](int i) { return val[i]; }
};
template <class T> struct Point3_ {
  T x, y, z;
  Point3_(T x_ = 0, T y_ = 0, T z_ = 0) : x(x_), y(y_), z(z_) {}
  bool operator==(const Point3_& o) const { return x == o.x && y == o.y && z == o.z; }
  bool operator!=(const Point3_& o) const { return !(*this == o); }
};
typedef Point3_<float> Point3f;
inline double norm(const Point2f& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
struct Rect {
  int x, y, width, height;
  Rect(int x_ = 0, int y_ = 0, int w = 0, int h = 0) : x(x_), y(y_), width(w), height(h) {}
};
struct MatStep {
  size_t p[2];
  MatStep() { p[0] = p[1] = 0; }
  size_t operator[](int i) const { return p[i]; }
  operator size_t() const { return p[0]; }
};
class Mat {
 public:
  Mat() : rows(0), cols(0), data(nullptr), type_(0) {}
  Mat(int r, int c, int type, void* d, size_t row_step) : rows(r), cols(c), data((uchar*)d), type_(type) {
    step.p[0] = row_step;
    step.p[1] = elemSize();
  }
  // owning, continuous rows (cv::Mat(rows, cols, type) allocates step = cols * elemSize())
  Mat(int r, int c, int type) : rows(r), cols(c), data(nullptr), type_(type) {
    step.p[1] = elemSize();
    step.p[0] = (size_t)c * elemSize();
    own_.reset(new uchar[(size_t)r * step.p[0] + 16], std::default_delete<uchar[]>());
    data = own_.get();
  }
  // deep copy into a matrix of the same geometry (allocated if it is not)
  void copyTo(Mat& dst) const {
    if (dst.rows != rows || dst.cols != cols || dst.type() != type_ || !dst.data) dst.create(rows, cols, type_);
    for (int y = 0; y < rows; ++y) std::memcpy(dst.ptr<uchar>(y), ptr<uchar>(y), (size_t)cols * elemSize());
  }
  // convertTo for the one conversion on the path, 8U -> 32F with a scale (dense_segmentation.cpp:181): OpenCV's 8U->32F
  // kernel works in float, dst = (float)src * (float)alpha (+ (float)beta); pinned by cv2 golden vectors
  // (tests/golden/cv2_thirdparty.npz).
  void convertTo(Mat& dst, int rtype, double alpha = 1, double beta = 0) const {
    if (depth() != CV_8U || (rtype & 7) != CV_32F) std::abort();
    const int t = CV_MAKETYPE(CV_32F, channels());
    if (dst.rows != rows || dst.cols != cols || dst.type() != t || !dst.data) dst.create(rows, cols, t);
    const float a = (float)alpha, b = (float)beta;
    for (int y = 0; y < rows; ++y) {
      const uchar* s = ptr<uchar>(y);
      float* d = dst.ptr<float>(y);
      for (int x = 0; x < cols * channels(); ++x) d[x] = (float)s[x] * a + b;
    }
  }
  void create(int r, int c, int type) { *this = Mat(r, c, type); }
  Mat row(int y) const { return Mat(*this, Rect(0, y, cols, 1)); }
  Mat col(int x) const { return Mat(*this, Rect(x, 0, 1, rows)); }
  // setTo for the element types in use (int / float / byte matrices, every channel the same value)
  Mat& setTo(double v) {
    for (int y = 0; y < rows; ++y)
      for (int x = 0; x < cols * channels(); ++x) {
        if (depth() == CV_32F) ptr<float>(y)[x] = (float)v;
        else if (depth() == CV_32S) ptr<int>(y)[x] = (int)v;
        else ptr<uchar>(y)[x] = (uchar)v;
      }
    return *this;
  }
  // region-of-interest header sharing the parent's rows
  Mat(const Mat& m, const Rect& roi) : rows(roi.height), cols(roi.width), data(m.data + (size_t)roi.y * m.step.p[0] + (size_t)roi.x * m.elemSize()), step(m.step), own_(m.own_), type_(m.type_) {}
  // reshape(cn) of a continuous matrix: same rows, cols * channels() / cn elements per row
  Mat reshape(int cn) const {
    Mat r(*this);
    r.type_ = CV_MAKETYPE(depth(), cn);
    r.cols = cols * channels() / cn;
    r.step.p[1] = r.elemSize();
    return r;
  }
  template <class T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step.p[0]); }
  template <class T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step.p[0]); }
  template <class T> T& at(int y, int x) { return ((T*)(data + (size_t)y * step.p[0]))[x]; }
  template <class T> const T& at(int y, int x) const { return ((const T*)(data + (size_t)y * step.p[0]))[x]; }
  int type() const { return type_; }
  int depth() const { return type_ & 7; }
  int channels() const { return (type_ >> 3) + 1; }
  size_t elemSize1() const { return depth() == CV_32F || depth() == CV_32S ? 4 : 1; }
  size_t elemSize() const { return elemSize1() * channels(); }
  size_t step1(int i = 0) const { return step.p[i] / elemSize1(); }
  bool empty() const { return data == nullptr; }
  Size size() const { return Size(cols, rows); }
  int rows, cols;
  uchar* data;
  MatStep step;
 private:
  std::shared_ptr<uchar> own_;
  int type_;
};
// the reference's debug drawing (segmentation_util.cpp:414-481) is not on the path: no-ops
enum { FONT_HERSHEY_PLAIN = 1 };
inline void line(Mat&, Point, Point, const Scalar&, int = 1, int = 8, int = 0) {}
template <class S> inline void putText(Mat&, const S&, Point, int, double, Scalar, int = 1, int = 8, bool = false) {}
inline void ellipse(Mat&, Point, Size, double, double, double, const Scalar&, int = 1, int = 8, int = 0) {}
// cv::minMaxLoc on a single-channel CV_32F matrix (values only): the extrema of the elements, as doubles.
// Semantics checked against cv2 4.13 golden vectors (tests/golden/cv2_thirdparty.npz).
inline void minMaxLoc(const Mat& m, double* min_val, double* max_val) {
  float lo = m.ptr<float>(0)[0], hi = lo;
  for (int y = 0; y < m.rows; ++y) {
    const float* p = m.ptr<float>(y);
    for (int x = 0; x < m.cols; ++x) {
      if (p[x] < lo) lo = p[x];
      if (p[x] > hi) hi = p[x];
    }
  }
  if (min_val) *min_val = lo;
  if (max_val) *max_val = hi;
}
}  // namespace cv
#endif
