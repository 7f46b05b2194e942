// TEST INFRASTRUCTURE ONLY.  Stand-in for <opencv2/core/core.hpp>: just enough of cv::Mat (a view of float / byte
// rows, optionally owning its buffer, ROI and reshape headers), cv::minMaxLoc and the point types for the reference's
// segmentation/pixel_distance.h and imagefilter/image_filter.cpp to compile unmodified into oracle/_ref.
#ifndef VSO_REF_SHIM_OPENCV_CORE_HPP_
#define VSO_REF_SHIM_OPENCV_CORE_HPP_
#include <stddef.h>
#include <stdint.h>
#include <cmath>
#include <memory>
#define CV_32F 5
#define CV_8U 0
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_32FC(n) CV_MAKETYPE(CV_32F, (n))
#define CV_32FC3 CV_MAKETYPE(CV_32F, 3)
#define CV_32FC2 CV_MAKETYPE(CV_32F, 2)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
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
  Point_ operator*(T f) const { return Point_(x * f, y * f); }
};
typedef Point_<int> Point;
typedef Point_<float> Point2f;
template <class T> struct Point3_ {
  T x, y, z;
  Point3_(T x_ = 0, T y_ = 0, T z_ = 0) : x(x_), y(y_), z(z_) {}
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
  size_t elemSize1() const { return depth() == CV_32F ? 4 : 1; }
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
