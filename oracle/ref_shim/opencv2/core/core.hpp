// TEST INFRASTRUCTURE ONLY.  Stand-in for <opencv2/core/core.hpp>: just enough of cv::Mat (a non-owning view of
// float / byte rows) for the reference's segmentation/pixel_distance.h to compile unmodified into oracle/_ref.
#ifndef VSO_REF_SHIM_OPENCV_CORE_HPP_
#define VSO_REF_SHIM_OPENCV_CORE_HPP_
#include <stddef.h>
#include <stdint.h>
#define CV_32F 5
#define CV_8U 0
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
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
  int type_;
};
}  // namespace cv
#endif
