// TEST INFRASTRUCTURE ONLY.  Stand-in for <opencv2/imgproc/imgproc.hpp>: cv::copyMakeBorder with BORDER_REPLICATE,
// the one imgproc call of the reference's imagefilter/image_filter.cpp (:206-207).  Semantics checked against cv2 4.13
// golden vectors (tests/golden/cv2_thirdparty.npz): every border texel copies the nearest source texel.
#ifndef VSO_REF_SHIM_OPENCV_IMGPROC_HPP_
#define VSO_REF_SHIM_OPENCV_IMGPROC_HPP_
#include <cstring>
#include "opencv2/core/core.hpp"
enum { CV_BGR2Lab = 44 };
namespace cv {
enum { BORDER_REPLICATE = 1 };
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int border_type) {
  (void)border_type;
  (void)bottom;
  (void)right;
  const size_t es = src.elemSize();
  for (int y = 0; y < dst.rows; ++y) {
    int sy = y - top;
    sy = sy < 0 ? 0 : (sy >= src.rows ? src.rows - 1 : sy);
    for (int x = 0; x < dst.cols; ++x) {
      int sx = x - left;
      sx = sx < 0 ? 0 : (sx >= src.cols ? src.cols - 1 : sx);
      std::memcpy(dst.ptr<uchar>(y) + (size_t)x * es, src.ptr<uchar>(sy) + (size_t)sx * es, es);
    }
  }
}
// Region-stage feature extraction (region_descriptor.cpp:59-89) is linked for its vtables only in oracle/_ref; the
// descriptor oracle is pinned through histograms.cpp and cv2 golden vectors instead.  Abort if reached.
// Functional only where the including library provides vso_shim_bgr2lab (oracle/ref_hier_wrap.cpp forwards to the
// oracle's 8-bit BGR->Lab, bit identical to cv2 4.13 over the whole colour cube); aborts elsewhere.
extern "C" void vso_shim_bgr2lab(const unsigned char* bgr, int w, int h, int row_stride, unsigned char* lab_out) __attribute__((weak));
inline void cvtColor(const Mat& src, Mat& dst, int code, int = 0) {
  if (code != CV_BGR2Lab || !vso_shim_bgr2lab || src.type() != CV_8UC3) std::abort();
  if (dst.rows != src.rows || dst.cols != src.cols || dst.type() != CV_8UC3 || !dst.data) dst.create(src.rows, src.cols, CV_8UC3);
  for (int y = 0; y < src.rows; ++y) vso_shim_bgr2lab(src.ptr<uchar>(y), src.cols, 1, (int)src.step[0], dst.ptr<uchar>(y));
}
inline Scalar mean(const Mat&) { std::abort(); }
// Not on the default path (PRESMOOTH_GAUSSIAN, compute_vectorization): third-party algorithms, abort if reached.
inline void GaussianBlur(const Mat&, Mat&, Size, double, double = 0, int = 4) { std::abort(); }
template <class A, class B> inline void approxPolyDP(const A&, B&, double, bool) { std::abort(); }
}  // namespace cv
#endif
