// TEST INFRASTRUCTURE ONLY.  C wrapper around the REFERENCE's own imagefilter::BilateralFilter
// (imagefilter/image_filter.cpp:184-277, compiled unmodified from /root/reference by `make -C oracle _ref`; the
// cv::Mat / copyMakeBorder / minMaxLoc stand-ins live in ref_shim/), used to pin the oracle's restatement of
// DenseSegmentation::PreprocessFeatures (oracle/vso_preprocess.cpp) against the reference itself.
#include <stdint.h>

#include <opencv2/core/core.hpp>

#include "imagefilter/image_filter.h"

extern "C" {

// dense_segmentation.cpp:164-198 with PRESMOOTH_BILATERAL: tmp = u8 * (1/255) as float (cv::Mat::convertTo, float work
// type: pinned by the cv2 golden vectors), then BilateralFilter(tmp, sigma_space, sigma_color, out).
void ref_preprocess_bilateral(const uint8_t* bgr, int width, int height, int row_stride, float sigma_space, float sigma_color, float* out) {
  cv::Mat tmp(height, width, CV_32FC3);
  const float alpha = (float)(1.0 / 255.0);
  for (int y = 0; y < height; ++y) {
    const uint8_t* s = bgr + (size_t)y * row_stride;
    float* d = tmp.ptr<float>(y);
    for (int x = 0; x < 3 * width; ++x) d[x] = (float)s[x] * alpha;
  }
  cv::Mat dst(height, width, CV_32FC3, out, (size_t)width * 3 * sizeof(float));
  imagefilter::BilateralFilter(tmp, sigma_space, sigma_color, &dst);
}

// The filter alone on a float image (1 or 3 channels).
void ref_bilateral_f32(const float* in, int width, int height, int channels, float sigma_space, float sigma_color, float* out) {
  cv::Mat src(height, width, CV_32FC(channels), (void*)in, (size_t)width * channels * sizeof(float));
  cv::Mat dst(height, width, CV_32FC(channels), out, (size_t)width * channels * sizeof(float));
  imagefilter::BilateralFilter(src, sigma_space, sigma_color, &dst);
}

}  // extern "C"
