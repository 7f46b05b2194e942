// vso_preprocess.cpp -- CPU ORACLE (test infrastructure): frame pre-processing.
// Restates DenseSegmentation::PreprocessFeatures (segmentation/dense_segmentation.cpp:164-198)
// and imagefilter::BilateralFilter (imagefilter/image_filter.cpp:130-167,184-277).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <thread>
#include <vector>

#include "vso_core.hpp"

namespace vso {

// cv::Mat::convertTo(CV_32FC3, 1.0 / 255.0) (dense_segmentation.cpp:180-181).
// OpenCV's 8U->32F convertTo uses a float work type: dst = (float)src * (float)alpha
// (+ 0.0f).  Verified against Python cv2 4.13 by tests/golden/make_golden.py.
void ConvertU8ToF32(const uint8_t* bgr, int w, int h, int row_stride, float* out) {
  const float alpha = (float)(1.0 / 255.0);
  for (int i = 0; i < h; ++i) {
    const uint8_t* src = bgr + (size_t)i * row_stride;
    float* dst = out + (size_t)i * w * 3;
    for (int j = 0; j < w * 3; ++j) dst[j] = (float)src[j] * alpha;
  }
}

// imagefilter::BilateralFilter, 3-channel case (image_filter.cpp:184-277) with the
// per-pixel kernel ParallelBilateralColor::operator() (image_filter.cpp:130-167).
// exp() is evaluated in double and rounded to float (the reference's unqualified
// exp(float) resolves to either overload depending on the libstdc++ vintage; the
// difference is <= 1 ulp of the LUT entry, far inside the 1e-5 edge-weight bar).
void BilateralFilter(const float* in, int w, int h, float sigma_space, float sigma_color,
                     float* out, int num_threads, float* lut_out, float* scale_out) {
  const int cn = 3;
  const int radius = sigma_space * 1.5f;                       // :201
  // cv::copyMakeBorder(..., BORDER_REPLICATE) (:204-207)
  const int bw = w + 2 * radius, bh = h + 2 * radius;
  std::vector<float> border((size_t)bw * bh * cn);
  for (int i = 0; i < bh; ++i) {
    const int sy = std::min(std::max(i - radius, 0), h - 1);
    for (int j = 0; j < bw; ++j) {
      const int sx = std::min(std::max(j - radius, 0), w - 1);
      std::memcpy(&border[((size_t)i * bw + j) * cn], &in[((size_t)sy * w + sx) * cn],
                  sizeof(float) * cn);
    }
  }
  const size_t step_floats = (size_t)bw * cn;

  // space offsets and weights (:210-225)
  std::vector<long> space_ofs;
  std::vector<float> space_weights;
  const float space_coeff = -0.5f / (sigma_space * sigma_space);
  for (int i = -radius; i <= radius; ++i) {
    for (int j = -radius; j <= radius; ++j) {
      const int r2 = i * i + j * j;
      if (r2 > radius * radius) continue;
      space_ofs.push_back((long)i * (long)step_floats + (long)j * cn);
      space_weights.push_back((float)std::exp((double)(space_coeff * (float)r2)));
    }
  }
  const int space_sz = (int)space_ofs.size();

  // cv::minMaxLoc over all channels (:227-230)
  double min_val = in[0], max_val = in[0];
  for (size_t k = 0, n = (size_t)w * h * cn; k < n; ++k) {
    min_val = std::min<double>(min_val, in[k]);
    max_val = std::max<double>(max_val, in[k]);
  }
  const float diff_range =
      std::max<float>(1e-3f, (max_val - min_val) * (max_val - min_val) * cn * 1.02f);  // :233-234
  const int num_bins = (1 << 12) * cn;                         // :237
  const float scale = (float)num_bins / diff_range;            // :238
  std::vector<float> lut(num_bins);
  const float color_coeff = -0.5 / (sigma_color * sigma_color);  // :241
  bool zero_reached = false;
  for (int i = 0; i < num_bins; ++i) {                         // :243-250
    if (!zero_reached) {
      lut[i] = (float)std::exp((double)((float)i / scale * color_coeff));
      zero_reached = (lut[i] < 1e-10);
    } else {
      lut[i] = 0;
    }
  }
  if (lut_out) std::memcpy(lut_out, lut.data(), sizeof(float) * num_bins);
  if (scale_out) *scale_out = scale;

  auto rows = [&](int r0, int r1) {
    for (int i = r0; i < r1; ++i) {
      const float* src_ptr = &border[((size_t)(i + radius) * bw + radius) * cn];
      float* dst_ptr = out + (size_t)i * w * cn;
      for (int j = 0; j < w; ++j, src_ptr += 3, dst_ptr += 3) {   // :134-166
        const float my_b = src_ptr[0], my_g = src_ptr[1], my_r = src_ptr[2];
        float weight_sum = 0, sum_r = 0, sum_g = 0, sum_b = 0;
        for (int k = 0; k < space_sz; ++k) {
          const float* local_ptr = src_ptr + space_ofs[k];
          const float diff_b = my_b - local_ptr[0];
          const float diff_g = my_g - local_ptr[1];
          const float diff_r = my_r - local_ptr[2];
          const int idx = (int)((diff_b * diff_b + diff_g * diff_g + diff_r * diff_r) * scale);
          const float weight = space_weights[k] * lut[idx];
          weight_sum += weight;
          sum_b += local_ptr[0] * weight;
          sum_g += local_ptr[1] * weight;
          sum_r += local_ptr[2] * weight;
        }
        if (weight_sum > 0) {
          weight_sum = 1.0 / weight_sum;
          dst_ptr[0] = sum_b * weight_sum;
          dst_ptr[1] = sum_g * weight_sum;
          dst_ptr[2] = sum_r * weight_sum;
        } else {
          dst_ptr[0] = dst_ptr[1] = dst_ptr[2] = 0.0f;
        }
      }
    }
  };
  if (num_threads <= 1) {
    rows(0, h);
  } else {
    // base::ParallelFor(BlockedRange(0, h, h / 8)) (:254-275, base/base.h:139-160):
    // row blocks are independent, any split gives identical output.
    std::vector<std::thread> th;
    const int blk = std::max(1, (h + num_threads - 1) / num_threads);
    for (int r0 = 0; r0 < h; r0 += blk) th.emplace_back(rows, r0, std::min(h, r0 + blk));
    for (auto& t : th) t.join();
  }
}

}  // namespace vso
