// vso_region.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle; never linked into the product).
// Restates the appearance descriptor of the region stage: 8-bit BGR -> Lab, the interpolated Lab histogram
// per region and the chi-square region distance.
//   AppearanceExtractor::AppearanceExtractor      segmentation/region_descriptor.cpp:59-89   (cv::cvtColor(CV_BGR2Lab), :73)
//   AppearanceDescriptor3D::AddFeatures           segmentation/region_descriptor.cpp:97-111
//   ColorHistogram::AddPixelInterpolated          segmentation/histograms.cpp:206-211
//   ColorHistogram::AddValueInterpolated          segmentation/histograms.cpp:140-204
//   ColorHistogram::NormalizeToOne                segmentation/histograms.cpp:340-360
//   ColorHistogram::ChiSquareDist / GenericDistance  segmentation/histograms.cpp:362-407
// cv::cvtColor is third-party arithmetic (OpenCV is not vendored): the integer RGB2Lab_b path is restated from its
// published algorithm with the tables of vso_lab_tables.inc; pinned against cv2 4.13 over the full colour cube
// (tests/golden/make_lab_golden.py, tests/golden/lab_bgr2lab_cv2.npz).
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "vso.h"
#include "vso_lab_tables.inc"

namespace {
inline int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }
inline uint8_t sat_u8(int v) { return (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v); }
}  // namespace

extern "C" {

// RGB2Lab_b with blue index 0 (BGR input), sRGB gamma: lab_shift 12, gamma_shift 3, lab_shift2 15.
void vso_bgr2lab(const uint8_t* bgr, int w, int h, int row_stride, uint8_t* lab_out) {
  const int lscale = (116 * 255 + 50) / 100;
  const int lshift = -((16 * 255 * (1 << 15) + 50) / 100);
  for (int y = 0; y < h; ++y) {
    const uint8_t* src = bgr + (size_t)y * row_stride;
    uint8_t* dst = lab_out + (size_t)y * w * 3;
    for (int x = 0; x < w; ++x, src += 3, dst += 3) {
      const int B = kLabGammaTab[src[0]], G = kLabGammaTab[src[1]], R = kLabGammaTab[src[2]];
      const int fX = kLabCbrtTab[descale(R * kLabCoeff[0] + G * kLabCoeff[1] + B * kLabCoeff[2], 12)];
      const int fY = kLabCbrtTab[descale(R * kLabCoeff[3] + G * kLabCoeff[4] + B * kLabCoeff[5], 12)];
      const int fZ = kLabCbrtTab[descale(R * kLabCoeff[6] + G * kLabCoeff[7] + B * kLabCoeff[8], 12)];
      dst[0] = sat_u8(descale(lscale * fY + lshift, 15));
      dst[1] = sat_u8(descale(500 * (fX - fY) + 128 * (1 << 15), 15));
      dst[2] = sat_u8(descale(200 * (fY - fZ) + 128 * (1 << 15), 15));
    }
  }
}

// AddFeatures over one frame: every pixel with a region id in [0, n_regions) adds its Lab value to its region's
// histogram, in raster order (the order of the reference's scan intervals), accumulating in float like the reference
// (exact == 0) or in double (exact != 0: the order-independent value the float sums scatter around).
// hist: [n_regions][lum_bins * color_bins * color_bins] accumulators (double storage; float mode rounds every add to
// float), weight_sum: [n_regions].  Call once per frame of the chunk set; then vso_hist_normalize.
void vso_region_hist_add(const uint8_t* lab, const int32_t* ids, int w, int h, int n_regions, int lum_bins,
                         int color_bins, int exact, double* hist, double* weight_sum) {
  const int sq = color_bins * color_bins, total = lum_bins * sq;
  const size_t n = (size_t)w * h;
  for (size_t i = 0; i < n; ++i) {
    const int r = ids[i];
    if (r < 0 || r >= n_regions) continue;
    const uint8_t* px = lab + i * 3;
    // AddPixelInterpolated (histograms.cpp:206-211), weight 1.0f
    const float x_bin = (float)px[0] * (1.0f / 255.f) * (lum_bins - 1);
    const float y_bin = (float)px[1] * (1.0f / 255.f) * (color_bins - 1);
    const float z_bin = (float)px[2] * (1.0f / 255.f) * (color_bins - 1);
    // AddValueInterpolated (histograms.cpp:140-204)
    const int int_x = x_bin, int_y = y_bin, int_z = z_bin;
    const float dx = x_bin - (float)int_x, dy = y_bin - (float)int_y, dz = z_bin - (float)int_z;
    const int xb[2] = {int_x, int_x + (dx >= 1e-6f)}, yb[2] = {int_y, int_y + (dy >= 1e-6f)}, zb[2] = {int_z, int_z + (dz >= 1e-6f)};
    const float xv[2] = {1.0f - dx, dx}, yv[2] = {1.0f - dy, dy}, zv[2] = {1.0f - dz, dz};
    double* H = hist + (size_t)r * total;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        for (int c = 0; c < 2; ++c) {
          const int bin = xb[a] * sq + yb[b] * color_bins + zb[c];
          const float value = xv[a] * yv[b] * zv[c] * 1.0f;
          if (exact) H[bin] += (double)value;
          else H[bin] = (double)((float)H[bin] + value);
        }
    weight_sum[r] += 1.0;
  }
}

// NormalizeToOne (histograms.cpp:340-360): bins *= 1.0f / weight_sum (float), regions without pixels stay zero.
void vso_hist_normalize(const double* hist, const double* weight_sum, int n_regions, int total_bins, int exact, float* out) {
  for (int r = 0; r < n_regions; ++r) {
    const double* H = hist + (size_t)r * total_bins;
    float* O = out + (size_t)r * total_bins;
    if (weight_sum[r] == 0) { memset(O, 0, sizeof(float) * total_bins); continue; }
    if (exact) {
      for (int b = 0; b < total_bins; ++b) O[b] = (float)(H[b] / weight_sum[r]);
    } else {
      const float denom = 1.0f / weight_sum[r];      // as written in the reference: double division, rounded to float
      for (int b = 0; b < total_bins; ++b) O[b] = (float)H[b] * denom;
    }
  }
}

// AppearanceDescriptor3D::RegionDistance = ChiSquareDist (histograms.cpp:391-407): 0.5 * sum over bins of
// (a - b)^2 / (a + b) (float terms, double sum; bins where |a + b| <= 1e-12 contribute nothing).
void vso_hist_chisquare(const float* hist, int total_bins, const int32_t* pairs, int n_pairs, float* out) {
  for (int p = 0; p < n_pairs; ++p) {
    const float* A = hist + (size_t)pairs[2 * p] * total_bins;
    const float* B = hist + (size_t)pairs[2 * p + 1] * total_bins;
    double sum = 0;
    for (int b = 0; b < total_bins; ++b) {
      const float add = A[b] + B[b];
      if (fabs(add) > 1e-12) {
        const float sub = A[b] - B[b];
        sum += sub * sub / add;
      }
    }
    out[p] = (float)(0.5 * sum);
  }
}

}  // extern "C"
