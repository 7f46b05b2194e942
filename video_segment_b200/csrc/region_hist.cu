// region_hist.cu -- appearance descriptor of the region stage, sm_100a.
// Replaces, for every region of a frame at once,
//   AppearanceExtractor::AppearanceExtractor (cv::cvtColor(CV_BGR2Lab), segmentation/region_descriptor.cpp:59-89),
//   AppearanceDescriptor3D::AddFeatures (:97-111) -> ColorHistogram::AddPixelInterpolated / AddValueInterpolated
//   (segmentation/histograms.cpp:140-211), NormalizeToOne (:340-360) and ChiSquareDist (:362-407).
//
// One fused pass per frame reads 3 B (BGR) + 4 B (region id) per pixel: the 8-bit Lab value is computed with
// OpenCV's integer tables (lab_tables.inc, bit identical to cv2 over the whole colour cube) and never stored,
// the 8 trilinear weights are formed in float exactly like the reference and added to dense per-region
// accumulators in 2^-26 fixed point, which makes the result independent of the order of the adds (the
// reference's float sums depend on its raster order; tests compare against both its float and an exact
// restatement).  Neighbouring pixels mostly share region and histogram cell: lanes with the same (region,
// cell) are found with __match_any_sync and their weights summed with __reduce_add_sync, one atomic per
// group and corner instead of one per pixel and corner.
#include <stdint.h>

#include "../../include/vsb200.h"
#include "common.cuh"
#include "region_kernels.cuh"

#define LAB_TAB_QUAL __device__ const
#include "lab_tables.inc"

namespace vsb {

constexpr int kLabCbrtSize = 256 * 3 / 2 * 8;
constexpr float kHistFix = 67108864.0f;            // 2^26: a weight <= 1, 32 of them fit 32 bits

struct LabTables {
  unsigned short gamma[256];
  unsigned short cbrt[kLabCbrtSize];
};

__device__ __forceinline__ void load_lab_tables(LabTables& T) {
  for (int i = threadIdx.x; i < 256; i += blockDim.x) T.gamma[i] = kLabGammaTab[i];
  for (int i = threadIdx.x; i < kLabCbrtSize; i += blockDim.x) T.cbrt[i] = kLabCbrtTab[i];
  __syncthreads();
}

__device__ __forceinline__ int lab_descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// RGB2Lab_b, blue index 0, sRGB gamma: lab_shift 12, gamma_shift 3, lab_shift2 15
__device__ __forceinline__ uchar3 bgr_to_lab(const LabTables& T, int b8, int g8, int r8) {
  const int B = T.gamma[b8], G = T.gamma[g8], R = T.gamma[r8];
  const int fX = T.cbrt[lab_descale(R * kLabCoeff[0] + G * kLabCoeff[1] + B * kLabCoeff[2], 12)];
  const int fY = T.cbrt[lab_descale(R * kLabCoeff[3] + G * kLabCoeff[4] + B * kLabCoeff[5], 12)];
  const int fZ = T.cbrt[lab_descale(R * kLabCoeff[6] + G * kLabCoeff[7] + B * kLabCoeff[8], 12)];
  const int lscale = (116 * 255 + 50) / 100;
  const int lshift = -((16 * 255 * (1 << 15) + 50) / 100);
  const int L = lab_descale(lscale * fY + lshift, 15);
  const int a = lab_descale(500 * (fX - fY) + 128 * (1 << 15), 15);
  const int bb = lab_descale(200 * (fY - fZ) + 128 * (1 << 15), 15);
  return make_uchar3((unsigned char)min(max(L, 0), 255), (unsigned char)min(max(a, 0), 255), (unsigned char)min(max(bb, 0), 255));
}

__global__ void __launch_bounds__(256) bgr2lab_kernel(const uint8_t* __restrict__ bgr, int stride, int w, int h,
                                                      uint8_t* __restrict__ lab) {
  __shared__ LabTables T;
  load_lab_tables(T);
  const size_t n = (size_t)w * h;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
    const uint8_t* px = bgr + (size_t)y * stride + (size_t)x * 3;
    const uchar3 o = bgr_to_lab(T, __ldg(px), __ldg(px + 1), __ldg(px + 2));
    lab[i * 3] = o.x; lab[i * 3 + 1] = o.y; lab[i * 3 + 2] = o.z;
  }
}

// acc: [n_regions][total_bins] u64 fixed-point sums; cnt: [n_regions] u32 pixel counts (weight_sum_)
__global__ void __launch_bounds__(256) region_hist_kernel(const uint8_t* __restrict__ bgr, int stride, const int* __restrict__ ids,
                                                          int w, int h, int n_regions, int lum_bins, int color_bins,
                                                          unsigned long long* __restrict__ acc, unsigned int* __restrict__ cnt) {
  __shared__ LabTables T;
  load_lab_tables(T);
  const size_t n = (size_t)w * h;
  const int sq = color_bins * color_bins, total = lum_bins * sq;
  const float lum_s = (float)(lum_bins - 1), col_s = (float)(color_bins - 1);
  const unsigned lane = threadIdx.x & 31u;
  const size_t step = (size_t)gridDim.x * blockDim.x;
  const size_t n_round = (n + 31) / 32 * 32;                 // whole warps run the loop body together
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += step) {
    int r = -1;
    if (i < n) r = __ldg(&ids[i]);
    const bool valid = r >= 0 && r < n_regions;
    unsigned wfix[8];
    int bin[8];
    unsigned long long key = ~0ull - lane;                    // invalid lanes: unique keys, no partners
    if (valid) {
      const int y = (int)(i / w), x = (int)(i - (size_t)y * w);
      const uint8_t* px = bgr + (size_t)y * stride + (size_t)x * 3;
      const uchar3 lab = bgr_to_lab(T, __ldg(px), __ldg(px + 1), __ldg(px + 2));
      // AddPixelInterpolated (histograms.cpp:206-211)
      const float x_bin = (float)lab.x * (1.0f / 255.f) * lum_s;
      const float y_bin = (float)lab.y * (1.0f / 255.f) * col_s;
      const float z_bin = (float)lab.z * (1.0f / 255.f) * col_s;
      // AddValueInterpolated (histograms.cpp:140-204), weight 1.0f
      const int ix = (int)x_bin, iy = (int)y_bin, iz = (int)z_bin;
      const float dx = x_bin - (float)ix, dy = y_bin - (float)iy, dz = z_bin - (float)iz;
      const int ux = dx >= 1e-6f, uy = dy >= 1e-6f, uz = dz >= 1e-6f;
      const float xv[2] = {1.0f - dx, dx}, yv[2] = {1.0f - dy, dy}, zv[2] = {1.0f - dz, dz};
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const int a = c >> 2, b = (c >> 1) & 1, d = c & 1;
        bin[c] = (ix + (a & ux)) * sq + (iy + (b & uy)) * color_bins + (iz + (d & uz));
        const float value = xv[a] * yv[b] * zv[d] * 1.0f;
        wfix[c] = __float2uint_rn(value * kHistFix);
      }
      const unsigned cell = (unsigned)(ix * sq + iy * color_bins + iz) * 8u + (unsigned)(ux * 4 + uy * 2 + uz);
      key = ((unsigned long long)(unsigned)r << 32) | cell;
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) { wfix[c] = 0u; bin[c] = 0; }
    }
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const bool leader = valid && lane == (unsigned)(__ffs(peers) - 1);
    unsigned long long* A = acc + (size_t)(valid ? r : 0) * total;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const unsigned s = __reduce_add_sync(peers, wfix[c]);
      if (leader && s) atomicAdd(&A[bin[c]], (unsigned long long)s);
    }
    if (leader) atomicAdd(&cnt[r], (unsigned)__popc(peers));
  }
}

// NormalizeToOne (histograms.cpp:340-360): bin / weight_sum
__global__ void region_hist_finish_kernel(const unsigned long long* __restrict__ acc, const unsigned int* __restrict__ cnt,
                                          int n_regions, int total, float* __restrict__ out, float* __restrict__ weight_out) {
  const size_t n = (size_t)n_regions * total;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / total);
    const unsigned c = cnt[r];
    // one rounding to float: the fixed-point sum is exact, so the division is done in double (the reference's
    // float sum * (1.0f / weight_sum) carries the rounding noise of its own accumulation order instead)
    float v = 0.f;
    if (c) v = (float)((double)acc[i] / (67108864.0 * (double)c));
    out[i] = v;
    if (weight_out && i == (size_t)r * total) weight_out[r] = (float)c;
  }
}

// ChiSquareDist (histograms.cpp:391-407): one warp per region pair
__global__ void __launch_bounds__(256) hist_chisquare_kernel(const float* __restrict__ hist, int total, const int* __restrict__ pairs,
                                                             int n_pairs, float* __restrict__ out) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  const unsigned lane = threadIdx.x & 31u;
  if (warp >= n_pairs) return;
  const float* A = hist + (size_t)pairs[2 * warp] * total;
  const float* B = hist + (size_t)pairs[2 * warp + 1] * total;
  double sum = 0.0;
  for (int b = (int)lane; b < total; b += 32) {
    const float a = __ldg(&A[b]), c = __ldg(&B[b]);
    const float add = a + c;
    if (fabs((double)add) > 1e-12) {              // double comparison, as written in the reference
      const float sub = a - c;
      sum += (double)(sub * sub / add);
    }
  }
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if (lane == 0) out[warp] = (float)(0.5 * sum);
}

static int grid_for(size_t work_items, int block) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const size_t want = (work_items + block - 1) / block;
  const size_t cap = (size_t)sms * 8;                       // a multiple of the SM count; grid-stride loops cover the rest
  return (int)(want < cap ? (want ? want : 1) : cap);
}

int launch_region_hist(const uint8_t* dev_bgr, int row_stride_bytes, const int* dev_region_ids, int w, int h, int n_regions,
                       int lum_bins, int color_bins, unsigned long long* acc, unsigned* cnt, cudaStream_t s) {
  region_hist_kernel<<<grid_for((size_t)w * h, 256), 256, 0, s>>>(dev_bgr, row_stride_bytes, dev_region_ids, w, h, n_regions, lum_bins,
                                                                   color_bins, acc, cnt);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vsb

using namespace vsb;

static int require_device() {
  static int cached = -1;                       // device enumeration is slow (cudaGetDeviceProperties)
  if (cached < 0) cached = vsb200_device_count();
  if (cached <= 0) {
    set_error("no sm_100 CUDA device available: this path has no CPU fallback");
    return VSB200_ERR_NO_DEVICE;
  }
  return 0;
}

extern "C" {

int vsb200_bgr2lab(const uint8_t* dev_bgr, int row_stride_bytes, int width, int height, uint8_t* dev_lab_out, void* stream) {
  if (int rc = require_device()) return rc;
  if (!dev_bgr || !dev_lab_out || width < 1 || height < 1 || row_stride_bytes < width * 3) { set_error("bgr2lab: bad arguments"); return VSB200_ERR_INVALID; }
  bgr2lab_kernel<<<grid_for((size_t)width * height, 256), 256, 0, (cudaStream_t)stream>>>(dev_bgr, row_stride_bytes, width, height, dev_lab_out);
  VSB_CUDA_OK(cudaGetLastError());
  return VSB200_OK;
}

size_t vsb200_region_hist_scratch_bytes(int n_regions, int lum_bins, int color_bins) {
  if (n_regions < 1 || lum_bins < 2 || color_bins < 2) return 0;
  const size_t total = (size_t)lum_bins * color_bins * color_bins;
  return (size_t)n_regions * total * sizeof(unsigned long long) + (((size_t)n_regions * sizeof(unsigned int) + 255) & ~(size_t)255);
}

static bool hist_args_ok(int n_regions, int lum_bins, int color_bins) {
  // the (cell, flags) part of the match key is 32 bits; the 8-bit Lab range needs at least two bins per axis
  return n_regions >= 1 && lum_bins >= 2 && color_bins >= 2 && (size_t)lum_bins * color_bins * color_bins * 8 < (1ull << 32);
}

int vsb200_region_hist_reset(void* dev_scratch, int n_regions, int lum_bins, int color_bins, void* stream) {
  if (int rc = require_device()) return rc;
  if (!dev_scratch || !hist_args_ok(n_regions, lum_bins, color_bins)) { set_error("region_hist_reset: bad arguments"); return VSB200_ERR_INVALID; }
  VSB_CUDA_OK(cudaMemsetAsync(dev_scratch, 0, vsb200_region_hist_scratch_bytes(n_regions, lum_bins, color_bins), (cudaStream_t)stream));
  return VSB200_OK;
}

int vsb200_region_hist_add(const uint8_t* dev_bgr, int row_stride_bytes, const int32_t* dev_region_ids, int width, int height,
                           int n_regions, int lum_bins, int color_bins, void* dev_scratch, void* stream) {
  if (int rc = require_device()) return rc;
  if (!dev_bgr || !dev_region_ids || !dev_scratch || width < 1 || height < 1 || row_stride_bytes < width * 3 ||
      !hist_args_ok(n_regions, lum_bins, color_bins)) { set_error("region_hist_add: bad arguments"); return VSB200_ERR_INVALID; }
  const size_t total = (size_t)lum_bins * color_bins * color_bins;
  unsigned long long* acc = (unsigned long long*)dev_scratch;
  unsigned int* cnt = (unsigned int*)(acc + (size_t)n_regions * total);
  region_hist_kernel<<<grid_for((size_t)width * height, 256), 256, 0, (cudaStream_t)stream>>>(
      dev_bgr, row_stride_bytes, dev_region_ids, width, height, n_regions, lum_bins, color_bins, acc, cnt);
  VSB_CUDA_OK(cudaGetLastError());
  return VSB200_OK;
}

int vsb200_region_hist_finish(const void* dev_scratch, int n_regions, int lum_bins, int color_bins, float* dev_hist_out,
                              float* dev_weight_sum_out, void* stream) {
  if (int rc = require_device()) return rc;
  if (!dev_scratch || !dev_hist_out || !hist_args_ok(n_regions, lum_bins, color_bins)) { set_error("region_hist_finish: bad arguments"); return VSB200_ERR_INVALID; }
  const int total = lum_bins * color_bins * color_bins;
  const unsigned long long* acc = (const unsigned long long*)dev_scratch;
  const unsigned int* cnt = (const unsigned int*)(acc + (size_t)n_regions * total);
  region_hist_finish_kernel<<<grid_for((size_t)n_regions * total, 256), 256, 0, (cudaStream_t)stream>>>(acc, cnt, n_regions, total, dev_hist_out, dev_weight_sum_out);
  VSB_CUDA_OK(cudaGetLastError());
  return VSB200_OK;
}

int vsb200_hist_chisquare(const float* dev_hist, int total_bins, const int32_t* dev_pairs, int n_pairs, float* dev_out, void* stream) {
  if (int rc = require_device()) return rc;
  if (!dev_hist || !dev_pairs || !dev_out || total_bins < 1 || n_pairs < 0) { set_error("hist_chisquare: bad arguments"); return VSB200_ERR_INVALID; }
  if (n_pairs == 0) return VSB200_OK;
  hist_chisquare_kernel<<<(n_pairs * 32 + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dev_hist, total_bins, dev_pairs, n_pairs, dev_out);
  VSB_CUDA_OK(cudaGetLastError());
  return VSB200_OK;
}

}  // extern "C"
