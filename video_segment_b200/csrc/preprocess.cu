// preprocess.cu -- u8 BGR -> float/255 -> joint bilateral filter, sm_100a.
// Replaces DenseSegmentation::PreprocessFeatures (segmentation/dense_segmentation.cpp:164-198)
// and imagefilter::BilateralFilter (imagefilter/image_filter.cpp:130-167,184-277).
//
// Data flow per frame (the float image between convertTo and the filter never
// touches HBM):  u8 frame --minmax_u8--> 2 words --build_lut--> 12288-entry exp LUT
//                u8 frame --bilateral_u8 (tile + 4 px halo in shared memory)--> f32 HWC.
// HBM bytes/frame: 3N (minmax) + ~3N (tile loads) + 12N (store) = ~18N, compute bound
// (49 taps x ~22 flops) -- see DESIGN.md.
#include <math.h>
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace vsb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// scratch layout (bytes): [0] uint max word, [4] uint (255 - min) word, [8] float scale,
// [64 ...] float lut[kLutBins]
size_t preprocess_scratch_bytes() { return 64 + sizeof(float) * kLutBins; }

__constant__ float c_space_w[kBilateralTaps];
static bool g_taps_ready[64] = {false};

// cv::minMaxLoc over all channels (image_filter.cpp:227-230).  (float)u8 * alpha is
// monotone in u8, so the extrema are taken on the bytes.
__global__ void minmax_u8_kernel(const uint8_t* __restrict__ bgr, int stride, int row_bytes, int h,
                                 unsigned int* __restrict__ words) {
  unsigned int mx = 0, mn = 255;
  const long long total4 = (long long)h * ((row_bytes + 3) / 4);
  const int per_row = (row_bytes + 3) / 4;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total4;
       t += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(t / per_row), c = (int)(t % per_row) * 4;
    const uint8_t* p = bgr + (size_t)y * stride + c;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (c + k < row_bytes) {
        const unsigned int v = p[k];
        mx = max(mx, v);
        mn = min(mn, v);
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&words[0], mx);
    atomicMax(&words[1], 255u - mn);
  }
}

// LUT of image_filter.cpp:232-250.  exp is evaluated in double and rounded to float
// (same definition as the oracle).  LUT entries after the first value < 1e-10 are 0.
__global__ void build_lut_kernel(const unsigned int* __restrict__ words, float* __restrict__ scale_out,
                                 float* __restrict__ lut) {
  const float alpha = (float)(1.0 / 255.0);
  const float max_f = (float)words[0] * alpha;
  const float min_f = (float)(255u - words[1]) * alpha;
  const double max_val = (double)max_f, min_val = (double)min_f;
  const float cand = (float)((max_val - min_val) * (max_val - min_val) * 3 * (double)1.02f);
  const float diff_range = fmaxf(1e-3f, cand);
  const float scale = (float)kLutBins / diff_range;
  const float color_coeff = -8.0f;   // -0.5 / (0.25f * 0.25f), image_filter.cpp:241
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *scale_out = scale;
  if (i >= kLutBins) return;
  const float v = (float)exp((double)((float)i / scale * color_coeff));
  bool zero = false;
  if (i > 0) {
    const float vp = (float)exp((double)((float)(i - 1) / scale * color_coeff));
    zero = ((double)vp < 1e-10);
  }
  lut[i] = zero ? 0.0f : v;
}

constexpr int kTW = 64, kTH = 16;          // output tile
constexpr int kSW = kTW + 2 * kBilateralRadius, kSH = kTH + 2 * kBilateralRadius;

// ParallelBilateralColor::operator() (image_filter.cpp:130-167) on a shared-memory tile with
// replicate border (cv::copyMakeBorder BORDER_REPLICATE, image_filter.cpp:204-207).
__global__ void __launch_bounds__(256) bilateral_u8_kernel(const uint8_t* __restrict__ bgr, int stride, int w, int h,
                                                           const float* __restrict__ scale_ptr,
                                                           const float* __restrict__ lut,
                                                           float* __restrict__ out) {
  __shared__ float tile[kSH][kSW * 3];
  const int x0 = blockIdx.x * kTW, y0 = blockIdx.y * kTH;
  const float alpha = (float)(1.0 / 255.0);
  const int tid = threadIdx.y * blockDim.x + threadIdx.x;
  for (int t = tid; t < kSH * kSW; t += 256) {
    const int ty = t / kSW, tx = t % kSW;
    const int sy = min(max(y0 + ty - kBilateralRadius, 0), h - 1);
    const int sx = min(max(x0 + tx - kBilateralRadius, 0), w - 1);
    const uint8_t* p = bgr + (size_t)sy * stride + sx * 3;
    tile[ty][tx * 3 + 0] = (float)p[0] * alpha;   // convertTo(CV_32FC3, 1/255), dense_segmentation.cpp:181
    tile[ty][tx * 3 + 1] = (float)p[1] * alpha;
    tile[ty][tx * 3 + 2] = (float)p[2] * alpha;
  }
  __syncthreads();
  const float scale = *scale_ptr;
  const int lx = threadIdx.x;
#pragma unroll 1
  for (int r = 0; r < 4; ++r) {
    const int ly = threadIdx.y + 4 * r;
    const int x = x0 + lx, y = y0 + ly;
    if (x >= w || y >= h) continue;
    const float* c = &tile[ly + kBilateralRadius][(lx + kBilateralRadius) * 3];
    const float my_b = c[0], my_g = c[1], my_r = c[2];
    float weight_sum = 0, sum_r = 0, sum_g = 0, sum_b = 0;
    int k = 0;   // compile-time after unrolling: taps in the reference's (i, j) order
#pragma unroll
    for (int i = -kBilateralRadius; i <= kBilateralRadius; ++i) {
#pragma unroll
      for (int j = -kBilateralRadius; j <= kBilateralRadius; ++j) {
        if (i * i + j * j > kBilateralRadius * kBilateralRadius) continue;
        const float* l = &tile[ly + kBilateralRadius + i][(lx + kBilateralRadius + j) * 3];
        const float lb = l[0], lg = l[1], lr = l[2];
        const float diff_b = my_b - lb, diff_g = my_g - lg, diff_r = my_r - lr;
        const int idx = (int)((diff_b * diff_b + diff_g * diff_g + diff_r * diff_r) * scale);
        const float weight = c_space_w[k] * __ldg(&lut[idx]);
        weight_sum += weight;
        sum_b += lb * weight;
        sum_g += lg * weight;
        sum_r += lr * weight;
        ++k;
      }
    }
    float* o = out + ((size_t)y * w + x) * 3;
    if (weight_sum > 0) {
      const float inv = (float)(1.0 / (double)weight_sum);
      o[0] = sum_b * inv;
      o[1] = sum_g * inv;
      o[2] = sum_r * inv;
    } else {
      o[0] = o[1] = o[2] = 0.0f;
    }
  }
}

// PRESMOOTH_NONE: plain convertTo.
__global__ void convert_u8_kernel(const uint8_t* __restrict__ bgr, int stride, int row_elems, int h,
                                  float* __restrict__ out) {
  const float alpha = (float)(1.0 / 255.0);
  const long long total = (long long)row_elems * h;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(t / row_elems), c = (int)(t % row_elems);
    out[t] = (float)bgr[(size_t)y * stride + c] * alpha;
  }
}

static int ensure_taps() {
  int dev = 0;
  VSB_CUDA_OK(cudaGetDevice(&dev));
  if (dev < 64 && g_taps_ready[dev]) return 0;
  float sw[kBilateralTaps];
  const float sigma_space = 3.0f;
  const float space_coeff = -0.5f / (sigma_space * sigma_space);   // image_filter.cpp:213
  int n = 0;
  for (int i = -kBilateralRadius; i <= kBilateralRadius; ++i)
    for (int j = -kBilateralRadius; j <= kBilateralRadius; ++j) {
      const int r2 = i * i + j * j;
      if (r2 > kBilateralRadius * kBilateralRadius) continue;
      sw[n++] = (float)exp((double)(space_coeff * (float)r2));    // image_filter.cpp:222
    }
  if (n != kBilateralTaps) { set_error("tap count %d", n); return 1; }
  VSB_CUDA_OK(cudaMemcpyToSymbol(c_space_w, sw, sizeof(sw)));
  if (dev < 64) g_taps_ready[dev] = true;
  return 0;
}

int launch_preprocess(const uint8_t* bgr, int stride, int w, int h, int presmoothing, float* out,
                      void* scratch, cudaStream_t s) {
  if (presmoothing == 0) {
    convert_u8_kernel<<<148 * 8, 256, 0, s>>>(bgr, stride, w * 3, h, out);
    VSB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (presmoothing != 2) { set_error("presmoothing %d unsupported (gaussian = cv::GaussianBlur, third-party)", presmoothing); return 5; }
  if (int rc = ensure_taps()) return rc;
  unsigned int* words = (unsigned int*)scratch;
  float* scale = (float*)((char*)scratch + 8);
  float* lut = (float*)((char*)scratch + 64);
  VSB_CUDA_OK(cudaMemsetAsync(words, 0, 8, s));
  minmax_u8_kernel<<<148 * 4, 256, 0, s>>>(bgr, stride, w * 3, h, words);
  build_lut_kernel<<<(kLutBins + 255) / 256, 256, 0, s>>>(words, scale, lut);
  dim3 grid((w + kTW - 1) / kTW, (h + kTH - 1) / kTH), block(kTW, 4);
  bilateral_u8_kernel<<<grid, block, 0, s>>>(bgr, stride, w, h, scale, lut, out);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vsb
