// Development probe (not a test): which 2-D fp32 TMA box / coordinate combinations does the B200 accept?
// usage: tma_probe <rows> <row_floats> <box_rows> <box_floats> <c0> <c1> [store]
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__global__ void load_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int box_rows, int box_floats, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) unsigned long long s_bar;
  const unsigned bar = smem_u32(&s_bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = (unsigned)(box_rows * box_floats * 4);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(smem)), "l"(&map), "r"(c0), "r"(c1), "r"(bar) : "memory");
  }
  unsigned ok = 0;
  do {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n selp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar) : "memory");
  } while (!ok);
  const float* s = reinterpret_cast<const float*>(smem);
  for (int i = threadIdx.x; i < box_rows * box_floats; i += blockDim.x) out[i] = s[i];
}
__global__ void store_kernel(const __grid_constant__ CUtensorMap map, int c0, int c1, int box_rows, int box_floats) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* s = reinterpret_cast<float*>(smem);
  for (int i = threadIdx.x; i < box_rows * box_floats; i += blockDim.x) s[i] = (float)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(&map), "r"(c0), "r"(c1), "r"(smem_u32(smem)) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  }
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const int rows = atoi(argv[1]), rowf = atoi(argv[2]), brows = atoi(argv[3]), bf = atoi(argv[4]), c0 = atoi(argv[5]), c1 = atoi(argv[6]);
  const bool store = argc > 7;
  void* f = nullptr; cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) { printf("no entry point\n"); return 1; }
  float* g; cudaMalloc(&g, (size_t)rows * rowf * 4);
  std::vector<float> h((size_t)rows * rowf); for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
  cudaMemcpy(g, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap m;
  const cuuint64_t dims[2] = {(cuuint64_t)rowf, (cuuint64_t)rows}; const cuuint64_t strides[1] = {(cuuint64_t)rowf * 4};
  const cuuint32_t box[2] = {(cuuint32_t)bf, (cuuint32_t)brows}; const cuuint32_t es[2] = {1, 1};
  CUresult r = ((EncodeTiledFn)f)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, g, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("encode rc %d ", (int)r);
  if (r != CUDA_SUCCESS) { printf("\n"); return 1; }
  const size_t smem = (size_t)brows * bf * 4;
  float* out; cudaMalloc(&out, smem);
  if (store) { cudaFuncSetAttribute(store_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); store_kernel<<<1, 128, smem>>>(m, c0, c1, brows, bf); }
  else { cudaFuncSetAttribute(load_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); load_kernel<<<1, 128, smem>>>(m, c0, c1, brows, bf, out); }
  cudaError_t e = cudaDeviceSynchronize();
  printf("%s rows %d rowf %d box %dx%d at (%d,%d): %s", store ? "store" : "load", rows, rowf, brows, bf, c0, c1, cudaGetErrorString(e));
  if (e == cudaSuccess && !store) { std::vector<float> o(smem / 4); cudaMemcpy(o.data(), out, smem, cudaMemcpyDeviceToHost); printf(" first %.0f %.0f %.0f %.0f last %.0f", o[0], o[1], o[2], o[3], o.back()); }
  if (e == cudaSuccess && store) { cudaMemcpy(h.data(), g, h.size() * 4, cudaMemcpyDeviceToHost); const int cc0 = c0 < 0 ? 0 : c0, cc1 = c1 < 0 ? 0 : c1; printf(" g[c1][c0] %.0f", h[(size_t)cc1 * rowf + cc0]); }
  printf("\n");
  return e == cudaSuccess ? 0 : 1;
}
