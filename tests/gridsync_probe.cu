// Development probe: cost of a cooperative grid barrier on this GPU (cg::grid_group::sync vs an atomic counter barrier).
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdio.h>
namespace cg = cooperative_groups;
__global__ void __launch_bounds__(512) k_cg(int iters, unsigned long long* out) {
  cg::grid_group g = cg::this_grid();
  unsigned long long t0 = 0, t1 = 0;
  g.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < iters; ++i) g.sync();
  if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); out[0] = t1 - t0; }
}
__device__ __forceinline__ void atomic_barrier(unsigned* counter, unsigned& phase, unsigned nblk) {
  __syncthreads();
  if (threadIdx.x == 0) {
    phase += nblk;
    __threadfence();
    atomicAdd(counter, 1u);
    while (*((volatile unsigned*)counter) < phase) { }
    __threadfence();
  }
  __syncthreads();
}
__global__ void __launch_bounds__(512) k_atomic(int iters, unsigned* counter, unsigned long long* out) {
  unsigned phase = 0;
  unsigned long long t0 = 0, t1 = 0;
  atomic_barrier(counter, phase, gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < iters; ++i) atomic_barrier(counter, phase, gridDim.x);
  if (blockIdx.x == 0 && threadIdx.x == 0) { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1)); out[0] = t1 - t0; }
}
int main() {
  int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* out; cudaMallocManaged(&out, 16); unsigned* counter; cudaMalloc(&counter, 4);
  int iters = 2000;
  for (int rep = 0; rep < 2; ++rep) {
    void* args[] = {&iters, &out};
    cudaLaunchCooperativeKernel((void*)k_cg, dim3(sms), dim3(512), args, 0, 0);
    cudaError_t e = cudaDeviceSynchronize();
    printf("cg grid.sync: %s %.3f us per barrier (%d blocks)\n", cudaGetErrorString(e), out[0] / 1000.0 / iters, sms);
    cudaMemset(counter, 0, 4);
    void* args2[] = {&iters, &counter, &out};
    cudaLaunchCooperativeKernel((void*)k_atomic, dim3(sms), dim3(512), args2, 0, 0);
    e = cudaDeviceSynchronize();
    printf("atomic barrier: %s %.3f us per barrier\n", cudaGetErrorString(e), out[0] / 1000.0 / iters);
  }
  return 0;
}
