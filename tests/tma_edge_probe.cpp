// Development probe (not a test): calls vsb200_edge_build (TMA path) through the C ABI on random frames
// and compares it with the staged-load kernel (VSB200_NO_TMA=1 in a second process).  usage: tma_edge_probe W H out.bin
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../include/vsb200.h"
int main(int argc, char** argv) {
  const int w = atoi(argv[1]), h = atoi(argv[2]);
  std::vector<float> a((size_t)w * h * 3), b(a.size());
  srand(1);
  for (auto& v : a) v = (rand() % 1000) / 1000.0f;
  for (auto& v : b) v = (rand() % 1000) / 1000.0f;
  float *da, *db, *sp, *tp;
  cudaMalloc(&da, a.size() * 4); cudaMalloc(&db, b.size() * 4);
  cudaMalloc(&sp, (size_t)w * h * 16); cudaMalloc(&tp, (size_t)w * h * 36);
  cudaMemcpy(da, a.data(), a.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b.data(), b.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(sp, 0, (size_t)w * h * 16); cudaMemset(tp, 0, (size_t)w * h * 36);
  int rc = vsb200_edge_build(da, db, nullptr, w, h, 0, sp, tp, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  printf("rc %d (%s) sync: %s\n", rc, vsb200_last_error(), cudaGetErrorString(e));
  if (rc || e != cudaSuccess) return 1;
  std::vector<float> hs((size_t)w * h * 4), ht((size_t)w * h * 9);
  cudaMemcpy(hs.data(), sp, hs.size() * 4, cudaMemcpyDeviceToHost);
  cudaMemcpy(ht.data(), tp, ht.size() * 4, cudaMemcpyDeviceToHost);
  FILE* f = fopen(argv[3], "wb");
  fwrite(hs.data(), 4, hs.size(), f); fwrite(ht.data(), 4, ht.size(), f);
  fclose(f);
  double cs = 0; for (float v : hs) cs += v; for (float v : ht) cs += v;
  printf("checksum %.6f\n", cs);
  return 0;
}
