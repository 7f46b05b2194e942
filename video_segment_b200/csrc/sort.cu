// sort.cu -- stable bucket (counting) sort of one chunk graph's edge lists, sm_100a.
// The reference never sorts: FastSegmentationGraph::AddEdge (segmentation/segmentation_graph.h:
// 158-162) appends every edge to bucket_lists_[list][bucket] at insert time and SegmentGraph
// (:367-374) walks bucket ascending, then list ascending, then insertion (raster, direction)
// order.  A stable counting sort on the 11-bit bucket key of the edges enumerated in
// (list, pixel, direction) order reproduces exactly that traversal order.
//
//   pass 1  hist     : one warp per 128 Ki-element range -> per-range bucket counts
//   pass 2  scan     : exclusive scan in (bucket-major, range-minor) order
//   pass 3  scatter  : one warp per range walks its elements in order; lanes with equal buckets
//                      are ranked with __match_any_sync, so the output stays stable
// Payload = 32-bit edge code (rank of the edge in (list, pixel, direction) order, common.cuh); the weights themselves
// are not moved.
#include "common.cuh"

namespace vsb {

#ifndef VSB_SORT_RANGE_LOG2
#define VSB_SORT_RANGE_LOG2 15
#endif
constexpr int kRangeLog2 = VSB_SORT_RANGE_LOG2;     // 32 Ki elements per warp (round 1: 128 Ki -- four times fewer warps in flight, scatter 5.2 ms per 1080p chunk under ncu)
constexpr unsigned kRangeElems = 1u << kRangeLog2;
constexpr int kWarpsPerBlock = 4;

struct RangeDesc {
  const float* base;     // weights of the list
  unsigned start;        // first element of the range inside the list
  unsigned count;        // elements in the range
  unsigned list;         // bucket-list index q
  unsigned code_base;    // edge code of the list's first element (edge_list_offset)
  unsigned nd;           // 4 (spatial) or 9 (temporal)
};

static unsigned num_ranges_for(int num_lists, int w, int h, const float* const* ptrs) {
  unsigned r = 0;
  const unsigned long long n = (unsigned long long)w * h;
  for (int q = 0; q < num_lists; ++q) {
    if (ptrs && !ptrs[q]) continue;
    const unsigned long long e = n * ((q & 1) ? 9 : 4);
    r += (unsigned)((e + kRangeElems - 1) >> kRangeLog2);
  }
  return r;
}

size_t sort_scratch_bytes(int num_lists, int w, int h) {
  const unsigned nr = num_ranges_for(num_lists, w, h, nullptr);
  size_t b = 0;
  b += (size_t)nr * sizeof(RangeDesc);            // range table
  b = (b + 255) & ~(size_t)255;
  b += (size_t)nr * kNumBuckets * sizeof(unsigned);   // hist / offsets [bucket][range]
  b += (kNumBuckets + 1) * sizeof(unsigned long long);   // row sums
  return b + 1024;
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) hist_kernel(const RangeDesc* __restrict__ ranges,
                                                                   unsigned num_ranges,
                                                                   unsigned* __restrict__ hist) {
  __shared__ unsigned cnt[kWarpsPerBlock][kNumBuckets];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned r = blockIdx.x * kWarpsPerBlock + warp;
  for (int i = lane; i < kNumBuckets; i += 32) cnt[warp][i] = 0;
  __syncwarp();
  if (r < num_ranges) {
    const RangeDesc rd = ranges[r];
    const float* p = rd.base + rd.start;
    for (unsigned i = lane; i < rd.count; i += 32) {
      const float wgt = __ldg(&p[i]);
      if (wgt >= 0.f) atomicAdd(&cnt[warp][bucket_of(wgt)], 1u);
    }
    __syncwarp();
    for (int i = lane; i < kNumBuckets; i += 32) hist[(size_t)i * num_ranges + r] = cnt[warp][i];
  }
}

__device__ __forceinline__ unsigned long long block_exclusive_scan(unsigned long long v, unsigned long long* smem,
                                                                   unsigned long long* total) {
  // blockDim.x == 256
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long inc = v;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) smem[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    unsigned long long s = (lane < 8) ? smem[lane] : 0;
    for (int o = 1; o < 8; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < 8) smem[8 + lane] = s;
  }
  __syncthreads();
  const unsigned long long warp_base = warp ? smem[8 + warp - 1] : 0;
  if (total) *total = smem[8 + 7];
  __syncthreads();
  return warp_base + inc - v;
}

// one block per bucket: row sum
__global__ void __launch_bounds__(256) rowsum_kernel(const unsigned* __restrict__ hist, unsigned num_ranges,
                                                     unsigned long long* __restrict__ rowsum) {
  __shared__ unsigned long long sm[16];
  const unsigned* row = hist + (size_t)blockIdx.x * num_ranges;
  unsigned long long s = 0;
  for (unsigned i = threadIdx.x; i < num_ranges; i += 256) s += row[i];
  unsigned long long tot;
  block_exclusive_scan(s, sm, &tot);
  if (threadIdx.x == 0) rowsum[blockIdx.x] = tot;
}

// single block: bucket_start[b] = exclusive scan of row sums, bucket_start[2048] = total
__global__ void __launch_bounds__(256) bucket_start_kernel(const unsigned long long* __restrict__ rowsum,
                                                           unsigned long long* __restrict__ bucket_start) {
  __shared__ unsigned long long sm[16];
  constexpr int per = kNumBuckets / 256;
  unsigned long long loc[per], s = 0;
  for (int k = 0; k < per; ++k) { loc[k] = rowsum[threadIdx.x * per + k]; s += loc[k]; }
  unsigned long long tot;
  unsigned long long base = block_exclusive_scan(s, sm, &tot);
  for (int k = 0; k < per; ++k) { bucket_start[threadIdx.x * per + k] = base; base += loc[k]; }
  if (threadIdx.x == 0) bucket_start[kNumBuckets] = tot;
}

// one block per bucket: in-place exclusive scan of the row + bucket_start[b]
__global__ void __launch_bounds__(256) rowscan_kernel(unsigned* __restrict__ hist, unsigned num_ranges,
                                                      const unsigned long long* __restrict__ bucket_start) {
  __shared__ unsigned long long sm[16];
  unsigned* row = hist + (size_t)blockIdx.x * num_ranges;
  const unsigned per = (num_ranges + 255) / 256;
  const unsigned b0 = threadIdx.x * per, b1 = min(num_ranges, b0 + per);
  unsigned long long s = 0;
  for (unsigned i = b0; i < b1; ++i) s += row[i];
  unsigned long long base = block_exclusive_scan(s, sm, nullptr) + bucket_start[blockIdx.x];
  for (unsigned i = b0; i < b1; ++i) {
    const unsigned c = row[i];
    row[i] = (unsigned)base;
    base += c;
  }
}

__global__ void __launch_bounds__(32 * kWarpsPerBlock) scatter_kernel(const RangeDesc* __restrict__ ranges,
                                                                      unsigned num_ranges,
                                                                      const unsigned* __restrict__ offsets,
                                                                      unsigned n_pix,
                                                                      uint32_t* __restrict__ codes) {
  __shared__ unsigned pos[kWarpsPerBlock][kNumBuckets];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned r = blockIdx.x * kWarpsPerBlock + warp;
  if (r >= num_ranges) return;                      // whole warp exits together
  for (int i = lane; i < kNumBuckets; i += 32) pos[warp][i] = offsets[(size_t)i * num_ranges + r];
  __syncwarp();
  const RangeDesc rd = ranges[r];
  const float* p = rd.base + rd.start;
  const unsigned lt_mask = (1u << lane) - 1u;
  for (unsigned i0 = 0; i0 < rd.count; i0 += 32) {
    const unsigned i = i0 + lane;
    float wgt = -1.f;
    if (i < rd.count) wgt = __ldg(&p[i]);
    const bool valid = wgt >= 0.f;
    const unsigned b = valid ? (unsigned)bucket_of(wgt) : 0xFFFFu;
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    unsigned base = 0;
    if (valid) base = pos[warp][b];
    __syncwarp();
    if (valid) {
      const unsigned rank = __popc(peers & lt_mask);
      if (rank == 0) pos[warp][b] = base + __popc(peers);
      codes[base + rank] = rd.code_base + rd.start + i;      // element index inside the list = pixel * nd + dir
    }
    __syncwarp();
  }
}

int launch_sort_edges(const float* const* seg_ptrs, int num_lists, int w, int h, uint32_t* codes,
                      unsigned long long* bucket_start, void* scratch, size_t scratch_bytes, cudaStream_t s) {
  const unsigned long long n = (unsigned long long)w * h;
  if (!edge_codes_fit(num_lists, n)) {
    set_error("sort_edges: %d lists x %llu pixels overflow the 32-bit edge code", num_lists, n);
    return 5;
  }
  const unsigned nr = num_ranges_for(num_lists, w, h, seg_ptrs);
  if (nr == 0) { set_error("sort_edges: no edge lists"); return 1; }
  if (scratch_bytes < sort_scratch_bytes(num_lists, w, h)) { set_error("sort_edges: scratch too small"); return 1; }
  // host range table
  static thread_local RangeDesc* h_tab = nullptr;
  static thread_local unsigned h_cap = 0;
  if (h_cap < nr) {
    if (h_tab) cudaFreeHost(h_tab);
    VSB_CUDA_OK(cudaMallocHost(&h_tab, sizeof(RangeDesc) * nr));
    h_cap = nr;
  } else {
    // the previous async copy from this pinned table must have completed before we overwrite it
    VSB_CUDA_OK(cudaStreamSynchronize(s));
  }
  unsigned k = 0;
  for (int q = 0; q < num_lists; ++q) {
    if (!seg_ptrs[q]) continue;
    const unsigned nd = (q & 1) ? 9 : 4;
    const unsigned long long e = n * nd;
    for (unsigned long long st = 0; st < e; st += kRangeElems) {
      RangeDesc rd;
      rd.base = seg_ptrs[q];
      rd.start = (unsigned)st;
      rd.count = (unsigned)((e - st < kRangeElems) ? (e - st) : kRangeElems);
      rd.list = (unsigned)q;
      rd.code_base = edge_list_offset(q, (uint32_t)n);
      rd.nd = nd;
      h_tab[k++] = rd;
    }
  }
  char* sc = (char*)scratch;
  RangeDesc* d_tab = (RangeDesc*)sc;
  size_t off = ((size_t)nr * sizeof(RangeDesc) + 255) & ~(size_t)255;
  unsigned* hist = (unsigned*)(sc + off);
  off += (size_t)nr * kNumBuckets * sizeof(unsigned);
  unsigned long long* rowsum = (unsigned long long*)(sc + off);
  VSB_CUDA_OK(cudaMemcpyAsync(d_tab, h_tab, sizeof(RangeDesc) * nr, cudaMemcpyHostToDevice, s));
  const unsigned blocks = (nr + kWarpsPerBlock - 1) / kWarpsPerBlock;
  hist_kernel<<<blocks, 32 * kWarpsPerBlock, 0, s>>>(d_tab, nr, hist);
  rowsum_kernel<<<kNumBuckets, 256, 0, s>>>(hist, nr, rowsum);
  bucket_start_kernel<<<1, 256, 0, s>>>(rowsum, bucket_start);
  rowscan_kernel<<<kNumBuckets, 256, 0, s>>>(hist, nr, bucket_start);
  scatter_kernel<<<blocks, 32 * kWarpsPerBlock, 0, s>>>(d_tab, nr, hist, (unsigned)n, codes);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vsb
