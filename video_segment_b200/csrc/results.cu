// results.cu -- label volume, N4 fix-up, run-length rasterisation, neighbour pairs, sm_100a.
// Replaces FlattenUnionFind (segmentation/segmentation_graph.h:596-629), the id-image fill,
// EnforceN4Connectivity and the RLE loop of DenseSegmentationGraph::ObtainResults
// (segmentation/dense_segmentation_graph.h:467-579,1303-1337) and the edge walk of
// DetermineNeighborIdsImpl (segmentation/segmentation_graph.h:466-496).
#include "common.cuh"
#include "results.cuh"

namespace vsb {

// ---- FlattenUnionFind: every node -> its representative (pointer jumping to the root) ----
__global__ void flatten_kernel(const int* __restrict__ parent, int* __restrict__ labels, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int x = (int)i, p = parent[x];
    while (p != x) { x = p; p = parent[x]; }
    labels[i] = x;
  }
}

int launch_flatten(const int* parent_in, int* /*unused*/, int* labels, long long n, cudaStream_t s) {
  const int* parent = parent_in;
  flatten_kernel<<<148 * 8, 256, 0, s>>>(parent, labels, n);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- EnforceN4Connectivity (dense_segmentation_graph.h:1303-1337) ----
// The reference scans rows top to bottom and, inside a row, left to right; at pixel (i, j) with
// id = img(i, j) it overwrites the pixel BELOW, img(i+1, j), when the diagonal neighbour below
// has the id but neither the horizontal neighbour nor the pixel below has (two tests: below-left
// with left, then below-right with right).  Row i is final when its scan starts (only row i+1 is
// written), so rows are processed one after the other; inside a row the only carried state is
// whether img(i+1, j-1) was overwritten one step earlier (the "below-left" value of step j).
// The per-column transfer function therefore is a map {0,1} -> {0,1} ("was the left column's
// lower pixel swapped") which is composed with a block-wide scan: exact, row-parallel.
//
// One block per frame slice; img has no border: out-of-frame neighbours compare as -1.
__global__ void __launch_bounds__(1024) n4_kernel(int* __restrict__ labels, int w, int h, int n_slices,
                                                  const int* __restrict__ slice_ids,
                                                  int* __restrict__ size_adjust) {
  extern __shared__ unsigned char sm_raw[];
  // per column: f0 = swapped(j) if left not swapped, f1 = swapped(j) if left swapped
  unsigned char* f = sm_raw;                                  // [w] packed 2 bits
  unsigned char* sw = sm_raw + ((w + 15) & ~15);              // [w] resolved swap flag
  __shared__ unsigned char warp_fn[32];
  __shared__ unsigned char warp_in[32];
  const int slice = slice_ids[blockIdx.x];
  int* img = labels + (size_t)slice * w * h;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int per = (w + nthr - 1) / nthr;                      // contiguous columns per thread
  const int j0 = tid * per, j1 = min(w, j0 + per);
  for (int i = 0; i + 1 < h; ++i) {
    const int* row = img + (size_t)i * w;
    int* below = img + (size_t)(i + 1) * w;
    // transfer functions
    unsigned char comp = 0x2;   // identity map encoded as bits: out(0) = bit0, out(1) = bit1 -> id = 0b10
    for (int j = j0; j < j1; ++j) {
      const int id = row[j];
      const int left = (j > 0) ? row[j - 1] : -1;
      const int right = (j + 1 < w) ? row[j + 1] : -1;
      const int b = below[j];
      const int bl_orig = (j > 0) ? below[j - 1] : -1;
      const int bl_swapped = (j > 0) ? row[j - 1] : -1;       // a swap writes the upper pixel's id
      const int br = (j + 1 < w) ? below[j + 1] : -1;
      unsigned char fn = 0;
#pragma unroll
      for (int s = 0; s < 2; ++s) {
        const int bl = s ? bl_swapped : bl_orig;
        int cur_b = b;
        bool swapped = false;
        if (bl == id && left != id && cur_b != id) { cur_b = id; swapped = true; }
        if (br == id && right != id && cur_b != id) { cur_b = id; swapped = true; }
        fn |= (unsigned char)((swapped ? 1 : 0) << s);
      }
      f[j] = fn;
      // compose: comp = fn o comp
      const unsigned char c0 = (fn >> (comp & 1)) & 1, c1 = (fn >> ((comp >> 1) & 1)) & 1;
      comp = (unsigned char)(c0 | (c1 << 1));
    }
    if (j0 >= j1) comp = 0x2;
    // block-wide inclusive composition scan over threads (function composition is associative)
    const int lane = tid & 31, wid = tid >> 5;
    unsigned char inc = comp;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned char prev = (unsigned char)__shfl_up_sync(0xffffffffu, (unsigned)inc, o);
      if (lane >= o) {   // inc = inc o prev  (prev applied first)
        const unsigned char c0 = (inc >> (prev & 1)) & 1, c1 = (inc >> ((prev >> 1) & 1)) & 1;
        inc = (unsigned char)(c0 | (c1 << 1));
      }
    }
    if (lane == 31) warp_fn[wid] = inc;
    __syncthreads();
    if (wid == 0) {
      const int nw = (nthr + 31) >> 5;
      unsigned char wf = (lane < nw) ? warp_fn[lane] : (unsigned char)0x2;
      unsigned char winc = wf;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned char prev = (unsigned char)__shfl_up_sync(0xffffffffu, (unsigned)winc, o);
        if (lane >= o) {
          const unsigned char c0 = (winc >> (prev & 1)) & 1, c1 = (winc >> ((prev >> 1) & 1)) & 1;
          winc = (unsigned char)(c0 | (c1 << 1));
        }
      }
      // state entering warp `lane` = (composition of all previous warps)(0)
      const unsigned char prev = (unsigned char)__shfl_up_sync(0xffffffffu, (unsigned)winc, 1);
      warp_in[lane] = (lane == 0) ? 0 : (prev & 1);
    }
    __syncthreads();
    // state entering this thread: apply the exclusive prefix inside the warp to the warp's input
    unsigned char excl = (unsigned char)__shfl_up_sync(0xffffffffu, (unsigned)inc, 1);
    unsigned char state = warp_in[wid];
    if (lane > 0) state = (excl >> state) & 1;
    for (int j = j0; j < j1; ++j) {
      state = (f[j] >> state) & 1;
      sw[j] = state;
    }
    __syncthreads();
    for (int j = j0; j < j1; ++j) {
      if (sw[j]) {
        const int id = row[j];
        // size bookkeeping of the reference: every executed test decrements the overwritten id
        // and increments the new one; when both tests fire the second overwrites id by id
        // (net zero), so one adjustment per swapped pixel is exact.
        atomicAdd(&size_adjust[below[j]], -1);
        atomicAdd(&size_adjust[id], 1);
      }
    }
    __syncthreads();
    for (int j = j0; j < j1; ++j)
      if (sw[j]) below[j] = row[j];
    __syncthreads();
  }
}

int launch_n4(int* labels, int w, int h, int n_slices, const int* dev_slice_ids, int* size_adjust, cudaStream_t s) {
  const size_t sm = 2 * ((w + 15) & ~15);
  n4_kernel<<<n_slices, 1024, sm, s>>>(labels, w, h, n_slices, dev_slice_ids, size_adjust);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- RLE (dense_segmentation_graph.h:533-559): run heads -> (slice, y, left_x, right_x, id) ----
// pass 1 counts runs per row, pass 2 (after an exclusive scan on the host side of the row counts
// being small: rows = slices * h) writes the runs in raster order.
__global__ void rle_count_kernel(const int* __restrict__ labels, int w, int h, const int* __restrict__ slice_ids,
                                 int n_rows_total, unsigned* __restrict__ row_counts) {
  const int warps_per_block = blockDim.x >> 5;
  const int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (r >= n_rows_total) return;
  const int lane = threadIdx.x & 31;
  const int slice = slice_ids[r / h], y = r % h;
  const int* row = labels + ((size_t)slice * h + y) * w;
  unsigned c = 0;
  for (int j = lane; j < w; j += 32) c += (j == 0 || row[j] != row[j - 1]) ? 1u : 0u;
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) row_counts[r] = c;
}

__global__ void rle_write_kernel(const int* __restrict__ labels, int w, int h, const int* __restrict__ slice_ids,
                                 int n_rows_total, const unsigned* __restrict__ row_offsets,
                                 RunRec* __restrict__ runs) {
  const int warps_per_block = blockDim.x >> 5;
  const int r = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  if (r >= n_rows_total) return;
  const int lane = threadIdx.x & 31;
  const int slice_pos = r / h, y = r % h;
  const int slice = slice_ids[slice_pos];
  const int* row = labels + ((size_t)slice * h + y) * w;
  unsigned base = row_offsets[r];
  for (int j0 = 0; j0 < w; j0 += 32) {
    const int j = j0 + lane;
    const bool head = (j < w) && (j == 0 || row[j] != row[j - 1]);
    const unsigned m = __ballot_sync(0xffffffffu, head);
    if (head) {
      const unsigned k = base + __popc(m & ((1u << lane) - 1u));
      runs[k].slice = slice; runs[k].y = y; runs[k].left_x = j; runs[k].id = row[j];
      if (k > row_offsets[r]) runs[k - 1].right_x = j - 1;   // previous run of this row ends here
    }
    base += __popc(m);
  }
  if (lane == 0) runs[base - 1].right_x = w - 1;
}

int launch_rle_count(const int* labels, int w, int h, const int* dev_slice_ids, int n_slices, unsigned* row_counts,
                     cudaStream_t s) {
  const int rows = n_slices * h;
  rle_count_kernel<<<(rows + 7) / 8, 256, 0, s>>>(labels, w, h, dev_slice_ids, rows, row_counts);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}
int launch_rle_write(const int* labels, int w, int h, const int* dev_slice_ids, int n_slices,
                     const unsigned* row_offsets, RunRec* runs, cudaStream_t s) {
  const int rows = n_slices * h;
  rle_write_kernel<<<(rows + 7) / 8, 256, 0, s>>>(labels, w, h, dev_slice_ids, rows, row_offsets, runs);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- gather per-region info for a list of representative ids ----
__global__ void gather_region_info_kernel(const int* __restrict__ ids, int n, const RegionRec* __restrict__ rec,
                                          const int* __restrict__ size_adjust, int2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int id = ids[i];
  out[i] = make_int2(rec[id].sz + size_adjust[id], rec[id].con);
}
int launch_gather_region_info(const int* ids, int n, const RegionRec* rec, const int* size_adjust, int2* out,
                              cudaStream_t s) {
  if (n == 0) return 0;
  gather_region_info_kernel<<<(n + 255) / 256, 256, 0, s>>>(ids, n, rec, size_adjust, out);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- relabel intervals (tubes split off by EnforceSpatialConnectedness, :866-893) ----
__global__ void relabel_kernel(const RunRec* __restrict__ runs, int n, int w, int h, int* __restrict__ node_labels) {
  const int r = blockIdx.x;
  if (r >= n) return;
  const RunRec rr = runs[r];
  int* row = node_labels + ((size_t)rr.slice * h + rr.y) * w;
  for (int x = rr.left_x + threadIdx.x; x <= rr.right_x; x += blockDim.x) row[x] = rr.id;
}
int launch_relabel(const RunRec* runs, int n, int w, int h, int* node_labels, cudaStream_t s) {
  if (n == 0) return 0;
  relabel_kernel<<<n, 64, 0, s>>>(runs, n, w, h, node_labels);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- neighbour pairs (DetermineNeighborIdsImpl): distinct (label_a, label_b) over all edges ----
__device__ __forceinline__ void pair_insert(unsigned long long* table, unsigned cap_mask, int a, int b,
                                            unsigned long long* out, unsigned long long* out_count,
                                            unsigned long long out_cap) {
  const unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
  const unsigned long long key = ((unsigned long long)lo << 32) | hi;
  unsigned hsh = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 32);
  for (unsigned probe = 0; probe <= cap_mask; ++probe) {
    const unsigned slot = (hsh + probe) & cap_mask;
    const unsigned long long cur = table[slot];
    if (cur == key) return;
    if (cur == ~0ull) {
      const unsigned long long old = atomicCAS(&table[slot], ~0ull, key);
      if (old == ~0ull) {
        const unsigned long long k = atomicAdd(out_count, 1ull);
        if (k < out_cap) out[k] = key;
        return;
      }
      if (old == key) return;
    }
  }
}

__global__ void neighbor_pairs_kernel(const int* __restrict__ roots, const int* __restrict__ labels, int w, int h, int slots,
                                      const float* __restrict__ flows, int virtual_slot0,
                                      unsigned long long* __restrict__ table, unsigned cap_mask,
                                      unsigned long long* __restrict__ out, unsigned long long* __restrict__ out_count,
                                      unsigned long long out_cap) {
  const long long n = (long long)w * h;
  const long long total = n * slots;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int slot = (int)(i / n), pix = (int)(i % n);
    const int x = pix % w, y = pix / w;
    // An edge contributes a neighbour pair iff its endpoints ended in different union-find
    // regions (edges inside one region were dropped or merged, segmentation_graph.h:375-440);
    // the pair is reported on the labels after the tube split (EnforceSpatialConnectedness).
    const int ra = roots[i], la = labels[i];
    int last = -2;
#define VSB_PAIR(J)                                                                         \
    { const long long j_ = (J); const int rb = roots[j_];                                   \
      if (rb != ra) { const int lb = labels[j_];                                            \
        if (lb != la && lb != last) { pair_insert(table, cap_mask, la, lb, out, out_count, out_cap); last = lb; } } }
    // spatial edges R, B, BL, BR (virtual slots have none, dense_segmentation_graph.h:327-367)
    if (!(virtual_slot0 && slot == 0)) {
      if (x + 1 < w) VSB_PAIR(i + 1)
      if (y + 1 < h) {
        VSB_PAIR(i + w)
        if (x > 0) VSB_PAIR(i + w - 1)
        if (x + 1 < w) VSB_PAIR(i + w + 1)
      }
    }
    if (slot > 0) {
      int px = x, py = y;
      if (flows) {
        const float* f = flows + ((size_t)slot * n + pix) * 2;
        px = max(0, min(w - 1, (int)((float)x + f[0])));
        py = max(0, min(h - 1, (int)((float)y + f[1])));
      }
      const long long pbase = (long long)(slot - 1) * n;
      for (int dy = -1; dy <= 1; ++dy)
        for (int dx = -1; dx <= 1; ++dx) {
          const int xx = px + dx, yy = py + dy;
          if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
          VSB_PAIR(pbase + (long long)yy * w + xx)
        }
    }
#undef VSB_PAIR
  }
}

// No-flow variant: one thread per voxel, one grid row per image row (32-bit index arithmetic, the
// 13 neighbour loads of a warp fall into 5 cache lines).
__global__ void __launch_bounds__(256) neighbor_pairs_rows_kernel(const int* __restrict__ roots, const int* __restrict__ labels,
                                                                  int w, int h, int virtual_slot0,
                                                                  unsigned long long* __restrict__ table, unsigned cap_mask,
                                                                  unsigned long long* __restrict__ out,
                                                                  unsigned long long* __restrict__ out_count,
                                                                  unsigned long long out_cap) {
  // The 256 voxels of a CTA meet a handful of distinct pairs, thousands of times: they are de-duplicated in a
  // shared-memory hash first and only the distinct ones go to the global table (one dependent L2 access each).
  constexpr unsigned kLocal = 512;
  __shared__ unsigned long long stab[kLocal];
  for (unsigned k = threadIdx.x; k < kLocal; k += blockDim.x) stab[k] = ~0ull;
  __syncthreads();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int row = blockIdx.y;                     // slot * h + y
  const int slot = row / h, y = row - slot * h;
  const size_t base = (size_t)row * w;
  const int* r0 = roots + base;
  const int* l0 = labels + base;
  auto local_insert = [&](int a, int b) {
    const unsigned lo = (unsigned)min(a, b), hi = (unsigned)max(a, b);
    const unsigned long long key = ((unsigned long long)lo << 32) | hi;
    unsigned hsh = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 40);
    for (unsigned probe = 0; probe < 16; ++probe) {
      const unsigned slot_ = (hsh + probe) & (kLocal - 1);
      const unsigned long long cur = stab[slot_];
      if (cur == key) return;
      if (cur == ~0ull) {
        const unsigned long long old = atomicCAS(&stab[slot_], ~0ull, key);
        if (old == ~0ull || old == key) return;
      }
    }
    pair_insert(table, cap_mask, a, b, out, out_count, out_cap);     // local table crowded: straight to the global one
  };
  if (x < w) {
  const int ra = __ldg(&r0[x]);
  int la = -1, last = -2;
#define VSB_PAIR2(RP, LP, XX)                                                               \
  { const int rb = __ldg(&(RP)[XX]);                                                        \
    if (rb != ra) { if (la == -1) la = __ldg(&l0[x]); const int lb = __ldg(&(LP)[XX]);      \
      if (lb != la && lb != last) { local_insert(la, lb); last = lb; } } }
  if (!(virtual_slot0 && slot == 0)) {
    if (x + 1 < w) VSB_PAIR2(r0, l0, x + 1)
    if (y + 1 < h) {
      const int* r1 = r0 + w; const int* l1 = l0 + w;
      VSB_PAIR2(r1, l1, x)
      if (x > 0) VSB_PAIR2(r1, l1, x - 1)
      if (x + 1 < w) VSB_PAIR2(r1, l1, x + 1)
    }
  }
  if (slot > 0) {
    const size_t pbase = base - (size_t)h * w;    // same pixel, previous slot
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = y + dy;
      if (yy < 0 || yy >= h) continue;
      const int* rp = roots + pbase + (ptrdiff_t)dy * w;
      const int* lp = labels + pbase + (ptrdiff_t)dy * w;
      if (x > 0) VSB_PAIR2(rp, lp, x - 1)
      VSB_PAIR2(rp, lp, x)
      if (x + 1 < w) VSB_PAIR2(rp, lp, x + 1)
    }
  }
#undef VSB_PAIR2
  }
  __syncthreads();
  for (unsigned k = threadIdx.x; k < kLocal; k += blockDim.x) {
    const unsigned long long key = stab[k];
    if (key != ~0ull) pair_insert(table, cap_mask, (int)(key >> 32), (int)(key & 0xffffffffu), out, out_count, out_cap);
  }
}

int launch_neighbor_pairs(const int* roots, const int* labels, int w, int h, int slots, const float* flows, int virtual_slot0,
                          unsigned long long* table, unsigned table_cap_pow2, unsigned long long* out,
                          unsigned long long* out_count, unsigned long long out_cap, cudaStream_t s) {
  VSB_CUDA_OK(cudaMemsetAsync(table, 0xff, sizeof(unsigned long long) * (size_t)table_cap_pow2, s));
  VSB_CUDA_OK(cudaMemsetAsync(out_count, 0, sizeof(unsigned long long), s));
  if (!flows && (long long)slots * h <= 65535) {
    dim3 grid((w + 255) / 256, slots * h);
    neighbor_pairs_rows_kernel<<<grid, 256, 0, s>>>(roots, labels, w, h, virtual_slot0, table, table_cap_pow2 - 1, out, out_count, out_cap);
  } else {
    neighbor_pairs_kernel<<<148 * 8, 256, 0, s>>>(roots, labels, w, h, slots, flows, virtual_slot0, table, table_cap_pow2 - 1,
                                                 out, out_count, out_cap);
  }
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// single-block exclusive scan (n up to a few hundred thousand)
__global__ void __launch_bounds__(1024) scan_u32_kernel(const unsigned* __restrict__ in, unsigned* __restrict__ out,
                                                        unsigned* __restrict__ total, int n) {
  __shared__ unsigned warp_sums[32];
  const int tid = threadIdx.x;
  const int per = (n + 1023) / 1024;
  const int b0 = tid * per, b1 = min(n, b0 + per);
  unsigned s = 0;
  for (int i = b0; i < b1; ++i) s += in[i];
  unsigned inc = s;
  const int lane = tid & 31, wid = tid >> 5;
  for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += t; }
  if (lane == 31) warp_sums[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    unsigned v = warp_sums[lane];
    for (int o = 1; o < 32; o <<= 1) { const unsigned t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
    warp_sums[lane] = v;
  }
  __syncthreads();
  unsigned base = (wid ? warp_sums[wid - 1] : 0u) + inc - s;
  for (int i = b0; i < b1; ++i) { const unsigned c = in[i]; out[i] = base; base += c; }
  if (tid == 0) *total = warp_sums[31];
}
int launch_scan_u32(const unsigned* in, unsigned* out_exclusive, unsigned* total, int n, cudaStream_t s) {
  scan_u32_kernel<<<1, 1024, 0, s>>>(in, out_exclusive, total, n);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void fill_i32_kernel(int* p, int v, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = v;
}
int launch_fill_i32(int* p, int value, long long n, cudaStream_t s) {
  fill_i32_kernel<<<148 * 4, 256, 0, s>>>(p, value, n);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}
__global__ void iota_kernel(int* p, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) p[i] = (int)i;
}
int launch_init_iota(int* p, long long n, cudaStream_t s) {
  iota_kernel<<<148 * 4, 256, 0, s>>>(p, n);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}
// certification scratch of the merge kernel (common.cuh: NodeScratch), idle state
__global__ void init_hull_kernel(NodeScratch* sc, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    int4* q = reinterpret_cast<int4*>(&sc[i]);
    q[0] = make_int4(0x7f7f7f7f, 0x7f7f7f7f, 0x7f7f7f7f, 0);
    q[1] = make_int4(0, 0, 0, kNoCon);
    q[2] = make_int4(0, -1, -1, 0);
    q[3] = make_int4(-1, -1, 0, 0);       // num = ~0 (epoch-tagged big-big key, see merge.cu), claim, frozen
  }
}
int launch_init_hull(NodeScratch* hull, long long n_nodes, cudaStream_t s) {
  init_hull_kernel<<<148 * 8, 256, 0, s>>>(hull, n_nodes);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vsb
