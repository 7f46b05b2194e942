// shape.cu -- K10 / K11 of the result stage on the device.
//
// Input: the scan intervals ("runs") of the label volume in raster order (results.cu: rle_*), row offsets included.
//   K11  run_cc_*      N4 connected components of every region in every frame (ConnectedComponents(raster, N4_CONNECT),
//                      segment_util/segmentation_util.cpp:1007-1101): lock-free union-find over the runs, a run links to
//                      the runs of the row above that carry its label and overlap it in x.  A component is named by
//                      its first run, so ascending names = the reference's "order of first scan interval".
//        sort_pairs    stable LSD radix sort (8-bit digits) of (key, run index): by component name for K10, by
//                      (frame, output rank of the region) for the final result order.
//   K10  group_runs_*  one RunGroup per run of equal keys in the sorted order: the group's scan intervals copied out
//                      contiguously and its ShapeMoments (ShapeMomentsFromRasterization, segmentation_util.cpp:652-693).
//                      The sums are floats accumulated interval by interval in raster order and the result is a field of
//                      the output message, so the "segmented reduce" keeps that order: one thread walks one group
//                      (groups are short -- a component of one frame -- and there are 10^5 of them).
//        relabel_groups pixels of components that became regions of their own (tubes.hpp) get their fresh label.
//
// HBM traffic per chunk is O(#runs) (20 B per run and pass), two to three orders of magnitude below the label volume.
#include "shape.cuh"

#include "shape_math.hpp"

#define VSB_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

namespace vsb {

// ---------------------------------------------------------------- K11: components over runs
__device__ __forceinline__ int cc_find(int* parent, int x) {
  // pointers only ever move to smaller run indices of the same tree, so halving with plain stores is safe under races
  for (;;) {
    const int p = ((volatile int*)parent)[x];
    if (p == x) return x;
    const int g = ((volatile int*)parent)[p];
    if (g != p) parent[x] = g;
    x = g;
  }
}
__device__ __forceinline__ void cc_unite(int* parent, int a, int b) {
  for (;;) {
    a = cc_find(parent, a);
    b = cc_find(parent, b);
    if (a == b) return;
    if (a < b) { const int t = a; a = b; b = t; }
    if (atomicCAS(&parent[a], a, b) == a) return;          // the larger root hangs under the smaller one
  }
}

__global__ void run_cc_link_kernel(const RunRec* __restrict__ runs, unsigned n_runs, const unsigned* __restrict__ row_offsets,
                                   int h, int slice0, int* parent) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_runs) return;
  const RunRec me = runs[i];
  if (me.y == 0) return;
  const int r = (me.slice - slice0) * h + me.y;
  unsigned lo = row_offsets[r - 1];
  const unsigned end = row_offsets[r];
  unsigned hi = end;
  while (lo < hi) {                                          // first run of the row above that reaches my left end
    const unsigned mid = (lo + hi) >> 1;
    if (runs[mid].right_x < me.left_x) lo = mid + 1; else hi = mid;
  }
  for (unsigned k = lo; k < end; ++k) {
    const RunRec up = runs[k];
    if (up.left_x > me.right_x) break;
    if (up.id == me.id) cc_unite(parent, (int)i, (int)k);
  }
}

__global__ void run_cc_name_kernel(int* parent, unsigned n_runs, unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_runs) return;
  keys[i] = (unsigned)cc_find(parent, (int)i);
  vals[i] = i;
}

int launch_run_components(const RunRec* runs, unsigned n_runs, const unsigned* row_offsets, int h, int slice0, int* parent,
                          unsigned* keys, unsigned* vals, cudaStream_t s) {
  if (n_runs == 0) return 0;
  VSB_RC(launch_init_iota(parent, n_runs, s));
  const unsigned blocks = (n_runs + 255) / 256;
  run_cc_link_kernel<<<blocks, 256, 0, s>>>(runs, n_runs, row_offsets, h, slice0, parent);
  run_cc_name_kernel<<<blocks, 256, 0, s>>>(parent, n_runs, keys, vals);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- stable radix sort of (key, value) pairs
constexpr int kSortTile = 2048;                              // elements per CTA: 8 rounds of 256 consecutive ones

__global__ void __launch_bounds__(256) radix_hist_kernel(const unsigned* __restrict__ keys, unsigned n, int shift,
                                                         unsigned* __restrict__ hist, unsigned n_tiles) {
  __shared__ unsigned h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const unsigned base = blockIdx.x * kSortTile;
  for (int r = 0; r < kSortTile / 256; ++r) {
    const unsigned i = base + r * 256 + threadIdx.x;
    if (i < n) atomicAdd(&h[(keys[i] >> shift) & 255u], 1u);
  }
  __syncthreads();
  hist[threadIdx.x * n_tiles + blockIdx.x] = h[threadIdx.x];   // digit-major: one exclusive scan gives every tile's bases
}

__global__ void __launch_bounds__(256) radix_scatter_kernel(const unsigned* __restrict__ keys_in, const unsigned* __restrict__ vals_in,
                                                            unsigned* __restrict__ keys_out, unsigned* __restrict__ vals_out,
                                                            unsigned n, int shift, const unsigned* __restrict__ bases,
                                                            unsigned n_tiles) {
  __shared__ unsigned next[256];                             // next free output slot of each digit for this tile
  __shared__ unsigned per_warp[8][256];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  next[tid] = bases[tid * n_tiles + blockIdx.x];
  const unsigned base = blockIdx.x * kSortTile;
  for (int r = 0; r < kSortTile / 256; ++r) {
    for (int k = 0; k < 8; ++k) per_warp[k][tid] = 0;
    __syncthreads();
    const unsigned i = base + r * 256 + tid;
    const bool live = i < n;
    const unsigned key = live ? keys_in[i] : 0u, val = live ? vals_in[i] : 0u;
    const unsigned digit = live ? ((key >> shift) & 255u) : 256u + lane;   // dead lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const unsigned before = __popc(peers & ((1u << lane) - 1u));
    if (live && before == 0) per_warp[warp][digit] = __popc(peers);
    __syncthreads();
    unsigned pos = 0;
    if (live) {
      pos = next[digit] + before;
      for (int k = 0; k < warp; ++k) pos += per_warp[k][digit];
    }
    __syncthreads();
    unsigned total = 0;
    for (int k = 0; k < 8; ++k) total += per_warp[k][tid];
    next[tid] += total;
    if (live) { keys_out[pos] = key; vals_out[pos] = val; }
    __syncthreads();
  }
}

// Sorts n pairs by the low `bits` bits of the key.  keys / vals: the input; *_alt: scratch of the same size; `hist`: 2 x
// 256 x ceil(n / 2048) words.  Returns through out_keys / out_vals which of the two buffers holds the result.
int launch_sort_pairs(unsigned* keys, unsigned* vals, unsigned* keys_alt, unsigned* vals_alt, unsigned n, int bits,
                      unsigned* hist, unsigned* scratch_total, unsigned** out_keys, unsigned** out_vals, cudaStream_t s) {
  *out_keys = keys; *out_vals = vals;
  if (n == 0) return 0;
  const unsigned tiles = (n + kSortTile - 1) / kSortTile;
  unsigned* bases = hist + (size_t)256 * tiles;
  for (int shift = 0; shift < bits; shift += 8) {
    radix_hist_kernel<<<tiles, 256, 0, s>>>(keys, n, shift, hist, tiles);
    VSB_RC(launch_scan_u32(hist, bases, scratch_total, (int)(256 * tiles), s));
    radix_scatter_kernel<<<tiles, 256, 0, s>>>(keys, vals, keys_alt, vals_alt, n, shift, bases, tiles);
    unsigned* t = keys; keys = keys_alt; keys_alt = t;
    t = vals; vals = vals_alt; vals_alt = t;
  }
  VSB_CUDA_OK(cudaGetLastError());
  *out_keys = keys; *out_vals = vals;
  return 0;
}

// ---------------------------------------------------------------- K10: groups of equal keys, their intervals and moments
constexpr int kHeadTile = 1024;

__global__ void __launch_bounds__(kHeadTile) group_heads_count_kernel(const unsigned* __restrict__ keys, unsigned n,
                                                                      unsigned* __restrict__ tile_counts) {
  const unsigned i = blockIdx.x * kHeadTile + threadIdx.x;
  const bool head = i < n && (i == 0 || keys[i] != keys[i - 1]);
  const int c = __syncthreads_count(head);
  if (threadIdx.x == 0) tile_counts[blockIdx.x] = (unsigned)c;
}

// position of every group's first element in the sorted order
__global__ void __launch_bounds__(kHeadTile) group_heads_write_kernel(const unsigned* __restrict__ keys, unsigned n,
                                                                      const unsigned* __restrict__ tile_bases,
                                                                      unsigned* __restrict__ head_pos) {
  __shared__ unsigned warp_heads[kHeadTile / 32];
  const unsigned i = blockIdx.x * kHeadTile + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool head = i < n && (i == 0 || keys[i] != keys[i - 1]);
  const unsigned m = __ballot_sync(0xffffffffu, head);
  if (lane == 0) warp_heads[warp] = __popc(m);
  __syncthreads();
  if (!head) return;
  unsigned g = tile_bases[blockIdx.x] + __popc(m & ((1u << lane) - 1u));
  for (int k = 0; k < warp; ++k) g += warp_heads[k];
  head_pos[g] = i;
}

// One warp per group (grid-stride over the groups): the lanes fetch 32 runs of the group at a time -- the gather
// through the sorted order is what costs -- and copy their intervals out; the moment sums then go through the 32 runs
// in order, every lane running the same accumulator on shuffled values, so the float result is the sequential one.
// pass 0 (components): tag = label of the runs, group_of_run[run] = index of its group.  pass 1 (result order): tag = key.
__global__ void __launch_bounds__(256) group_runs_kernel(const unsigned* __restrict__ keys, const unsigned* __restrict__ vals,
                                                         unsigned n, const RunRec* __restrict__ runs,
                                                         const unsigned* __restrict__ head_pos,
                                                         const unsigned* __restrict__ n_groups_dev, int pass,
                                                         RunGroup* __restrict__ groups, int* __restrict__ group_of_run,
                                                         int3* __restrict__ intervals) {
  const unsigned n_groups = *n_groups_dev;
  const int lane = threadIdx.x & 31;
  const unsigned warps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; g < n_groups; g += warps) {
    const unsigned begin = head_pos[g], end = g + 1 < n_groups ? head_pos[g + 1] : n;
    vsbs::MomentSum sum;
    int area = 0, label = 0, slice = 0;
    for (unsigned base = begin; base < end; base += 32) {
      const unsigned p = base + lane;
      const int cnt = (int)min(32u, end - base);
      int y = 0, lx = 0, rx = 0;
      if (p < end) {
        const unsigned run = vals[p];
        const RunRec rr = runs[run];
        y = rr.y; lx = rr.left_x; rx = rr.right_x;
        intervals[p] = make_int3(y, lx, rx);
        if (pass == 0) group_of_run[run] = (int)g;
        if (p == begin) { label = rr.id; slice = rr.slice; }
      }
      for (int j = 0; j < cnt; ++j) {
        const int yj = __shfl_sync(0xffffffffu, y, j), lj = __shfl_sync(0xffffffffu, lx, j), rj = __shfl_sync(0xffffffffu, rx, j);
        sum.add(yj, lj, rj);
        area += rj - lj + 1;
      }
    }
    if (lane == 0) {
      const vsbs::Moments mo = sum.mean();
      RunGroup out;
      out.first = (int)begin; out.count = (int)(end - begin); out.tag = pass == 0 ? label : (int)keys[begin]; out.slice = slice; out.area = area;
      out.mean_x = mo.mean_x; out.mean_y = mo.mean_y; out.xx = mo.xx; out.xy = mo.xy; out.yy = mo.yy;
      groups[g] = out;
    }
  }
}

// tile_counts / tile_bases: ceil(n / 1024) words each; head_pos: n words; n_groups (device) receives the number of
// groups.  `groups` must hold at least as many records as there are groups -- n is always enough.
int launch_group_runs(const unsigned* keys, const unsigned* vals, unsigned n, const RunRec* runs, int pass, unsigned* tile_counts,
                      unsigned* tile_bases, unsigned* head_pos, unsigned* n_groups, RunGroup* groups, int* group_of_run,
                      int3* intervals, cudaStream_t s) {
  if (n == 0) { VSB_CUDA_OK(cudaMemsetAsync(n_groups, 0, sizeof(unsigned), s)); return 0; }
  const unsigned tiles = (n + kHeadTile - 1) / kHeadTile;
  group_heads_count_kernel<<<tiles, kHeadTile, 0, s>>>(keys, n, tile_counts);
  VSB_RC(launch_scan_u32(tile_counts, tile_bases, n_groups, (int)tiles, s));
  group_heads_write_kernel<<<tiles, kHeadTile, 0, s>>>(keys, n, tile_bases, head_pos);
  group_runs_kernel<<<148 * 8, 256, 0, s>>>(keys, vals, n, runs, head_pos, n_groups, pass, groups, group_of_run, intervals);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------- result order and fresh labels
__global__ void result_keys_kernel(const RunRec* __restrict__ runs, unsigned n_runs, const int* __restrict__ group_of_run,
                                   const int* __restrict__ rank_of_group, int slice0, unsigned n_ranks,
                                   unsigned* __restrict__ keys, unsigned* __restrict__ vals) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_runs) return;
  keys[i] = (unsigned)(runs[i].slice - slice0) * n_ranks + (unsigned)rank_of_group[group_of_run[i]];
  vals[i] = i;
}
int launch_result_keys(const RunRec* runs, unsigned n_runs, const int* group_of_run, const int* rank_of_group, int slice0,
                       unsigned n_ranks, unsigned* keys, unsigned* vals, cudaStream_t s) {
  if (n_runs == 0) return 0;
  result_keys_kernel<<<(n_runs + 255) / 256, 256, 0, s>>>(runs, n_runs, group_of_run, rank_of_group, slice0, n_ranks, keys, vals);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void relabel_groups_kernel(const RunRec* __restrict__ runs, unsigned n_runs, const int* __restrict__ group_of_run,
                                      const int* __restrict__ label_of_group, int w, int h, int* __restrict__ node_labels) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_runs) return;
  const int fresh = label_of_group[group_of_run[i]];
  if (fresh < 0) return;
  const RunRec rr = runs[i];
  int* row = node_labels + ((size_t)rr.slice * h + rr.y) * w;
  for (int x = rr.left_x; x <= rr.right_x; ++x) row[x] = fresh;
}
int launch_relabel_groups(const RunRec* runs, unsigned n_runs, const int* group_of_run, const int* label_of_group, int w, int h,
                          int* node_labels, cudaStream_t s) {
  if (n_runs == 0) return 0;
  relabel_groups_kernel<<<(n_runs + 255) / 256, 256, 0, s>>>(runs, n_runs, group_of_run, label_of_group, w, h, node_labels);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vsb
