// common.cuh -- shared declarations of the sm_100a dense over-segmentation kernels.
// All arithmetic that feeds parity (pixel conversion, bilateral weights, edge
// weights, descriptor means) is compiled with -fmad=false so that every float
// operation rounds exactly like the reference's scalar x86-64 code.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace vsb {

constexpr int kNumBuckets = 2048;           // segmentation/dense_segmentation_graph.h:303
constexpr int kLutBins = (1 << 12) * 3;     // imagefilter/image_filter.cpp:237
constexpr int kBilateralRadius = 4;         // int(3.0f * 1.5f), image_filter.cpp:201
constexpr int kBilateralTaps = 49;          // i*i + j*j <= 16

// thread-local last error text, set by the launch wrappers
void set_error(const char* fmt, ...);

#define VSB_CUDA_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      ::vsb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 3;                                                                        \
    }                                                                                  \
  } while (0)

// Bucket index, FastSegmentationGraph::AddEdge (segmentation/segmentation_graph.h:158-162)
// with scale_ = 2048 / (1.0f + 1e-6f) (segmentation_graph.h:336).
__host__ __device__ inline float bucket_scale() { return 2048 / (1.0f + 1e-6f); }
__host__ __device__ inline int bucket_of(float w) {
  const float v = w * bucket_scale();
  return (int)(v < 2048.f ? v : 2048.f);
}

// Edge code: the edge's rank in (list, pixel, direction) order over ALL lists of the chunk graph, i.e.
// offset(list) + pixel * nd + dir with nd = 4 (spatial lists, even) / 9 (temporal lists, odd) and
// offset(list) = (list / 2) * 13 N + (list odd ? 4 N : 0).  list = reference bucket-list index (2*slot spatial,
// 2*slot-1 temporal; dense_segmentation_graph.h:962,1077).  Dense, so 3840x2160 x 21 slots (2.19 G edges) still
// fits 32 bits -- the reference's own (node, direction) packing overflows there (dense_segmentation_graph.h:314-324).
__host__ __device__ inline uint32_t edge_list_offset(int list, uint32_t n_pix) {
  return (uint32_t)(list >> 1) * 13u * n_pix + ((list & 1) ? 4u * n_pix : 0u);
}
__host__ __device__ inline uint32_t edge_code(int list, uint32_t n_pix, uint32_t pixel, uint32_t dir) {
  return edge_list_offset(list, n_pix) + pixel * ((list & 1) ? 9u : 4u) + dir;
}
__host__ __device__ inline void edge_decode(uint32_t code, uint32_t n_pix, int& list, uint32_t& pixel, uint32_t& dir) {
  const uint32_t pair = code / (13u * n_pix);
  uint32_t r = code - pair * 13u * n_pix;
  if (r < 4u * n_pix) { list = (int)(2u * pair); pixel = r >> 2; dir = r & 3u; }
  else { r -= 4u * n_pix; list = (int)(2u * pair + 1u); pixel = r / 9u; dir = r - pixel * 9u; }
}
// largest graph the 32-bit code can address: offset(last list) + its size <= 2^32 - 2 (0xFFFFFFFF marks executed entries)
__host__ inline bool edge_codes_fit(int num_lists, unsigned long long n_pix) {
  const unsigned long long total = (unsigned long long)(num_lists / 2) * 13ull * n_pix + ((num_lists & 1) ? 4ull * n_pix : 0ull);
  return total < 0xFFFFFFFFull;
}

// ---------------- launch wrappers (host) ----------------
// preprocess.cu
size_t preprocess_scratch_bytes();
int launch_preprocess(const uint8_t* bgr, int stride, int w, int h, int presmoothing, float* out,
                      void* scratch, cudaStream_t s);
// edges.cu
int launch_edge_build(const float* curr, const float* prev, const float* flow, int w, int h, bool l1,
                      float* spatial, float* temporal, cudaStream_t s);
// sort.cu
size_t sort_scratch_bytes(int num_lists, int w, int h);
int launch_sort_edges(const float* const* seg_ptrs, int num_lists, int w, int h, uint32_t* codes,
                      unsigned long long* bucket_start, void* scratch, size_t scratch_bytes,
                      cudaStream_t s);

// merge.cu : region record, 32 bytes (one sector per root access)
struct __align__(32) RegionRec {
  int sz;            // voxels
  int con;           // constraint id, -1 = unconstrained
  float d0, d1, d2;  // mean colour descriptor (ColorMeanDescriptorTraits)
  int fin;           // region_finalized
  int pad0, pad1;
};

// Per-node scratch of the window certification stage (merge.cu).  A small root uses its slot as the
// record of the sub-cluster it represents, a big root ("hub") as its absorption bound.  Idle state
// between windows: mn = 0x7f7f7f7f, mx = 0, flags = 0, con = kNoCon, mass = 0, hubs = -1, rbits = 0.
struct __align__(16) NodeScratch {
  int mn[3];                 // colour hull of the members (float bits; colours are >= 0)
  int mx[3];
  int flags;                 // kSc* bits
  int con;                   // single constraint id met so far (kNoCon = none)
  int mass;                  // voxels of the members / of the atoms a hub may absorb
  int hub0, hub1;            // big roots adjacent to the sub-cluster (-1 = none)
  int rbits;                 // hub: max distance to an absorbable atom (float bits)
  unsigned long long num;    // ordered rounds: earliest pending edge to another big region (epoch tagged, never reset)
  int claim;                 // window tag of the last once-per-atom claim (never reset)
  int frozen;                // window tag in which this hub's decision-relevant state is certified constant
};
constexpr int kNoCon = 0x7f7f7f7f;
constexpr size_t kScanQueueWords = 1024 * 2 * 2048;   // up to 1024 CTAs x 2 x kScanMax words

struct MergeParams {
  int w, h, slots;                 // graph geometry; nodes = slots * w * h
  int min_region_size;
  float force_merge_weight;        // 0.001f (L2) / 0.002f (L1), dense_segmentation.cpp:259-264
  int has_constraints;             // chunk > 0
  float y_merge, y_force, y_split; // squared-distance images of the distance gates (set by launch_merge): sqrtf(y) < 0.05f <=> y < y_merge,
                                   // (double)sqrtf(y) < 0.2 <=> y < y_force, sqrtf(y) > 0.15f <=> y >= y_split
  unsigned long long window_target, residual_split, segment_min;   // window sizing (defaults in merge.cu; VSB200_WINDOW_TARGET / _RESIDUAL_SPLIT / _SEGMENT_MIN override them for sweeps)
  int dev_flags;                   // development switches (VSB200_MERGE_FLAGS): 1 = no hub-pair certificates, 2 = fixed window target, 4 = no block-0 rounds, 16 = no group-parallel scans, 32 / 64 = diagnostics for constrained chunks, 128 = one-thread exact scans; default 17
  const float* flows;              // optional [slots][h][w][2] (slot 0 unused), nullable
  const uint32_t* codes;           // sorted edge codes
  const unsigned long long* bucket_start;   // [2049]
  int* parent;                     // [nodes]
  RegionRec* rec;                  // [nodes]
  unsigned long long* res;         // [nodes] epoch-tagged reservations
  unsigned long long* acc;         // [nodes][4] fixed-point accumulators: sz, d0, d1, d2
  int* cl;                         // [nodes] scratch cluster union-find
  NodeScratch* hull;               // [nodes] certification scratch of the current window (idle between windows)
  unsigned char* done;             // [max bucket edges] per-position done flags of the current bucket
  uint32_t* live_a;                // live edge buffers: (code, ru, rv, position)
  uint32_t* live_b;
  uint32_t* live_c;                // live_a = live list of the window, live_b / live_c = lists of the ordered rounds
  uint32_t* scan_queue;            // [kScanQueueWords] per-CTA (position, group) queues of the group-parallel exact scans
  unsigned long long live_cap;     // in triples
  unsigned long long* counters;    // [8] device counters
  unsigned long long* stats;       // [8] rounds, commits, safe merges, ...
  unsigned long long* trace;       // optional mapped host memory: live progress markers (debugging hangs)
  unsigned long long* debug;       // optional [2048][4]: ns, cumulative rounds, edges, pending after pruning
};
int launch_merge(const MergeParams& p, cudaStream_t s);
int launch_init_nodes(const float* frame, const int* constraint_ids, int slot, int w, int h, int* parent,
                      RegionRec* rec, cudaStream_t s);
int launch_init_virtual_nodes(const int* constraint_ids, int slot, int w, int h, int* parent, RegionRec* rec,
                              int* first_of_id, int max_id, cudaStream_t s);

// results.cu
int launch_flatten(const int* parent_in, int* unused, int* labels, long long n, cudaStream_t s);
const char* last_error();

}  // namespace vsb
