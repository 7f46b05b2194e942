// region_stage.cu -- hierarchical region stage behind the C ABI (vsb200_region_*), sm_100a.
// Takes the place of RegionSegmentationUnit / RegionSegmentation (segmentation/segmentation_unit.cpp:180-331,
// segmentation/region_segmentation.cpp:97-365) with the hierarchical half of Segmentation
// (segmentation/segmentation.cpp:80-270,305-773), RegionAgglomerationGraph
// (segmentation/region_segmentation_graph.cpp:33-503) and the three default region descriptors
// (segmentation/region_descriptor.cpp:83-135,377-553; segmentation/histograms.cpp).
//
// Split of the work:
//   device  per frame: region-id image painted from the over-segmentation's scan intervals; fused BGR->Lab +
//           interpolated Lab histogram of EVERY region at once (region_hist.cu); one 16-bin flow histogram per
//           (region, frame) (flow_hist_kernel); at a chunk-set boundary: normalisation, then all descriptor
//           arithmetic of the agglomeration -- merging two descriptor slots into a new one (merge_slots_kernel)
//           and the distances of all incident edges of a merge / of a whole level in one launch
//           (slot_distance_kernel: chi-square over 4000 Lab bins, per-frame flow chi-square, squared-OR
//           combination, size penaliser).  Descriptor slots are immutable once written, so the levels of the
//           hierarchy share them instead of copying (the reference deep-copies every descriptor per level).
//   host    the O(#regions) control: chunk-set arithmetic, neighbour lists, the bucket queue of the greedy
//           agglomeration (same order semantics as the reference: first edge of the lowest bucket, edges
//           appended at the end, lazily dropped when unmergeable), hierarchy levels, constraints between
//           chunk sets, result records.
// Sums that the reference runs in the iteration order of a std::unordered_map (sparse histograms) run in bin
// order here, and level-0 histograms are accumulated exactly (fixed point) instead of in float: distances
// agree to ~1e-5, so an edge close to a bucket boundary (1/2028 wide) may change bucket.  tests/ hold the
// stage to the oracle on structure (frames, level sizes, trees, sizes) and on partition similarity per level.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <deque>
#include <limits>
#include <memory>
#include <unordered_map>
#include <vector>

#include "../../include/vsb200.h"
#include "common.cuh"
#include "region_raster.hpp"
#include "region_kernels.cuh"

using namespace vsb;

namespace {

#define RS_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VSB200_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)
#define RS_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

constexpr int kFlowBinsMax = 32;

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
struct PaintRun { int y, lx, rx, idx; };

// one warp per scan interval
__global__ void paint_runs_kernel(const PaintRun* __restrict__ runs, int n, int w, int* __restrict__ img) {
  const int warp = (int)((blockIdx.x * (size_t)blockDim.x + threadIdx.x) >> 5);
  if (warp >= n) return;
  const PaintRun r = runs[warp];
  int* row = img + (size_t)r.y * w;
  for (int x = r.lx + (int)(threadIdx.x & 31u); x <= r.rx; x += 32) row[x] = r.idx;
}

// FlowDescriptor::AddFeatures (region_descriptor.cpp:434-463) -> VectorHistogram::AddVector (histograms.cpp:466-479):
// bin = NormAngle(x, y) * bins with NormAngle = float(atan2(y, x) / (2 pi + 1e-4) + 0.5) (double arithmetic: the
// reference calls the C library's atan2 / hypot), magnitude hypot(x, y) added to the bin.  acc: [regions][frames][bins]
// fixed point 2^-24 (exact, order independent), cnt: [regions][frames] vectors.
__global__ void __launch_bounds__(256) flow_hist_kernel(const float* __restrict__ flow, const int* __restrict__ ids, size_t n,
                                                        int n_regions, int frame_slot, int frame_cap, int bins,
                                                        unsigned long long* __restrict__ acc, unsigned* __restrict__ cnt) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int r = __ldg(&ids[i]);
    if (r < 0 || r >= n_regions) continue;
    const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + i);
    const float ang = (float)(atan2((double)f.y, (double)f.x) / (2.0 * 3.14159265358979323846 + 1e-4) + 0.5);
    const int bin = (int)(ang * (float)bins);
    const double mag = hypot((double)f.x, (double)f.y);
    const size_t slot = (size_t)r * frame_cap + frame_slot;
    atomicAdd(&acc[slot * bins + bin], (unsigned long long)__double2ll_rn(mag * 16777216.0));
    atomicAdd(&cnt[slot], 1u);
  }
}

// Descriptor slot layout (float words): [0, B) Lab histogram, normalised; then per frame of the chunk set: bins flow
// histogram values (normalised to one).  Meta (separate arrays): weight_sum (double), per frame num_vectors (int, 0 =
// no histogram at that frame).

// base slots from the accumulators: ColorHistogram::NormalizeToOne (histograms.cpp:340-360) and, per frame,
// VectorHistogram::NormalizeToOne (:584-596) on the float bin values.
__global__ void __launch_bounds__(256) finish_slots_kernel(const unsigned long long* __restrict__ hist_acc, const unsigned* __restrict__ hist_cnt,
                                                           const unsigned long long* __restrict__ flow_acc, const unsigned* __restrict__ flow_cnt,
                                                           int n_regions, int B, int frames, int frame_cap, int bins, size_t slot_words,
                                                           float* __restrict__ slots, double* __restrict__ weight_sum, int* __restrict__ num_vectors) {
  const int r = blockIdx.x;
  if (r >= n_regions) return;
  float* S = slots + (size_t)r * slot_words;
  const unsigned c = hist_cnt[r];
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    float v = 0.f;
    if (c) v = (float)((double)hist_acc[(size_t)r * B + b] / (67108864.0 * (double)c));
    S[b] = v;
  }
  if (threadIdx.x == 0) weight_sum[r] = (double)c;
  for (int t = threadIdx.x; t < frames; t += blockDim.x) {
    const size_t slot = (size_t)r * frame_cap + t;
    const unsigned nv = flow_acc ? flow_cnt[slot] : 0u;
    num_vectors[(size_t)r * frames + t] = (int)nv;
    float* F = S + B + (size_t)t * bins;
    if (!nv) { for (int k = 0; k < bins; ++k) F[k] = 0.f; continue; }
    float vals[kFlowBinsMax];
    float sum = 0;
    for (int k = 0; k < bins; ++k) { vals[k] = (float)((double)flow_acc[slot * bins + k] * (1.0 / 16777216.0)); sum += vals[k]; }
    if (sum > 0) {
      sum = (float)(1.0 / (double)sum);
      for (int k = 0; k < bins; ++k) vals[k] *= sum;
    }
    for (int k = 0; k < bins; ++k) F[k] = vals[k];
  }
}

// MergeDescriptorsFrom(a) then (b) into a fresh slot (segmentation_common.cpp:70-90): ColorHistogram::MergeWithHistogram
// on normalised histograms (histograms.cpp:262-338) and FlowDescriptor::MergeWithDescriptor (region_descriptor.cpp:512-553)
// -> VectorHistogram::MergeWithHistogram (histograms.cpp:518-531).  One CTA.
__global__ void __launch_bounds__(256) merge_slots_kernel(float* __restrict__ slots, double* __restrict__ weight_sum, int* __restrict__ num_vectors,
                                                          int a, int b, int dst, int B, int frames, int bins, size_t slot_words) {
  __shared__ double red[256];
  const float* A = slots + (size_t)a * slot_words;
  const float* Bs = slots + (size_t)b * slot_words;
  float* D = slots + (size_t)dst * slot_words;
  const double wa = weight_sum[a], wb = weight_sum[b];
  const double n = wa + wb;
  if (n == 0) {
    for (int k = threadIdx.x; k < B; k += blockDim.x) D[k] = A[k];
  } else {
    const float n_l = (float)(wa / n), n_r = (float)(wb / n);
    double part = 0;
    for (int k = threadIdx.x; k < B; k += blockDim.x) {
      const float v = A[k] * n_l + Bs[k] * n_r;
      D[k] = v;
      part += (double)v;
    }
    red[threadIdx.x] = part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    const float denom = (float)(1.0 / red[0]);
    for (int k = threadIdx.x; k < B; k += blockDim.x) D[k] *= denom;
  }
  if (threadIdx.x == 0) weight_sum[dst] = n;
  for (int t = threadIdx.x; t < frames; t += blockDim.x) {
    const int na = num_vectors[(size_t)a * frames + t], nb = num_vectors[(size_t)b * frames + t];
    const float* FA = A + B + (size_t)t * bins;
    const float* FB = Bs + B + (size_t)t * bins;
    float* FD = D + B + (size_t)t * bins;
    int nd = 0;
    if (na && nb) {
      const float n_l = (float)na, n_r = (float)nb;
      const float inv = 1.0f / (n_l + n_r);
      float vals[kFlowBinsMax];
      float sum = 0;
      for (int k = 0; k < bins; ++k) { vals[k] = (FA[k] * n_l + FB[k] * n_r) * inv; sum += vals[k]; }
      if (sum > 0) {
        sum = (float)(1.0 / (double)sum);
        for (int k = 0; k < bins; ++k) vals[k] *= sum;
      }
      for (int k = 0; k < bins; ++k) FD[k] = vals[k];
      nd = na + nb;
    } else if (na) {
      for (int k = 0; k < bins; ++k) FD[k] = FA[k];
      nd = na;
    } else if (nb) {
      for (int k = 0; k < bins; ++k) FD[k] = FB[k];
      nd = nb;
    } else {
      for (int k = 0; k < bins; ++k) FD[k] = 0.f;
    }
    num_vectors[(size_t)dst * frames + t] = nd;
  }
}

struct DistJob { int slot_a, slot_b, size_a, size_b; };

// RegionInformation::DescriptorDistances + SquaredORDistance[SizePenalized]::Evaluate (region_descriptor.h:195-230) for a
// list of region pairs, one CTA per pair: ColorHistogram::ChiSquareDist (histograms.cpp:391-407) over all four warps,
// FlowDescriptor::RegionDistance (region_descriptor.cpp:465-498) on warp 0, RegionSizePenalizer::RegionDistance
// (:377-383).  `jobs` and `out` are mapped pinned host memory: a greedy merge costs one launch and one stream
// synchronisation, no copies (a batch is ~170 pairs: 3 KB in, 700 B out over PCIe from inside the kernel).
constexpr int kDistThreads = 128;
__global__ void __launch_bounds__(kDistThreads) slot_distance_kernel(const float* __restrict__ slots, const int* __restrict__ num_vectors,
                                                                     const DistJob* __restrict__ jobs, int n_jobs, int B, int frames, int bins,
                                                                     size_t slot_words, int use_appearance, int use_flow, int use_size,
                                                                     float penalizer, float inv_av_region_size, float* __restrict__ out) {
  __shared__ DistJob job;
  __shared__ double part[kDistThreads / 32];
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) job = jobs[blockIdx.x];
  __syncthreads();
  const DistJob j = job;
  const float* A = slots + (size_t)j.slot_a * slot_words;
  const float* Bs = slots + (size_t)j.slot_b * slot_words;
  float result = 1.0f;
  if (use_appearance) {
    double sum = 0.0;
    for (int k = (int)threadIdx.x; k < B; k += kDistThreads) {
      const float a = __ldg(&A[k]), c = __ldg(&Bs[k]);
      const float add = a + c;
      if (fabs((double)add) > 1e-12) { const float sub = a - c; sum += (double)(sub * sub / add); }
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) part[warp] = sum;
    __syncthreads();
    double total = 0.0;
    for (int k = 0; k < kDistThreads / 32; ++k) total += part[k];
    result *= (1.0f - (float)(0.5 * total));
  }
  if (warp != 0) return;
  if (use_flow) {
    double sum = 0, sum_w = 0;
    for (int t = (int)lane; t < frames; t += 32) {
      const int na = num_vectors[(size_t)j.slot_a * frames + t], nb = num_vectors[(size_t)j.slot_b * frames + t];
      if (!na || !nb) continue;
      const float* FA = A + B + (size_t)t * bins;
      const float* FB = Bs + B + (size_t)t * bins;
      float s = 0;
      for (int k = 0; k < bins; ++k) {
        const float add = FA[k] + FB[k];
        if (add) { const float sub = FA[k] - FB[k]; s += sub * sub / add; }
      }
      const float chi = (float)(0.5 * (double)s);
      const float weight = (float)min(na, nb);
      sum += (double)(chi * weight);
      sum_w += (double)weight;
    }
    for (int o = 16; o > 0; o >>= 1) { sum += __shfl_xor_sync(0xffffffffu, sum, o); sum_w += __shfl_xor_sync(0xffffffffu, sum_w, o); }
    const float fd = sum_w > 0 ? (float)(sum / sum_w) : 0.f;
    result *= (1.0f - fd);
  }
  if (lane == 0) {
    result = 1.0f - result;
    float d = result * result;
    if (use_size) {
      const int min_sz = min(j.size_a, j.size_b);
      const float size_scale = (float)(1.0 + (double)penalizer * log((double)((float)min_sz * inv_av_region_size)) / log(2.0));
      d = fmaxf(0.f, fminf(1.f, d * fminf(1.0f, size_scale)));
    }
    out[blockIdx.x] = d;
  }
}

int grid_of(size_t items, int block) {
  const size_t want = (items + block - 1) / block;
  return (int)std::max<size_t>(1, std::min<size_t>(want, 148 * 8));
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
using vsbr::Raster;
using vsbr::Raster3D;
using vsbr::Slice;

template <class T> bool insert_sorted_unique(const T& t, std::vector<T>* v) {
  auto pos = std::lower_bound(v->begin(), v->end(), t);
  if (pos == v->end() || *pos != t) { v->insert(pos, t); return true; }
  return false;
}

// One node of a hierarchy level (RegionInformation, segmentation_common.h:39-116).
struct Node {
  int index = -1, size = 0, parent_idx = -1;
  bool removed = false;
  std::vector<int> neighbors;                 // indices into the node's own level, sorted
  std::shared_ptr<Raster3D> raster;           // base level and the merged first level
  std::vector<int> children;
  bool has_children = false;
  Node* counterpart = nullptr;
  int constrained_id = -1, region_id = -1;
  std::vector<int> counterpart_ids;           // region ids of the counterpart's ancestors, level 1 upwards
  bool has_counterpart_ids = false;
  int slot = -1;                              // descriptor slot on the device
};
typedef std::vector<std::unique_ptr<Node>> Level;

struct FrameOutR {
  int width, height, chunk_id, chunk_size, overlap_start, hierarchy_frame_idx;
  std::vector<int32_t> flat;                  // record in the layout of vsb200_region_pop
};

struct Options {
  int min_region_num = 10, max_region_num = 10000;
  float level_cutoff_fraction = 0.8f, small_region_penalizer = 0.25f;
  int luminance_bins = 10, color_bins = 20, flow_bins = 16;
  int chunk_set_size = 6, chunk_set_overlap = 2, constraint_chunks = 1;
  bool use_appearance = true, use_flow = true, use_size_penalizer = true;
  int num_buckets = 2048;
};

struct Device {                                // per handle
  cudaStream_t stream = nullptr;
  uint8_t* d_bgr = nullptr; uint8_t* h_bgr = nullptr;
  float* d_flow = nullptr; float* h_flow = nullptr;
  int* d_ids = nullptr;
  PaintRun* d_runs = nullptr; PaintRun* h_runs = nullptr; size_t runs_cap = 0;
  DistJob* d_jobs = nullptr; DistJob* h_jobs = nullptr; float* d_dist = nullptr; float* h_dist = nullptr; size_t jobs_cap = 0;
  double kernel_ms = 0;
  long long launches = 0;
  // development taps (VSB200_STAGE_DEBUG): wall-clock ms since the last chunk-set boundary
  double t_add_frame = 0, t_evaluate = 0, t_hierarchy = 0, t_retrieve = 0, t_push = 0, t_begin = 0;
  long long n_evaluate = 0, n_jobs = 0, n_merge_launches = 0;
};
inline double wall_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// The regions of one chunk set and everything computed over them (the hierarchical half of Segmentation).
class ChunkSet {
 public:
  ChunkSet(const Options& o, int w, int h, int chunk_id, Device* dev) : opt_(o), w_(w), h_(h), chunk_id_(chunk_id), dev_(dev) {
    B_ = o.luminance_bins * o.color_bins * o.color_bins;
  }
  ~ChunkSet() { release(); }
  void release() {
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(hist_acc_); F(flow_acc_); F(slots_); F(weight_sum_); F(num_vectors_);
    hist_acc_ = nullptr; flow_acc_ = nullptr; slots_ = nullptr; weight_sum_ = nullptr; num_vectors_ = nullptr;
  }
  int frames() const { return frame_number_; }
  int levels() const { return (int)levels_.size(); }

  // InitializeBaseHierarchyLevel (segmentation.cpp:80-198)
  int begin_chunk(const vsb200_frame_result& r, std::unordered_map<int, Node*>* input_mapping, std::unordered_map<int, Node*>* output_mapping) {
    if (levels_.size() != 1) { levels_.clear(); levels_.emplace_back(); }
    if (output_mapping) output_mapping->clear();
    Level& base = levels_[0];
    for (int c = 0; c < r.n_compound; ++c) {
      const int id = r.compound[4 * c], size = r.compound[4 * c + 1];
      auto it = by_id_.find(id);
      Node* n;
      if (it == by_id_.end()) {
        base.emplace_back(new Node);
        n = base.back().get();
        n->index = (int)base.size() - 1;
        n->size = size;
        n->raster.reset(new Raster3D);
        n->slot = n->index;
        if (input_mapping) { auto cp = input_mapping->find(id); if (cp != input_mapping->end()) n->counterpart = cp->second; }
        by_id_[id] = n;
      } else {
        n = it->second;
        n->size += size;
      }
      if (output_mapping) (*output_mapping)[id] = n;
    }
    for (int c = 0; c < r.n_compound; ++c) {
      Node* n = by_id_[r.compound[4 * c]];
      for (int q = r.neighbor_offset[c]; q < r.neighbor_offset[c + 1]; ++q) {
        auto nb = by_id_.find(r.neighbor_id[q]);
        if (nb == by_id_.end()) { set_error("region stage: neighbour %d of region %d is not part of the chunk", r.neighbor_id[q], r.compound[4 * c]); return VSB200_ERR_INVALID; }
        insert_sorted_unique(nb->second->index, &n->neighbors);
      }
    }
    return grow_accumulators((int)base.size());
  }

  // AddOverSegmentation (segmentation.cpp:200-239): rasters on the host, descriptors on the device.  d_ids is scratch;
  // the frame (and flow) are already resident.
  int add_frame(const vsb200_frame_result& r, bool flow_valid) {
    const double t_in = wall_ms();
    Level& base = levels_[0];
    size_t n_runs = 0;
    for (int k = 0; k < r.n_regions; ++k) n_runs += (size_t)(r.interval_offset[k + 1] - r.interval_offset[k]);
    if (n_runs > dev_->runs_cap) {
      if (dev_->d_runs) cudaFree(dev_->d_runs);
      if (dev_->h_runs) cudaFreeHost(dev_->h_runs);
      dev_->runs_cap = n_runs + n_runs / 2 + 1024;
      RS_CUDA(cudaMalloc(&dev_->d_runs, dev_->runs_cap * sizeof(PaintRun)));
      RS_CUDA(cudaMallocHost(&dev_->h_runs, dev_->runs_cap * sizeof(PaintRun)));
    }
    size_t w = 0;
    for (int k = 0; k < r.n_regions; ++k) {
      auto it = by_id_.find(r.region_id[k]);
      if (it == by_id_.end()) { set_error("region stage: region %d has no entry in the chunk's hierarchy", r.region_id[k]); return VSB200_ERR_INVALID; }
      Node* n = it->second;
      auto raster = std::make_shared<Raster>();
      for (int q = r.interval_offset[k]; q < r.interval_offset[k + 1]; ++q) {
        const int y = r.intervals[3 * q], lx = r.intervals[3 * q + 1], rx = r.intervals[3 * q + 2];
        raster->push_back(vsbr::Interval{y, lx, rx});
        dev_->h_runs[w++] = PaintRun{y, lx, rx, n->index};
      }
      n->raster->push_back(Slice{frame_number_, raster});
    }
    if (frame_number_ >= frame_cap_) RS_RC(grow_frames(frame_number_ + 1));
    cudaStream_t s = dev_->stream;
    const size_t npx = (size_t)w_ * h_;
    RS_CUDA(cudaMemsetAsync(dev_->d_ids, 0xff, npx * sizeof(int), s));
    RS_CUDA(cudaMemcpyAsync(dev_->d_runs, dev_->h_runs, n_runs * sizeof(PaintRun), cudaMemcpyHostToDevice, s));
    if (n_runs) paint_runs_kernel<<<(unsigned)((n_runs * 32 + 255) / 256), 256, 0, s>>>(dev_->d_runs, (int)n_runs, w_, dev_->d_ids);
    const int R = (int)base.size();
    if (opt_.use_appearance)
      RS_RC(launch_region_hist(dev_->d_bgr, w_ * 3, dev_->d_ids, w_, h_, R, opt_.luminance_bins, opt_.color_bins, hist_acc_, hist_cnt(), s));
    if (opt_.use_flow && flow_valid)
      flow_hist_kernel<<<grid_of(npx, 256), 256, 0, s>>>(dev_->d_flow, dev_->d_ids, npx, R, frame_number_, frame_cap_, opt_.flow_bins, flow_acc_, flow_cnt());
    RS_CUDA(cudaGetLastError());
    RS_CUDA(cudaStreamSynchronize(s));          // h_runs / the frame staging buffers are reused by the next frame
    dev_->launches += 3;
    ++frame_number_;
    dev_->t_add_frame += wall_ms() - t_in;
    return 0;
  }

  // PullCounterpartSegmentationResult (segmentation.cpp:241-270)
  void pull_counterparts(const ChunkSet& prev) {
    const int L = prev.levels();
    for (auto& n : levels_[0]) {
      if (!n->counterpart) continue;
      n->constrained_id = n->counterpart->region_id;
      n->counterpart_ids.assign(L - 1, 0);
      int cur = n->counterpart->parent_idx;
      for (int l = 1; l < L; ++l) { n->counterpart_ids[l - 1] = prev.levels_[l][cur]->region_id; cur = prev.levels_[l][cur]->parent_idx; }
      n->has_counterpart_ids = true;
    }
    constrained_ = true;
  }

  int run_hierarchy();                        // RunHierarchicalSegmentation(distance, enforce_max_region_num = true)
  void constrain_to_interval(int lhs, int rhs);
  void adjust_area(int lhs, int rhs);
  void assign_ids(bool use_constrained, const std::vector<int>& offsets, std::vector<int>* max_ids);
  void discard_bottom();
  void retrieve(int frame, bool with_hierarchy, FrameOutR* out) const;

 private:
  unsigned* hist_cnt() const { return reinterpret_cast<unsigned*>(hist_acc_ + (size_t)region_cap_ * B_); }
  unsigned* flow_cnt() const { return reinterpret_cast<unsigned*>(flow_acc_ + (size_t)region_cap_ * frame_cap_ * opt_.flow_bins); }

  int realloc_accumulators(int region_cap, int frame_cap) {
    cudaStream_t s = dev_->stream;
    unsigned long long* nh = nullptr;
    unsigned long long* nf = nullptr;
    const size_t hist_bytes = (size_t)region_cap * B_ * 8 + (size_t)region_cap * 4;
    const size_t flow_bytes = (size_t)region_cap * frame_cap * opt_.flow_bins * 8 + (size_t)region_cap * frame_cap * 4;
    RS_CUDA(cudaMalloc(&nh, hist_bytes));
    RS_CUDA(cudaMemsetAsync(nh, 0, hist_bytes, s));
    if (opt_.use_flow) { RS_CUDA(cudaMalloc(&nf, flow_bytes)); RS_CUDA(cudaMemsetAsync(nf, 0, flow_bytes, s)); }
    if (hist_acc_) {
      RS_CUDA(cudaMemcpyAsync(nh, hist_acc_, (size_t)region_cap_ * B_ * 8, cudaMemcpyDeviceToDevice, s));
      RS_CUDA(cudaMemcpyAsync(nh + (size_t)region_cap * B_, hist_cnt(), (size_t)region_cap_ * 4, cudaMemcpyDeviceToDevice, s));
    }
    if (flow_acc_ && nf) {
      // rows are [region][frame]: re-pitch when the frame capacity changes
      RS_CUDA(cudaMemcpy2DAsync(nf, (size_t)frame_cap * opt_.flow_bins * 8, flow_acc_, (size_t)frame_cap_ * opt_.flow_bins * 8,
                                (size_t)frame_cap_ * opt_.flow_bins * 8, region_cap_, cudaMemcpyDeviceToDevice, s));
      unsigned* ncnt = reinterpret_cast<unsigned*>(nf + (size_t)region_cap * frame_cap * opt_.flow_bins);
      RS_CUDA(cudaMemcpy2DAsync(ncnt, (size_t)frame_cap * 4, flow_cnt(), (size_t)frame_cap_ * 4, (size_t)frame_cap_ * 4, region_cap_, cudaMemcpyDeviceToDevice, s));
    }
    RS_CUDA(cudaStreamSynchronize(s));
    if (hist_acc_) cudaFree(hist_acc_);
    if (flow_acc_) cudaFree(flow_acc_);
    hist_acc_ = nh; flow_acc_ = nf; region_cap_ = region_cap; frame_cap_ = frame_cap;
    return 0;
  }
  int grow_accumulators(int regions) {
    if (regions <= region_cap_) return 0;
    return realloc_accumulators(std::max(regions + regions / 2, 1024), std::max(frame_cap_, 32));
  }
  int grow_frames(int frames) {
    if (frames <= frame_cap_) return 0;
    return realloc_accumulators(std::max(region_cap_, 1024), std::max(frames + frames / 2, 32));
  }

  int finish_descriptors();                   // PopulatingDescriptorsFinished for the base level -> descriptor slots
  int evaluate(const std::vector<DistJob>& jobs, float inv_av, std::vector<float>* out);
  int merge_slots(int a, int b, int dst);
  void setup_constraints(int level, std::vector<int>* ids, std::unordered_map<int, std::vector<int>>* skeleton) const;

  friend class Agglomeration;
  Options opt_;
  int w_, h_, chunk_id_, B_ = 0;
  Device* dev_;
  int frame_number_ = 0;
  bool constrained_ = false, constrained_ids_assigned_ = false;
  std::vector<Level> levels_;
  std::vector<Level> retired_;                // discarded bottom level: counterparts of the next set point into it
  std::unordered_map<int, Node*> by_id_;
  // device
  unsigned long long* hist_acc_ = nullptr; unsigned long long* flow_acc_ = nullptr;
  int region_cap_ = 0, frame_cap_ = 0;
  float* slots_ = nullptr; double* weight_sum_ = nullptr; int* num_vectors_ = nullptr;
  size_t slot_words_ = 0;
  int slot_cap_ = 0, next_slot_ = 0, set_frames_ = 0;
};

int ChunkSet::finish_descriptors() {
  const int R = (int)levels_[0].size();
  set_frames_ = std::max(1, frame_number_);
  slot_words_ = (size_t)B_ + (size_t)set_frames_ * opt_.flow_bins;
  slot_cap_ = 2 * R + 8;                       // every merge of every level writes one new slot: fewer than R in total
  next_slot_ = R;
  RS_CUDA(cudaMalloc(&slots_, (size_t)slot_cap_ * slot_words_ * sizeof(float)));
  RS_CUDA(cudaMalloc(&weight_sum_, (size_t)slot_cap_ * sizeof(double)));
  RS_CUDA(cudaMalloc(&num_vectors_, (size_t)slot_cap_ * set_frames_ * sizeof(int)));
  finish_slots_kernel<<<R, 256, 0, dev_->stream>>>(hist_acc_, hist_cnt(), opt_.use_flow ? flow_acc_ : nullptr, opt_.use_flow ? flow_cnt() : nullptr, R, B_,
                                                   set_frames_, frame_cap_, opt_.flow_bins, slot_words_, slots_, weight_sum_, num_vectors_);
  RS_CUDA(cudaGetLastError());
  ++dev_->launches;
  return 0;
}

int ChunkSet::evaluate(const std::vector<DistJob>& jobs, float inv_av, std::vector<float>* out) {
  out->resize(jobs.size());
  if (jobs.empty()) return 0;
  const double t_in = wall_ms();
  ++dev_->n_evaluate; dev_->n_jobs += (long long)jobs.size();
  if (jobs.size() > dev_->jobs_cap) {
    if (dev_->h_jobs) cudaFreeHost(dev_->h_jobs);
    if (dev_->h_dist) cudaFreeHost(dev_->h_dist);
    dev_->h_jobs = nullptr; dev_->h_dist = nullptr; dev_->jobs_cap = 0;
    const size_t cap = jobs.size() * 2 + 1024;
    RS_CUDA(cudaHostAlloc(&dev_->h_jobs, cap * sizeof(DistJob), cudaHostAllocMapped));
    RS_CUDA(cudaHostAlloc(&dev_->h_dist, cap * sizeof(float), cudaHostAllocMapped));
    RS_CUDA(cudaHostGetDevicePointer((void**)&dev_->d_jobs, dev_->h_jobs, 0));
    RS_CUDA(cudaHostGetDevicePointer((void**)&dev_->d_dist, dev_->h_dist, 0));
    dev_->jobs_cap = cap;
  }
  memcpy(dev_->h_jobs, jobs.data(), jobs.size() * sizeof(DistJob));
  cudaStream_t s = dev_->stream;
  slot_distance_kernel<<<(unsigned)jobs.size(), kDistThreads, 0, s>>>(
      slots_, num_vectors_, dev_->d_jobs, (int)jobs.size(), B_, set_frames_, opt_.flow_bins, slot_words_, opt_.use_appearance ? 1 : 0,
      opt_.use_flow ? 1 : 0, opt_.use_size_penalizer ? 1 : 0, opt_.small_region_penalizer, inv_av, dev_->d_dist);
  RS_CUDA(cudaGetLastError());
  RS_CUDA(cudaStreamSynchronize(s));
  memcpy(out->data(), dev_->h_dist, jobs.size() * sizeof(float));
  ++dev_->launches;
  dev_->t_evaluate += wall_ms() - t_in;
  return 0;
}

int ChunkSet::merge_slots(int a, int b, int dst) {
  merge_slots_kernel<<<1, 256, 0, dev_->stream>>>(slots_, weight_sum_, num_vectors_, a, b, dst, B_, set_frames_, opt_.flow_bins, slot_words_);
  RS_CUDA(cudaGetLastError());
  ++dev_->launches;
  ++dev_->n_merge_launches;
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Greedy agglomeration of one level (RegionAgglomerationGraph, region_segmentation_graph.cpp).  Buckets are
// append-only arrays with tombstones and a head cursor (the reference keeps std::lists and iterators): same
// first-in-first-out order inside a bucket.
// ---------------------------------------------------------------------------------------------
class Agglomeration {
 public:
  Agglomeration(ChunkSet* set, const Level& level, float inv_av) : set_(set), level_(level), inv_av_(inv_av) {
    num_buckets_ = set->opt_.num_buckets;
    max_weight_ = 1.0f * 1.01f;
    edge_scale_ = num_buckets_ * (1.0f / max_weight_);
    buckets_.resize(num_buckets_ + 1);
    heads_.assign(num_buckets_ + 1, 0);
  }

  struct Rep {                                 // RegionAgglomerationGraph::Region
    int id, constraint_id, sz;
    const Node* info;                          // the level's node, or the merged node below
    std::unique_ptr<Node> merged;
  };

  // AddRegionEdges[Constrained] (:45-71,257-311): level 0 evaluates every neighbour pair in ONE launch, later levels
  // reuse the quantised weights of the level below.
  int add_regions(const std::vector<int>& constraint_ids, const std::unordered_map<long long, float>* weights,
                  const std::unordered_map<int, std::vector<int>>* skeleton) {
    const int n = (int)level_.size();
    reps_.reserve(n);
    std::vector<std::pair<int, int>> order;    // undirected edges in the reference's insertion order
    for (int i = 0; i < n; ++i) {
      for (int nb : level_[i]->neighbors) {
        const long long key = ekey(i, nb);
        if (seen_.insert(std::make_pair(key, 0)).second) order.emplace_back(i, nb);
      }
    }
    seen_.clear();
    std::vector<float> w(order.size());
    if (weights) {
      for (size_t k = 0; k < order.size(); ++k) {
        auto it = weights->find(ekey(order[k].first, order[k].second));
        if (it == weights->end()) { set_error("region stage: edge weight of the previous level is missing"); return VSB200_ERR_INVALID; }
        w[k] = it->second;
      }
    } else {
      std::vector<DistJob> jobs(order.size());
      for (size_t k = 0; k < order.size(); ++k) {
        const Node& a = *level_[order[k].first];
        const Node& b = *level_[order[k].second];
        jobs[k] = DistJob{a.slot, b.slot, a.size, b.size};
      }
      RS_RC(set_->evaluate(jobs, inv_av_, &w));
    }
    for (int i = 0; i < n; ++i) reps_.push_back(Rep{i, constraint_ids[i], 1, level_[i].get(), nullptr});
    for (size_t k = 0; k < order.size(); ++k) add_edge(order[k].first, order[k].second, w[k]);
    if (skeleton)
      for (const auto& e : *skeleton) {
        int prev = e.second.front();
        for (size_t q = 1; q < e.second.size(); ++q) { add_edge(prev, e.second[q], max_weight_ * 2); prev = e.second[q]; }
      }
    return 0;
  }

  // SegmentGraph (:73-177)
  int segment(bool merge_rasters, float cutoff, int* merges_out) {
    merge_rasters_ = merge_rasters;
    int num_merges = (int)(reps_.size() * (1.0f - cutoff));
    const int constraint_merges = (int)(live_count(num_buckets_) * cutoff);
    num_merges -= constraint_merges;
    num_merges = std::min<int>(num_merges, (int)reps_.size() - 1);
    int lowest = 0;
    while (lowest < num_buckets_ && empty(lowest)) ++lowest;
    int actual = 0;
    for (int merge = 0; merge < num_merges; ++merge) {
      if (lowest >= num_buckets_) break;
      bool done = false;
      while (!done) {
        const Entry e = front(lowest);
        Rep* r1 = find(e.a);
        Rep* r2 = find(e.b);
        bool at_end;
        if (!mergable(r1->constraint_id, r2->constraint_id)) {
          // different constraints: dropped from the bucket, kept in the position map as "no position"
          position_[ekey(e.a, e.b)].index = -1;
          pop_front(lowest);
          at_end = empty(lowest);
        } else {
          float min_dist = 0;
          RS_RC(merge_reps(r1, r2, &min_dist));
          const int min_bucket = (int)(min_dist * edge_scale_);
          ++actual;
          if (min_bucket < lowest) { lowest = min_bucket; break; }
          at_end = empty(lowest);
          done = true;
        }
        if (at_end) {
          do { ++lowest; } while (lowest < num_buckets_ && empty(lowest));
          if (lowest >= num_buckets_) break;
        }
      }
    }
    // forced merges along the skeleton (virtual edges, last bucket)
    for (size_t q = 0; q < buckets_[num_buckets_].size(); ++q) {
      const Entry e = buckets_[num_buckets_][q];
      Rep* r1 = find(e.a);
      Rep* r2 = find(e.b);
      if (r1 != r2) { float d; RS_RC(merge_reps(r1, r2, &d)); ++actual; }
    }
    *merges_out = actual;
    return 0;
  }

  // ObtainSegmentationResult (:181-255)
  void result(Level* prev_level, Level* next_level, std::unordered_map<long long, float>* weights) {
    std::unordered_map<int, Node*> assigned;
    std::vector<int> rep_of;
    int next_idx = 0;
    for (int child = 0; child < (int)prev_level->size(); ++child) {
      Rep* r = find(child);
      if (assigned.find(r->id) == assigned.end()) {
        if (r->info != r->merged.get()) {       // never merged: basic copy (descriptor slot shared, not copied)
          std::unique_ptr<Node> c(new Node);
          c->size = r->info->size;
          c->neighbors = r->info->neighbors;
          c->slot = r->info->slot;
          if (merge_rasters_) c->raster.reset(new Raster3D(*r->info->raster));
          r->merged.swap(c);
          r->info = r->merged.get();
        }
        r->merged->index = next_idx++;
        r->merged->constrained_id = r->constraint_id;
        r->merged->has_children = true;
        assigned[r->id] = r->merged.get();
        next_level->push_back(std::move(r->merged));
        rep_of.push_back(r->id);
      }
      Node* res = assigned[r->id];
      res->children.push_back(child);
      (*prev_level)[child]->parent_idx = res->index;
    }
    weights->clear();
    const float inv_scale = 1.0f / edge_scale_;
    for (auto& node : *next_level) {
      std::vector<int> mapped;
      for (int nb : node->neighbors) {
        const Rep* nr = find(nb);
        const int nidx = nr->info->index;
        (*weights)[ekey(node->index, nidx)] = inv_scale * position_[ekey(rep_of[node->index], nr->id)].bucket;
        insert_sorted_unique(nidx, &mapped);
      }
      node->neighbors.swap(mapped);
    }
  }

 private:
  struct Entry { int a, b; bool dead; };
  struct Position { int bucket = -1, index = -1; };      // index -1: not in a bucket (unmergeable)
  static long long ekey(int a, int b) { return a < b ? ((long long)a << 32) | (unsigned)b : ((long long)b << 32) | (unsigned)a; }
  static bool mergable(int c1, int c2) { return c1 < 0 || c2 < 0 || c1 == c2; }

  bool empty(int b) {
    auto& v = buckets_[b];
    size_t& h = heads_[b];
    while (h < v.size() && v[h].dead) ++h;
    return h >= v.size();
  }
  size_t live_count(int b) const { size_t c = 0; for (const auto& e : buckets_[b]) c += e.dead ? 0 : 1; return c; }
  Entry front(int b) { empty(b); return buckets_[b][heads_[b]]; }
  void pop_front(int b) { buckets_[b][heads_[b]].dead = true; }

  bool add_edge(int r1, int r2, float weight) {           // AddEdge (:320-349)
    const int bucket = std::min(num_buckets_, (int)(weight * edge_scale_));
    const bool ok = mergable(reps_[r1].constraint_id, reps_[r2].constraint_id);
    int index = -1;
    if (ok) { index = (int)buckets_[bucket].size(); buckets_[bucket].push_back(Entry{std::min(r1, r2), std::max(r1, r2), false}); }
    if (bucket != num_buckets_) { Position p; p.bucket = bucket; p.index = index; position_[ekey(r1, r2)] = p; }
    return ok;
  }

  Rep* find(int id) {                                      // GetRegion (:351-369)
    int root = id;
    while (reps_[root].id != root) root = reps_[root].id;
    while (reps_[id].id != root) { const int nx = reps_[id].id; reps_[id].id = root; id = nx; }
    return &reps_[root];
  }

  void remove_edges(int region, const std::vector<int>& neighbors, int other, std::vector<int>* removed) {   // :371-405
    for (int n : neighbors) {
      const int nb = find(n)->id;
      auto it = position_.find(ekey(region, nb));
      if (it == position_.end()) continue;
      if (it->second.index >= 0) buckets_[it->second.bucket][it->second.index].dead = true;
      position_.erase(it);
      if (nb != other) insert_sorted_unique(nb, removed);
    }
  }

  int merge_reps(Rep* r1, Rep* r2, float* min_dist_out) {  // MergeRegions (:409-503)
    const Node& i1 = *r1->info;
    const Node& i2 = *r2->info;
    const int id1 = r1->id, id2 = r2->id;
    std::vector<int> nbs;
    remove_edges(id1, i1.neighbors, id2, &nbs);
    remove_edges(id2, i2.neighbors, id1, &nbs);
    Rep* m = r1->sz > r2->sz ? r1 : r2;
    m->sz = r1->sz + r2->sz;
    r1->id = m->id;
    r2->id = m->id;
    m->constraint_id = std::max(r1->constraint_id, r2->constraint_id);
    std::unique_ptr<Node> node(new Node);
    node->size = i1.size + i2.size;
    node->neighbors.swap(nbs);
    node->slot = set_->next_slot_++;
    if (node->slot >= set_->slot_cap_) { set_error("region stage: descriptor slots exhausted"); return VSB200_ERR_CAPACITY; }
    RS_RC(set_->merge_slots(i1.slot, i2.slot, node->slot));
    if (merge_rasters_) {
      node->raster.reset(new Raster3D);
      merge_raster3d(*i1.raster, *i2.raster, node->raster.get());
    }
    std::vector<DistJob> jobs;
    jobs.reserve(node->neighbors.size());
    for (int nb : node->neighbors) { const Node* ni = reps_[nb].info; jobs.push_back(DistJob{node->slot, ni->slot, node->size, ni->size}); }
    std::vector<float> d;
    RS_RC(set_->evaluate(jobs, inv_av_, &d));
    float min_dist = 1.e6f;
    for (size_t k = 0; k < jobs.size(); ++k)
      if (add_edge(m->id, node->neighbors[k], d[k])) min_dist = std::min(min_dist, d[k]);
    m->merged.swap(node);
    m->info = m->merged.get();
    *min_dist_out = min_dist;
    return 0;
  }

  static void merge_raster3d(const Raster3D& a, const Raster3D& b, Raster3D* out) {   // MergeRasterization3D (segmentation_util.cpp:607-642)
    size_t i = 0, j = 0;
    while (i < a.size() || j < b.size()) {
      const int fa = i < a.size() ? a[i].frame : std::numeric_limits<int>::max();
      const int fb = j < b.size() ? b[j].frame : std::numeric_limits<int>::max();
      if (fa < fb) { out->push_back(Slice{fa, std::make_shared<Raster>(*a[i].raster)}); ++i; }
      else if (fb < fa) { out->push_back(Slice{fb, std::make_shared<Raster>(*b[j].raster)}); ++j; }
      else {
        auto m = std::make_shared<Raster>();
        vsbr::merge_rasters(*a[i].raster, *b[j].raster, m.get());
        out->push_back(Slice{fa, m});
        ++i; ++j;
      }
    }
  }

  ChunkSet* set_;
  const Level& level_;
  float inv_av_;
  int num_buckets_ = 0;
  float max_weight_ = 1.0f, edge_scale_ = 1.0f;
  bool merge_rasters_ = false;
  std::vector<std::vector<Entry>> buckets_;
  std::vector<size_t> heads_;
  std::unordered_map<long long, Position> position_;
  std::unordered_map<long long, int> seen_;
  std::vector<Rep> reps_;
};

void ChunkSet::setup_constraints(int level, std::vector<int>* ids, std::unordered_map<int, std::vector<int>>* skeleton) const {   // segmentation.cpp:600-669
  ids->clear();
  for (const auto& n : levels_[level]) {
    int child = n->index;
    if (level > 0) {
      for (int l = level; l > 0; --l) {
        bool found = false;
        for (int c : levels_[l][child]->children)
          if (levels_[l - 1][c]->constrained_id >= 0) { child = c; found = true; break; }
        if (!found) { child = -1; break; }
      }
    } else if (n->constrained_id < 0) {
      child = -1;
    }
    int id = -1;
    if (child >= 0) {
      const Node& base = *levels_[0][child];
      if (base.has_counterpart_ids && level < (int)base.counterpart_ids.size()) id = base.counterpart_ids[level];
    }
    ids->push_back(id);
    if (id >= 0) (*skeleton)[id].push_back(n->index);
  }
}

int ChunkSet::run_hierarchy() {                             // segmentation.cpp:305-389
  RS_RC(finish_descriptors());
  int level = 0;
  int count = (int)levels_[0].size();
  std::unordered_map<long long, float> weights;
  while (count > opt_.min_region_num) {
    const Level& cur = levels_[level];
    float inv_av = 1.0f;                                    // RegionSizePenalizerUpdater::InitializeUpdate (region_descriptor.cpp:392-415)
    if (opt_.use_size_penalizer && !cur.empty()) {
      std::vector<int> sizes;
      sizes.reserve(cur.size());
      for (const auto& n : cur) sizes.push_back(n->size);
      auto median = sizes.begin() + sizes.size() / 2;
      std::nth_element(sizes.begin(), median, sizes.end());
      inv_av = *median > 0 ? 1.0f / *median : 1.f;
    }
    Agglomeration graph(this, cur, inv_av);
    std::vector<int> constraint_ids(cur.size(), -1);
    std::unordered_map<int, std::vector<int>> skeleton;
    if (constrained_) setup_constraints(level, &constraint_ids, &skeleton);
    RS_RC(graph.add_regions(constraint_ids, level == 0 ? nullptr : &weights, constrained_ ? &skeleton : nullptr));
    int merges = 0;
    if (level == 0) {
      const float cutoff = std::min(1.0f, opt_.max_region_num * (1.0f / levels_[0].size()));
      RS_RC(graph.segment(true, cutoff, &merges));
    } else {
      RS_RC(graph.segment(false, opt_.level_cutoff_fraction, &merges));
      if (!merges) break;
    }
    levels_.emplace_back();
    graph.result(&levels_[level], &levels_.back(), &weights);
    count = (int)levels_[level].size();
    ++level;
  }
  return 0;
}

void ChunkSet::constrain_to_interval(int lhs, int rhs) {    // segmentation.cpp:392-422
  for (auto& n : levels_[0])
    if (!n->raster || n->raster->empty() || n->raster->front().frame >= rhs || n->raster->back().frame < lhs) n->removed = true;
  for (size_t l = 1; l < levels_.size(); ++l)
    for (auto& n : levels_[l]) {
      bool removed = true;
      for (int c : n->children) if (!levels_[l - 1][c]->removed) { removed = false; break; }
      n->removed = removed;
    }
}

void ChunkSet::adjust_area(int lhs, int rhs) {              // segmentation.cpp:424-456
  std::unordered_map<int, int> prev;
  for (auto& n : levels_[0]) {
    int inc = 0;
    if (!n->raster) continue;
    for (const auto& s : *n->raster) if (s.frame < lhs || s.frame >= rhs) inc -= vsbr::raster_area(*s.raster);
    n->size += inc;
    prev[n->index] = inc;
  }
  for (size_t l = 1; l < levels_.size(); ++l) {
    std::unordered_map<int, int> cur;
    for (auto& n : levels_[l]) {
      int inc = 0;
      for (int c : n->children) inc += prev[c];
      n->size += inc;
      cur[n->index] = inc;
    }
    prev.swap(cur);
  }
}

void ChunkSet::assign_ids(bool use_constrained, const std::vector<int>& offsets, std::vector<int>* max_ids) {   // segmentation.cpp:549-582
  constrained_ids_assigned_ = use_constrained;
  std::vector<int> local = offsets;
  if (local.size() < levels_.size()) local.resize(levels_.size());
  for (size_t l = 0; l < levels_.size(); ++l) {
    int max_id = -1;
    for (auto& n : levels_[l]) {
      n->region_id = (use_constrained && n->constrained_id >= 0) ? n->constrained_id : n->index + local[l];
      max_id = std::max(max_id, n->region_id);
    }
    if (max_ids) (*max_ids)[l] = std::max(offsets[l], max_id + 1);
  }
}

void ChunkSet::discard_bottom() {                           // segmentation.cpp:584-598
  if (levels_.size() < 2) return;
  for (auto& n : levels_[1]) { n->children.clear(); n->has_children = false; }
  retired_.push_back(std::move(levels_[0]));
  levels_.erase(levels_.begin());
}

// RetrieveSegmentation3D (segmentation.cpp:458-533) into the flat record of vsb200_region_pop.
void ChunkSet::retrieve(int frame, bool with_hierarchy, FrameOutR* out) const {
  struct Item { int id; const Raster* raster; };
  std::vector<Item> items;
  for (const auto& n : levels_[0]) {
    if (!n->raster) continue;
    auto it = std::lower_bound(n->raster->begin(), n->raster->end(), frame, [](const Slice& s, int f) { return s.frame < f; });
    if (it == n->raster->end() || it->frame != frame) continue;
    items.push_back(Item{n->region_id, it->raster.get()});
  }
  if (constrained_ids_assigned_) std::sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.id < b.id; });
  std::vector<int32_t>& f = out->flat;
  f.clear();
  auto bits = [](float v) { int32_t b; memcpy(&b, &v, 4); return b; };
  const int n_levels = with_hierarchy ? (int)levels_.size() : 0;
  const int32_t head[8] = {w_, h_, chunk_id_, out->chunk_size, out->overlap_start, out->hierarchy_frame_idx, (int32_t)items.size(), n_levels};
  f.insert(f.end(), head, head + 8);
  for (const Item& it : items) {
    f.push_back(it.id);
    f.push_back((int32_t)it.raster->size());
    for (const auto& s : *it.raster) { f.push_back(s.y); f.push_back(s.lx); f.push_back(s.rx); }
    const vsbs::Moments m = vsbr::moments_of(*it.raster);
    for (float v : {m.size, m.mean_x, m.mean_y, m.xx, m.xy, m.yy}) f.push_back(bits(v));
  }
  if (!with_hierarchy) return;
  std::unordered_map<int, std::pair<int, int>> prev_bound, cur_bound;
  struct Comp { int id, size, parent, sf, ef; std::vector<int> nb, ch; };
  for (int l = 0; l < n_levels; ++l) {
    const Level& list = levels_[l];
    std::vector<Comp> comps;
    cur_bound.clear();
    for (const auto& np : list) {                           // AddCompoundRegionToSegmentationDesc (:702-773)
      const Node& n = *np;
      if (n.removed) continue;
      Comp c;
      c.id = n.region_id; c.size = n.size; c.parent = -1;
      for (int nb : n.neighbors) if (!list[nb]->removed) c.nb.push_back(list[nb]->region_id);
      if (constrained_ids_assigned_) std::sort(c.nb.begin(), c.nb.end());
      if (l < n_levels - 1) c.parent = levels_[l + 1][n.parent_idx]->region_id;
      int mn = std::numeric_limits<int>::max(), mx = 0;
      if (l > 0) {
        for (int ch : n.children) {
          if (levels_[l - 1][ch]->removed) continue;
          c.ch.push_back(levels_[l - 1][ch]->region_id);
          const auto b = prev_bound.find(ch);
          mn = std::min(mn, b->second.first);
          mx = std::max(mx, b->second.second);
        }
        if (constrained_ids_assigned_) std::sort(c.ch.begin(), c.ch.end());
      } else {
        mn = n.raster->front().frame;
        mx = n.raster->back().frame;
      }
      c.sf = mn; c.ef = mx;
      cur_bound[n.index] = std::make_pair(mn, mx);
      comps.push_back(std::move(c));
    }
    prev_bound.swap(cur_bound);
    if (constrained_ids_assigned_) std::stable_sort(comps.begin(), comps.end(), [](const Comp& a, const Comp& b) { return a.id < b.id; });
    f.push_back((int32_t)comps.size());
    for (const Comp& c : comps) {
      f.push_back(c.id); f.push_back(c.size); f.push_back(c.parent); f.push_back(c.sf); f.push_back(c.ef);
      f.push_back((int32_t)c.nb.size()); f.push_back((int32_t)c.ch.size());
      f.insert(f.end(), c.nb.begin(), c.nb.end());
      f.insert(f.end(), c.ch.begin(), c.ch.end());
    }
  }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// handle: RegionSegmentation (region_segmentation.cpp:97-365)
// ---------------------------------------------------------------------------------------------
struct vsb200_region {
  Options opt;
  int w = 0, h = 0, device = 0;
  bool use_flow = false;
  Device dev;
  std::unique_ptr<ChunkSet> seg, new_seg;
  int read_chunks = 0, chunk_sets = 0, overlap_start = -1, lookahead_start = -1, num_output_frames = 0;
  std::vector<int> max_region_ids;
  std::deque<std::unique_ptr<FrameOutR>> ready;
  std::unique_ptr<FrameOutR> last;
  bool flushed = false;

  ~vsb200_region() {
    seg.reset(); new_seg.reset();
    auto F = [](void* p) { if (p) cudaFree(p); };
    auto FH = [](void* p) { if (p) cudaFreeHost(p); };
    F(dev.d_bgr); F(dev.d_flow); F(dev.d_ids); F(dev.d_runs);      // d_jobs / d_dist alias the mapped host buffers
    FH(dev.h_bgr); FH(dev.h_flow); FH(dev.h_runs); FH(dev.h_jobs); FH(dev.h_dist);
    if (dev.stream) cudaStreamDestroy(dev.stream);
  }

  int segment_and_output(int overlap_start_, int lookahead_start_) {   // SegmentAndOutputChunk (:313-365)
    const double t_h0 = wall_ms();
    RS_RC(seg->run_hierarchy());
    dev.t_hierarchy += wall_ms() - t_h0;
    const double t_r0 = wall_ms();
    const int computed = seg->levels();
    if (computed > (int)max_region_ids.size()) max_region_ids.resize(computed, 0);
    seg->constrain_to_interval(0, lookahead_start_);
    seg->adjust_area(0, overlap_start_);
    std::vector<int> new_max(max_region_ids.size());
    seg->assign_ids(chunk_sets > 0, max_region_ids, &new_max);
    max_region_ids.swap(new_max);
    if (new_seg) new_seg->pull_counterparts(*seg);
    seg->discard_bottom();
    const int hierarchy_frame_idx = num_output_frames;
    for (int f = 0; f < overlap_start_; ++f) {
      std::unique_ptr<FrameOutR> out(new FrameOutR);
      out->hierarchy_frame_idx = hierarchy_frame_idx;
      out->chunk_size = lookahead_start_;
      out->overlap_start = overlap_start_;
      seg->retrieve(f, f == 0, out.get());
      ready.push_back(std::move(out));
      ++num_output_frames;
    }
    ++chunk_sets;
    dev.t_retrieve += wall_ms() - t_r0;
    if (getenv("VSB200_STAGE_DEBUG"))
      fprintf(stderr, "[vsb200 region] chunk set %d: %d frames, %d levels; add_frame %.1f ms, hierarchy %.1f ms (distance batches %lld with %lld pairs: %.1f ms; slot merges %lld), ids + records %.1f ms; begin_chunk %.1f ms; all pushes so far %.1f ms\n",
              chunk_sets - 1, seg->frames(), seg->levels(), dev.t_add_frame, dev.t_hierarchy, dev.n_evaluate, dev.n_jobs, dev.t_evaluate, dev.n_merge_launches, dev.t_retrieve, dev.t_begin, dev.t_push + (wall_ms() - t_h0));
    dev.t_add_frame = dev.t_evaluate = dev.t_hierarchy = dev.t_retrieve = dev.t_begin = 0;
    dev.n_evaluate = dev.n_jobs = dev.n_merge_launches = 0;
    return 0;
  }

  int boundary(bool flush) {                               // ChunkBoundaryOutput (:292-311)
    if (!flush) {
      const int look_ahead = lookahead_start > 0 ? lookahead_start : seg->frames();
      RS_RC(segment_and_output(overlap_start, look_ahead));
    } else {
      RS_RC(segment_and_output(seg->frames(), seg->frames()));
    }
    overlap_start = -1;
    lookahead_start = -1;
    if (!flush) { seg.swap(new_seg); new_seg.reset(); }
    else seg.reset();
    return 0;
  }
};

extern "C" {

void vsb200_region_default_opts(vsb200_region_opts* o) {
  if (!o) return;
  o->min_region_num = 10; o->max_region_num = 10000; o->level_cutoff_fraction = 0.8f; o->small_region_penalizer = 0.25f;
  o->luminance_bins = 10; o->color_bins = 20; o->flow_bins = 16;
  o->chunk_set_size = 6; o->chunk_set_overlap = 2; o->constraint_chunks = 1;
  o->save_descriptors = 0; o->use_appearance = 1; o->use_flow = 1; o->use_size_penalizer = 1; o->compute_vectorization = 0;
  o->device = 0;
}

int vsb200_region_create(const vsb200_region_opts* o, int width, int height, vsb200_region** out) {
  if (!o || !out || width < 2 || height < 2) { set_error("region_create: bad arguments"); return VSB200_ERR_INVALID; }
  // the CHECKs of RegionSegmentation::RegionSegmentation (region_segmentation.cpp:47-95)
  if (o->chunk_set_size <= 1) { set_error("At least two chunks per chunk_set required."); return VSB200_ERR_INVALID; }
  if (o->chunk_set_overlap <= 0) { set_error("At least one chunk in overlap expected."); return VSB200_ERR_INVALID; }
  if (o->chunk_set_overlap >= o->chunk_set_size) { set_error("Overlap has to be strictly smaller than a chunk set."); return VSB200_ERR_INVALID; }
  if (o->constraint_chunks > o->chunk_set_overlap) { set_error("Constraints must be smaller or equal to overlap"); return VSB200_ERR_INVALID; }
  if (!o->use_appearance && !o->use_flow) { set_error("At least apperance or flow need to be set."); return VSB200_ERR_INVALID; }
  if (o->compute_vectorization || o->save_descriptors) { set_error("compute_vectorization / save_descriptors are not built"); return VSB200_ERR_UNSUPPORTED; }
  if (o->luminance_bins < 2 || o->color_bins < 2 || o->flow_bins < 1 || o->flow_bins > kFlowBinsMax) { set_error("region_create: bad histogram bins"); return VSB200_ERR_INVALID; }
  if (vsb200_device_count() <= 0) { set_error("no sm_100 CUDA device available: this path has no CPU fallback"); return VSB200_ERR_NO_DEVICE; }
  std::unique_ptr<vsb200_region> r(new vsb200_region);
  r->w = width; r->h = height; r->device = o->device; r->use_flow = o->use_flow != 0;
  Options& p = r->opt;
  p.min_region_num = o->min_region_num; p.max_region_num = o->max_region_num; p.level_cutoff_fraction = o->level_cutoff_fraction;
  p.small_region_penalizer = o->small_region_penalizer; p.luminance_bins = o->luminance_bins; p.color_bins = o->color_bins; p.flow_bins = o->flow_bins;
  p.chunk_set_size = o->chunk_set_size; p.chunk_set_overlap = o->chunk_set_overlap; p.constraint_chunks = o->constraint_chunks;
  p.use_appearance = o->use_appearance != 0; p.use_flow = o->use_flow != 0; p.use_size_penalizer = o->use_size_penalizer != 0;
  RS_CUDA(cudaSetDevice(o->device));
  RS_CUDA(cudaStreamCreateWithFlags(&r->dev.stream, cudaStreamNonBlocking));
  const size_t npx = (size_t)width * height;
  RS_CUDA(cudaMalloc(&r->dev.d_bgr, npx * 3));
  RS_CUDA(cudaMallocHost(&r->dev.h_bgr, npx * 3));
  RS_CUDA(cudaMalloc(&r->dev.d_ids, npx * sizeof(int)));
  if (r->use_flow) { RS_CUDA(cudaMalloc(&r->dev.d_flow, npx * 8)); RS_CUDA(cudaMallocHost(&r->dev.h_flow, npx * 8)); }
  *out = r.release();
  return VSB200_OK;
}

// RegionSegmentation::ProcessFrame(false, desc, features, results) (region_segmentation.cpp:97-205)
int vsb200_region_push(vsb200_region* r, const vsb200_frame_result* overseg, const uint8_t* bgr, int row_stride_bytes,
                       const float* flow_xy, int flow_row_stride_bytes, int* n_ready) {
  if (n_ready) *n_ready = 0;
  if (!r || !overseg || !bgr || row_stride_bytes < r->w * 3) { set_error("region_push: bad arguments"); return VSB200_ERR_INVALID; }
  if (r->flushed) { set_error("region_push after flush"); return VSB200_ERR_INVALID; }
  if (overseg->width != r->w || overseg->height != r->h) { set_error("region_push: over-segmentation of another frame size"); return VSB200_ERR_INVALID; }
  RS_CUDA(cudaSetDevice(r->device));
  const double t_push0 = wall_ms();
  struct PushTimer { Device& d; double t0; ~PushTimer() { d.t_push += wall_ms() - t0; } } push_timer{r->dev, t_push0};
  const size_t before = r->ready.size();
  const Options& o = r->opt;
  if (!r->seg) r->seg.reset(new ChunkSet(o, r->w, r->h, r->chunk_sets, &r->dev));
  // features to the device: the frame (Lab is computed inside the histogram kernel) and the flow field
  for (int y = 0; y < r->h; ++y) memcpy(r->dev.h_bgr + (size_t)y * r->w * 3, bgr + (size_t)y * row_stride_bytes, (size_t)r->w * 3);
  RS_CUDA(cudaMemcpyAsync(r->dev.d_bgr, r->dev.h_bgr, (size_t)r->w * r->h * 3, cudaMemcpyHostToDevice, r->dev.stream));
  const bool flow_valid = r->use_flow && flow_xy;
  if (flow_valid) {
    for (int y = 0; y < r->h; ++y) memcpy(r->dev.h_flow + (size_t)y * r->w * 2, (const char*)flow_xy + (size_t)y * flow_row_stride_bytes, (size_t)r->w * 8);
    RS_CUDA(cudaMemcpyAsync(r->dev.d_flow, r->dev.h_flow, (size_t)r->w * r->h * 8, cudaMemcpyHostToDevice, r->dev.stream));
  }
  const int overlap_start_chunk = o.chunk_set_size - o.chunk_set_overlap;
  const int lookahead_start_chunk = overlap_start_chunk + o.constraint_chunks;
  bool boundary = false;
  if (overseg->n_compound > 0) { ++r->read_chunks; boundary = true; }
  if (r->read_chunks > 0 && r->read_chunks % o.chunk_set_size == 0 && boundary) RS_RC(r->boundary(false));
  if (!r->seg) r->seg.reset(new ChunkSet(o, r->w, r->h, r->chunk_sets, &r->dev));
  if (r->read_chunks % o.chunk_set_size >= overlap_start_chunk) {
    if (!r->new_seg) r->new_seg.reset(new ChunkSet(o, r->w, r->h, r->chunk_sets + 1, &r->dev));
    if (r->overlap_start < 0) r->overlap_start = r->seg->frames();
    if (boundary) {
      const double tb = wall_ms();
      std::unordered_map<int, Node*> mapping;
      std::unordered_map<int, Node*>* mp = (r->read_chunks % o.chunk_set_size < lookahead_start_chunk) ? &mapping : nullptr;
      RS_RC(r->seg->begin_chunk(*overseg, nullptr, mp));
      RS_RC(r->new_seg->begin_chunk(*overseg, mp, nullptr));
      r->dev.t_begin += wall_ms() - tb;
    }
    RS_RC(r->seg->add_frame(*overseg, flow_valid));
    RS_RC(r->new_seg->add_frame(*overseg, flow_valid));
  } else {
    if (boundary) { const double tb = wall_ms(); RS_RC(r->seg->begin_chunk(*overseg, nullptr, nullptr)); r->dev.t_begin += wall_ms() - tb; }
    RS_RC(r->seg->add_frame(*overseg, flow_valid));
  }
  if (r->read_chunks % o.chunk_set_size >= lookahead_start_chunk && r->lookahead_start < 0) r->lookahead_start = r->seg->frames();
  if (n_ready) *n_ready = (int)(r->ready.size() - before);
  return VSB200_OK;
}

int vsb200_region_flush(vsb200_region* r, int* n_ready) {
  if (n_ready) *n_ready = 0;
  if (!r) return VSB200_ERR_INVALID;
  if (r->flushed || !r->seg) { r->flushed = true; return VSB200_OK; }
  RS_CUDA(cudaSetDevice(r->device));
  const size_t before = r->ready.size();
  RS_RC(r->boundary(true));
  r->new_seg.reset();
  r->flushed = true;
  if (n_ready) *n_ready = (int)(r->ready.size() - before);
  return VSB200_OK;
}

long long vsb200_region_pop(vsb200_region* r, const int32_t** record) {
  if (!r || !record || r->ready.empty()) return 0;
  r->last = std::move(r->ready.front());
  r->ready.pop_front();
  *record = r->last->flat.data();
  return (long long)r->last->flat.size();
}

void vsb200_region_stats(vsb200_region* r, double out[2]) {
  if (!r || !out) return;
  out[0] = (double)r->dev.launches;
  out[1] = (double)r->chunk_sets;
}

void vsb200_region_destroy(vsb200_region* r) { delete r; }

}  // extern "C"
