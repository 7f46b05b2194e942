// engine.cu -- streaming dense over-segmentation engine behind the C ABI (include/vsb200.h).
// Host-side mirror of DenseSegmentation (segmentation/dense_segmentation.cpp:50-162,268-432) and of
// the result half of Segmentation (segmentation/segmentation.cpp:392-582,671-773): chunk
// arithmetic, constraint hand-over between chunks, region ids, per-frame SegmentationDesc
// arrays.  Every per-pixel / per-edge step is a CUDA kernel launch (preprocess.cu, edges.cu,
// sort.cu, merge.cu, results.cu); the host only keeps O(#regions + #scan intervals) bookkeeping.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/vsb200.h"
#include "common.cuh"
#include "results.cuh"
#include "shape.cuh"
#include "tubes.hpp"

using namespace vsb;

namespace {

// One over-segmentation region of the current chunk (RegionInformation, segmentation_common.h:39-116).
struct Region {
  int index = -1;          // position in the chunk's region list (first-seen order)
  int label = -1;          // device label (representative node id, or a fresh id for a split-off tube)
  int size = 0;
  int constrained_id = -1;
  int region_id = -1;
  bool removed = false;    // FLAGGED_FOR_REMOVAL
  std::vector<int> neighbors;                       // sorted region indices
  std::vector<vsbt::Piece> pieces;                  // its N4 components, ascending by (frame, first scan interval)
  std::vector<std::pair<int, int>> frames;          // (frame, area in that frame), ascending
};

int bits_for(unsigned long long n) { int b = 1; while ((1ull << b) < n) ++b; return b; }

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct FrameOut {
  int width, height, chunk_id, chunk_size, overlap_start, hierarchy_frame_idx, connectedness;
  std::vector<int32_t> region_id, interval_offset, intervals;
  std::vector<float> moments;
  std::vector<int32_t> compound, neighbor_offset, neighbor_id;
  std::vector<int32_t> id_map;   // optional
  int64_t pts;
};

#define ENG_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VSB200_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)
#define ENG_RC(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

}  // namespace

struct vsb200_dense {
  vsb200_dense_opts o;
  int w = 0, h = 0, n = 0;
  bool use_flow = false, l1 = false;
  cudaStream_t stream = nullptr;
  // DenseSegmentation members (dense_segmentation.h:198-230)
  int input_frames = 0, chunk_id = 0, overlap_frames = 2, constraint_frames = 1;
  int max_region_id = 0, num_output_frames = 0, curr_chunk_start = 0;
  bool flushed = false;            // PostProcess ran: the chain is finished, further pushes are refused
  bool import_pending = false;     // import_halo done, the group's first frame (slot 1) not pushed yet
  int buffered = 0;              // == feature_buffer_.size(): graph slots in use
  int max_slots = 0;
  int min_region_size = 0;
  // device memory
  uint8_t* d_bgr = nullptr; uint8_t* h_bgr = nullptr;     // staging (pinned host + device)
  float* h_flow_stage = nullptr;
  void* d_pre_scratch = nullptr;
  std::vector<float*> d_frames, d_spatial, d_temporal;     // per slot
  float* d_flows = nullptr;                                  // [max_slots][n][2]
  std::vector<std::vector<float>> h_flows;                   // host copies per slot (tube matching)
  uint32_t* d_codes = nullptr; unsigned long long* d_bstart = nullptr; void* d_sort_scratch = nullptr;
  size_t sort_scratch_bytes_ = 0;
  int* d_parent = nullptr; RegionRec* d_rec = nullptr;
  MergeParams mp{};
  unsigned long long live_cap_alloc = 0;
  int* d_labels = nullptr; int* d_roots = nullptr; int* d_idimg = nullptr; int* d_size_adjust = nullptr;
  int* d_slice_ids = nullptr; unsigned* d_row_counts = nullptr; unsigned* d_row_offsets = nullptr;
  unsigned* d_total = nullptr;
  RunRec* d_runs = nullptr; size_t runs_cap = 0;
  // shape stage (shape.cu): components of the runs, the two sorted run orders and their groups
  int* d_cc_parent = nullptr; int* d_group_of_run = nullptr; int* d_group_tab = nullptr;
  unsigned* d_skeys[2] = {nullptr, nullptr}; unsigned* d_svals[2] = {nullptr, nullptr};
  unsigned* d_shist = nullptr; unsigned* d_head_pos = nullptr; unsigned* d_tile_counts = nullptr; unsigned* d_tile_bases = nullptr; unsigned* d_ngroups = nullptr;
  RunGroup* d_groups[2] = {nullptr, nullptr}; int3* d_intervals[2] = {nullptr, nullptr};
  size_t shape_cap = 0;
  RunGroup* h_groups[2] = {nullptr, nullptr}; size_t h_groups_cap[2] = {0, 0};     // pinned
  vsbs::Interval* h_intervals[2] = {nullptr, nullptr}; size_t h_intervals_cap[2] = {0, 0};
  int* d_tmp_ids = nullptr; int2* d_tmp_info = nullptr; size_t tmp_cap = 0;
  unsigned long long* d_pair_table = nullptr; unsigned long long* d_pairs = nullptr; unsigned long long* d_pair_count = nullptr;
  unsigned pair_table_cap = 1u << 22; unsigned long long pairs_cap = 1u << 21;
  int* d_con_ids[2] = {nullptr, nullptr};   // constraint id maps for slot 0 (virtual) and slot 1
  int* d_first_of_id = nullptr; size_t first_of_id_cap = 0;
  // host results
  std::vector<std::unique_ptr<FrameOut>> overlap_out;   // overlap_segmentations_
  std::deque<std::unique_ptr<FrameOut>> ready;
  std::deque<int64_t> pts_queue;
  std::unique_ptr<FrameOut> last_popped;
  std::vector<uint8_t> proto_buf;
  bool have_import = false;
  double stats[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  double h2d_bytes = 0, d2h_bytes = 0, edge_ms = 0, edge_launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> edge_events;   // timing of the edge-build launches
  int time_edges = 0;
  int host_threads = 1;          // worker threads of the O(#scan intervals) host shaping

  ~vsb200_dense() { release(); }
  void release();
  int init();
  int push(const uint8_t* bgr, int stride, const float* flow, int flow_stride, int64_t pts, int* n_ready, bool device_input);
  int flush(int* n_ready);
  int add_frame_to_graph(int slot, const int* d_constraints);
  int start_constrained_chunk();
  int chunk_boundary(bool flush_all, int* n_ready);
  int segment_and_output(bool flush_all, std::vector<std::unique_ptr<FrameOut>>* results);
  int merge_constrained_regions(int slots);
  int upload_id_map(const FrameOut& f, int* dst);
  int ensure_tmp_capacity(size_t n_regions);
  int ensure_shape_capacity(size_t n_runs);
  int ensure_host_groups(int which, size_t n_groups, size_t n_runs);
};

void vsb200_dense::release() {
  auto F = [](void* p) { if (p) cudaFree(p); };
  F(d_bgr); F(d_pre_scratch); F(d_flows); F(d_codes); F(d_bstart); F(d_sort_scratch); F(d_parent); F(d_rec);
  F(mp.res); F(mp.acc); F(mp.cl); F(mp.hull); F(mp.live_a); F(mp.live_b); F(mp.live_c); F(mp.done); F(mp.counters); F(mp.debug); F(mp.scan_queue);
  F(d_labels); F(d_roots); F(d_idimg); F(d_size_adjust); F(d_slice_ids); F(d_row_counts); F(d_row_offsets); F(d_total);
  F(d_runs); F(d_tmp_ids); F(d_tmp_info); F(d_pair_table); F(d_pairs); F(d_pair_count);
  F(d_con_ids[0]); F(d_con_ids[1]); F(d_first_of_id);
  F(d_cc_parent); F(d_group_of_run); F(d_group_tab); F(d_shist); F(d_head_pos); F(d_tile_counts); F(d_tile_bases); F(d_ngroups);
  for (int k = 0; k < 2; ++k) {
    F(d_skeys[k]); F(d_svals[k]); F(d_groups[k]); F(d_intervals[k]);
    if (h_groups[k]) cudaFreeHost(h_groups[k]);
    if (h_intervals[k]) cudaFreeHost(h_intervals[k]);
  }
  for (auto p : d_frames) F(p);
  for (auto p : d_spatial) F(p);
  for (auto p : d_temporal) F(p);
  if (h_bgr) cudaFreeHost(h_bgr);
  if (h_flow_stage) cudaFreeHost(h_flow_stage);
  if (stream) cudaStreamDestroy(stream);
  d_bgr = nullptr; stream = nullptr;
}

int vsb200_dense::init() {
  ENG_CUDA(cudaSetDevice(o.device));
  ENG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  n = w * h;
  max_slots = o.chunk_size + 1;                 // later chunks: virtual + constrained + chunk_size - 1
  const size_t nodes = (size_t)n * max_slots;
  if (!edge_codes_fit(2 * max_slots - 1, (unsigned long long)n) || nodes >= (1ull << 31)) {
    set_error("frame %dx%d x %d slots overflows the 32-bit edge code / node id", w, h, max_slots);
    return VSB200_ERR_UNSUPPORTED;
  }
  // min_region_size: float product truncated (dense_segmentation.cpp:270-272)
  min_region_size = (int)(o.frac_min_region_size * w * o.frac_min_region_size * h * o.chunk_size);
  ENG_CUDA(cudaMalloc(&d_bgr, (size_t)n * 3));
  ENG_CUDA(cudaMallocHost(&h_bgr, (size_t)n * 3));
  ENG_CUDA(cudaMalloc(&d_pre_scratch, preprocess_scratch_bytes()));
  d_frames.assign(max_slots, nullptr); d_spatial.assign(max_slots, nullptr); d_temporal.assign(max_slots, nullptr);
  for (int s = 0; s < max_slots; ++s) {
    ENG_CUDA(cudaMalloc(&d_frames[s], (size_t)n * 3 * sizeof(float)));
    ENG_CUDA(cudaMalloc(&d_spatial[s], (size_t)n * 4 * sizeof(float)));
    ENG_CUDA(cudaMalloc(&d_temporal[s], (size_t)n * 9 * sizeof(float)));
  }
  if (use_flow) {
    ENG_CUDA(cudaMalloc(&d_flows, nodes * 2 * sizeof(float)));
    ENG_CUDA(cudaMemsetAsync(d_flows, 0, nodes * 2 * sizeof(float), stream));
    ENG_CUDA(cudaMallocHost(&h_flow_stage, (size_t)n * 2 * sizeof(float)));
    h_flows.assign(max_slots, std::vector<float>());
  }
  const size_t total_elems = (size_t)n * (4 * max_slots + 9 * (max_slots - 1));
  ENG_CUDA(cudaMalloc(&d_codes, total_elems * sizeof(uint32_t)));
  ENG_CUDA(cudaMalloc(&d_bstart, sizeof(unsigned long long) * (kNumBuckets + 1)));
  sort_scratch_bytes_ = sort_scratch_bytes(2 * max_slots - 1, w, h);
  ENG_CUDA(cudaMalloc(&d_sort_scratch, sort_scratch_bytes_));
  ENG_CUDA(cudaMalloc(&d_parent, nodes * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_rec, nodes * sizeof(RegionRec)));
  ENG_CUDA(cudaMalloc(&mp.res, nodes * 8));
  ENG_CUDA(cudaMalloc(&mp.acc, nodes * 32));
  ENG_CUDA(cudaMalloc(&mp.cl, nodes * 4));
  ENG_CUDA(cudaMalloc(&mp.hull, nodes * sizeof(NodeScratch)));
  ENG_CUDA(cudaMalloc(&mp.counters, (16 + 4096) * 8));
  ENG_CUDA(cudaMalloc(&mp.scan_queue, kScanQueueWords * sizeof(uint32_t)));   // + per-block counts of the ordered compaction
  mp.stats = mp.counters + 8;
  mp.debug = nullptr;
  mp.trace = nullptr;
  ENG_CUDA(cudaMemsetAsync(mp.res, 0xff, nodes * 8, stream));
  ENG_CUDA(cudaMemsetAsync(mp.acc, 0, nodes * 32, stream));
  ENG_CUDA(cudaMemsetAsync(mp.counters, 0, (16 + 4096) * 8, stream));
  ENG_RC(launch_init_iota(mp.cl, (long long)nodes, stream));
  ENG_RC(launch_init_hull(mp.hull, (long long)nodes, stream));
  ENG_CUDA(cudaMalloc(&d_labels, nodes * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_idimg, nodes * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_roots, nodes * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_size_adjust, (nodes + 1) * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_slice_ids, max_slots * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_row_counts, (size_t)max_slots * h * sizeof(unsigned)));
  ENG_CUDA(cudaMalloc(&d_row_offsets, (size_t)max_slots * h * sizeof(unsigned)));
  ENG_CUDA(cudaMalloc(&d_total, sizeof(unsigned)));
  ENG_CUDA(cudaMalloc(&d_pair_table, sizeof(unsigned long long) * pair_table_cap));
  ENG_CUDA(cudaMalloc(&d_pairs, sizeof(unsigned long long) * pairs_cap));
  ENG_CUDA(cudaMalloc(&d_pair_count, sizeof(unsigned long long)));
  ENG_CUDA(cudaMalloc(&d_con_ids[0], (size_t)n * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_con_ids[1], (size_t)n * sizeof(int)));
  ENG_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

// Per-region scratch (ids in, records out), sized by the number of regions of the chunk.
int vsb200_dense::ensure_tmp_capacity(size_t n_regions) {
  if (n_regions <= tmp_cap) return 0;
  if (d_tmp_ids) cudaFree(d_tmp_ids);
  if (d_tmp_info) cudaFree(d_tmp_info);
  d_tmp_ids = nullptr; d_tmp_info = nullptr; tmp_cap = 0;
  const size_t cap = n_regions * 2 + 1024;
  ENG_CUDA(cudaMalloc(&d_tmp_ids, cap * 2 * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_tmp_info, cap * sizeof(RegionRec)));
  tmp_cap = cap;
  return 0;
}

// Device buffers of the shape stage, sized by the number of scan intervals of the chunk.
int vsb200_dense::ensure_shape_capacity(size_t n_runs) {
  if (n_runs <= shape_cap) return 0;
  auto F = [](void* p) { if (p) cudaFree(p); };
  F(d_cc_parent); F(d_group_of_run); F(d_group_tab); F(d_shist); F(d_head_pos); F(d_tile_counts); F(d_tile_bases); F(d_ngroups);
  d_cc_parent = d_group_of_run = d_group_tab = nullptr;
  d_shist = d_head_pos = d_tile_counts = d_tile_bases = d_ngroups = nullptr;
  for (int k = 0; k < 2; ++k) {
    F(d_skeys[k]); F(d_svals[k]); F(d_groups[k]); F(d_intervals[k]);
    d_skeys[k] = d_svals[k] = nullptr; d_groups[k] = nullptr; d_intervals[k] = nullptr;
  }
  shape_cap = 0;                                 // a failed allocation below leaves a consistent (empty) state behind
  const size_t cap = n_runs + n_runs / 2 + 4096;
  ENG_CUDA(cudaMalloc(&d_cc_parent, cap * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_group_of_run, cap * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_group_tab, cap * 2 * sizeof(int)));
  ENG_CUDA(cudaMalloc(&d_shist, (cap / 2048 + 2) * 512 * sizeof(unsigned)));
  ENG_CUDA(cudaMalloc(&d_head_pos, cap * sizeof(unsigned)));
  ENG_CUDA(cudaMalloc(&d_tile_counts, (cap / 1024 + 2) * sizeof(unsigned)));
  ENG_CUDA(cudaMalloc(&d_tile_bases, (cap / 1024 + 2) * sizeof(unsigned)));
  ENG_CUDA(cudaMalloc(&d_ngroups, sizeof(unsigned)));
  for (int k = 0; k < 2; ++k) {
    ENG_CUDA(cudaMalloc(&d_skeys[k], cap * sizeof(unsigned)));
    ENG_CUDA(cudaMalloc(&d_svals[k], cap * sizeof(unsigned)));
    ENG_CUDA(cudaMalloc(&d_groups[k], cap * sizeof(RunGroup)));
    ENG_CUDA(cudaMalloc(&d_intervals[k], cap * sizeof(int3)));
  }
  shape_cap = cap;
  return 0;
}

// Pinned host copies of one pass's groups and intervals.
int vsb200_dense::ensure_host_groups(int which, size_t n_groups, size_t n_runs) {
  if (n_groups > h_groups_cap[which]) {
    if (h_groups[which]) cudaFreeHost(h_groups[which]);
    h_groups[which] = nullptr; h_groups_cap[which] = 0;
    const size_t cap = n_groups + n_groups / 2 + 1024;
    ENG_CUDA(cudaMallocHost(&h_groups[which], cap * sizeof(RunGroup)));
    h_groups_cap[which] = cap;
  }
  if (n_runs > h_intervals_cap[which]) {
    if (h_intervals[which]) cudaFreeHost(h_intervals[which]);
    h_intervals[which] = nullptr; h_intervals_cap[which] = 0;
    const size_t cap = n_runs + n_runs / 2 + 4096;
    ENG_CUDA(cudaMallocHost(&h_intervals[which], cap * sizeof(vsbs::Interval)));
    h_intervals_cap[which] = cap;
  }
  return 0;
}

// AddGenericImage[Constrained] + ConnectTemporally (segmentation.h:310-338): nodes of the slot and
// the weights of bucket lists 2*slot (spatial) and 2*slot-1 (temporal to slot-1).
int vsb200_dense::add_frame_to_graph(int slot, const int* d_constraints) {
  ENG_RC(launch_init_nodes(d_frames[slot], d_constraints, slot, w, h, d_parent, d_rec, stream));
  const bool temporal = slot >= 1 && !(chunk_id > 0 && slot == 1);   // slot 1 of later chunks: virtual edges only
  const float* flow = (use_flow && temporal) ? d_flows + (size_t)slot * n * 2 : nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (time_edges && temporal) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventRecord(e0, stream); }
  ENG_RC(launch_edge_build(d_frames[slot], temporal ? d_frames[slot - 1] : nullptr, flow, w, h, l1, d_spatial[slot],
                           temporal ? d_temporal[slot] : nullptr, stream));
  if (e0) { cudaEventRecord(e1, stream); edge_events.emplace_back(e0, e1); }
  stats[7] += 2;
  return 0;
}

int vsb200_dense::push(const uint8_t* bgr, int stride, const float* flow, int flow_stride, int64_t pts, int* n_ready, bool device_input) {
  if (n_ready) *n_ready = 0;
  if (!bgr || stride < w * 3) { set_error("push: bad frame buffer"); return VSB200_ERR_INVALID; }
  if (use_flow && (input_frames > 0 || import_pending) && !flow) { set_error("push: flow missing (created with use_flow)"); return VSB200_ERR_INVALID; }
  if (buffered >= max_slots) { set_error("push: internal slot overflow"); return VSB200_ERR_INVALID; }
  if (flushed) { set_error("push after flush: a flushed handle is finished (create a new one for the next sequence)"); return VSB200_ERR_INVALID; }
  ENG_CUDA(cudaSetDevice(o.device));
  const double t0 = now_ms();
  const int slot = buffered;
  if (device_input) {
    // frame already resident in HBM (bench.py `value` leg / on-device decoders): no staging copy
    ENG_RC(launch_preprocess(bgr, stride, w, h, o.presmoothing, d_frames[slot], d_pre_scratch, stream));
  } else {
    // H2D: pinned staging (the caller may release its buffer when push returns)
    for (int y = 0; y < h; ++y) memcpy(h_bgr + (size_t)y * w * 3, bgr + (size_t)y * stride, (size_t)w * 3);
    ENG_CUDA(cudaMemcpyAsync(d_bgr, h_bgr, (size_t)n * 3, cudaMemcpyHostToDevice, stream));
    h2d_bytes += (double)n * 3;
    ENG_RC(launch_preprocess(d_bgr, w * 3, w, h, o.presmoothing, d_frames[slot], d_pre_scratch, stream));
  }
  stats[7] += (o.presmoothing == 2) ? 3 : 1;
  if (use_flow) {
    if ((input_frames == 0 && !import_pending) || !flow) {
      h_flows[slot].clear();
      ENG_CUDA(cudaMemsetAsync(d_flows + (size_t)slot * n * 2, 0, (size_t)n * 2 * sizeof(float), stream));
    } else {
      h_flows[slot].resize((size_t)n * 2);
      for (int y = 0; y < h; ++y)
        memcpy(h_flows[slot].data() + (size_t)y * w * 2, (const char*)flow + (size_t)y * flow_stride, sizeof(float) * 2 * w);
      ENG_CUDA(cudaStreamSynchronize(stream));       // staging buffer reuse
      memcpy(h_flow_stage, h_flows[slot].data(), sizeof(float) * 2 * n);
      ENG_CUDA(cudaMemcpyAsync(d_flows + (size_t)slot * n * 2, h_flow_stage, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, stream));
      h2d_bytes += (double)n * 8;
    }
  }
  if (import_pending) {
    // first frame of a successor group: it is the predecessor's second overlap frame, re-added under the
    // imported labels next to the virtual slot (ChunkBoundaryOutput, dense_segmentation.cpp:290-328)
    ENG_RC(start_constrained_chunk());
    import_pending = false;
  } else {
    ENG_RC(add_frame_to_graph(slot, nullptr));
  }
  ENG_CUDA(cudaStreamSynchronize(stream));            // h_bgr is reused by the next push
  stats[0] += now_ms() - t0;
  pts_queue.push_back(pts);            // after the last step that can fail: pts and results stay in step
  ++buffered;
  ++input_frames;
  if (buffered - curr_chunk_start >= o.chunk_size) return chunk_boundary(false, n_ready);
  return 0;
}

int vsb200_dense::flush(int* n_ready) {
  if (n_ready) *n_ready = 0;
  if (buffered == 0) { flushed = true; return 0; }
  ENG_CUDA(cudaSetDevice(o.device));
  const int rc = chunk_boundary(true, n_ready);
  if (rc == 0) flushed = true;
  return rc;
}

// Renders the region-id image of a result (SegmentationDescToIdImage, segmentation_util.cpp:741-770)
// and uploads it as the constraint ids of the next chunk.
int vsb200_dense::upload_id_map(const FrameOut& f, int* dst) {
  std::vector<int32_t> img((size_t)n, 0);
  for (size_t k = 0; k < f.region_id.size(); ++k)
    for (int q = f.interval_offset[k]; q < f.interval_offset[k + 1]; ++q) {
      const int y = f.intervals[3 * q], lx = f.intervals[3 * q + 1], rx = f.intervals[3 * q + 2];
      std::fill(img.begin() + (size_t)y * w + lx, img.begin() + (size_t)y * w + rx + 1, f.region_id[k]);
    }
  ENG_CUDA(cudaMemcpyAsync(dst, img.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice, stream));
  ENG_CUDA(cudaStreamSynchronize(stream));
  return 0;
}

// ChunkBoundaryOutput, second half (dense_segmentation.cpp:290-328): slot 0 = virtual copy of the
// last output frame, slot 1 = next frame re-added with its previous labels as constraints.
int vsb200_dense::start_constrained_chunk() {
  // d_con_ids[0/1] hold the id maps; slot 1 already owns its smoothed frame (pointer swap done by caller)
  size_t need = (size_t)max_region_id + 2;
  if (need > first_of_id_cap) {
    if (d_first_of_id) cudaFree(d_first_of_id);
    first_of_id_cap = need * 2;
    ENG_CUDA(cudaMalloc(&d_first_of_id, first_of_id_cap * sizeof(int)));
  }
  ENG_RC(launch_init_virtual_nodes(d_con_ids[0], 0, w, h, d_parent, d_rec, d_first_of_id, max_region_id + 1, stream));
  ENG_RC(add_frame_to_graph(1, d_con_ids[1]));
  stats[7] += 3;
  return 0;
}

int vsb200_dense::chunk_boundary(bool flush_all, int* n_ready) {
  std::vector<std::unique_ptr<FrameOut>> results;
  ENG_RC(segment_and_output(flush_all, &results));
  for (auto& r : results) {
    r->pts = pts_queue.front();
    pts_queue.pop_front();
    ready.push_back(std::move(r));
  }
  if (n_ready) *n_ready = (int)results.size();
  if (flush_all) return 0;
  // new Segmentation with curr_chunk_start_ + chunk_size slots (dense_segmentation.cpp:294-298)
  ENG_RC(upload_id_map(*overlap_out[0], d_con_ids[0]));
  ENG_RC(upload_id_map(*overlap_out[1], d_con_ids[1]));
  ENG_RC(start_constrained_chunk());
  overlap_out.clear();
  return 0;
}

// MergeConstrainedRegions (segmentation_graph.h:703-786), host assisted.  The reference walks every non-virtual node
// whose OWN (possibly stale) record still carries a constraint id, in node order, and handles the node's current
// representative: look its constraint id up in a map; insert it, or merge with / split from the map's holder.  Repeated
// visits of one representative with nothing in between are idempotent from the fourth on, so the device lists, in node
// order, the first four nodes of every run of visited nodes with one representative (O(#scan runs) records) and the
// host replays exactly those visits; then the virtual nodes, which always merge (:763-785).
namespace {
__global__ void mcr_visits_kernel(const int* __restrict__ labels, const RegionRec* __restrict__ rec, int base, long long total,
                                  int2* __restrict__ out, unsigned* __restrict__ count, unsigned cap) {
  for (long long i = base + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    if (rec[i].con < 0) continue;
    const int root = labels[i];
    int pos = 0;                                  // position inside the run of visited nodes with this representative
    while (pos < 4 && i - pos - 1 >= base && rec[i - pos - 1].con >= 0 && labels[i - pos - 1] == root) ++pos;
    if (pos >= 4) continue;
    const unsigned k = atomicAdd(count, 1u);
    if (k < cap) out[k] = make_int2((int)i, root);
  }
}
__global__ void gather_i32_kernel(const int* __restrict__ ids, int m, const int* __restrict__ src, int* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = src[ids[i]];
}
__global__ void gather_rec_kernel(const int* __restrict__ ids, int m, const RegionRec* __restrict__ rec, RegionRec* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < m) out[i] = rec[ids[i]];
}
__global__ void scatter_rec_kernel(const int* __restrict__ ids, const int* __restrict__ parents, int m,
                                   const RegionRec* __restrict__ recs, RegionRec* __restrict__ rec, int* __restrict__ parent) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m) return;
  // only the merge-relevant fields of a representative change here; a node's own stale constraint id stays as it is
  RegionRec r = recs[i];
  rec[ids[i]] = r;
  parent[ids[i]] = parents[i];
}
}  // namespace

int vsb200_dense::merge_constrained_regions(int slots) {
  const long long total = (long long)n * slots;
  ENG_RC(launch_flatten(d_parent, nullptr, d_labels, total, stream));
  // visit records: at most four per scan run of the label volume; sized generously, grown on overflow
  if (runs_cap < (size_t)n) {
    if (d_runs) cudaFree(d_runs);
    runs_cap = (size_t)n * 2;
    ENG_CUDA(cudaMalloc(&d_runs, runs_cap * sizeof(RunRec)));
  }
  unsigned cnt = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const unsigned cap = (unsigned)std::min<size_t>(runs_cap * sizeof(RunRec) / sizeof(int2), 0x7fffffffu);
    ENG_CUDA(cudaMemsetAsync(d_total, 0, sizeof(unsigned), stream));
    mcr_visits_kernel<<<148 * 8, 256, 0, stream>>>(d_labels, d_rec, n, total, (int2*)d_runs, d_total, cap);
    ENG_CUDA(cudaMemcpyAsync(&cnt, d_total, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
    ENG_CUDA(cudaStreamSynchronize(stream));
    if (cnt <= cap) break;
    cudaFree(d_runs);
    runs_cap = ((size_t)cnt * sizeof(int2) / sizeof(RunRec) + 1) * 2;
    ENG_CUDA(cudaMalloc(&d_runs, runs_cap * sizeof(RunRec)));
  }
  std::vector<int2> list(cnt);
  if (cnt) ENG_CUDA(cudaMemcpy(list.data(), d_runs, sizeof(int2) * cnt, cudaMemcpyDeviceToHost));
  d2h_bytes += 8.0 * cnt;
  std::sort(list.begin(), list.end(), [](const int2& a, const int2& b) { return a.x < b.x; });
  std::vector<int> ids;               // device node ids whose records take part
  std::unordered_map<int, int> pos;   // node id -> local index
  pos.reserve(1 << 14);
  auto local = [&](int id) { auto it = pos.find(id); if (it != pos.end()) return it->second; pos[id] = (int)ids.size(); ids.push_back(id); return (int)ids.size() - 1; };
  std::vector<int> visit_rep(cnt);
  for (unsigned k = 0; k < cnt; ++k) visit_rep[k] = local(list[k].y);
  std::vector<std::pair<int, int>> virt;   // (constraint id, local index of the virtual representative), first pixel of every id on slot 0
  {
    std::vector<int> h_first((size_t)max_region_id + 1);
    ENG_CUDA(cudaMemcpy(h_first.data(), d_first_of_id, sizeof(int) * ((size_t)max_region_id + 1), cudaMemcpyDeviceToHost));
    // the virtual representative's CURRENT representative (virtual nodes merge during the scan like any other node)
    std::vector<int> want;
    for (int c = 0; c <= max_region_id; ++c) if (h_first[c] != 0x7f7f7f7f) want.push_back(h_first[c]);
    std::vector<int> roots(want.size());
    if (!want.empty()) {
      ENG_RC(ensure_tmp_capacity(want.size()));
      ENG_CUDA(cudaMemcpyAsync(d_tmp_ids, want.data(), sizeof(int) * want.size(), cudaMemcpyHostToDevice, stream));
      gather_i32_kernel<<<(unsigned)((want.size() + 255) / 256), 256, 0, stream>>>(d_tmp_ids, (int)want.size(), d_labels, d_tmp_ids + tmp_cap);
      ENG_CUDA(cudaMemcpyAsync(roots.data(), d_tmp_ids + tmp_cap, sizeof(int) * want.size(), cudaMemcpyDeviceToHost, stream));
      ENG_CUDA(cudaStreamSynchronize(stream));
    }
    size_t k = 0;
    for (int c = 0; c <= max_region_id; ++c)
      if (h_first[c] != 0x7f7f7f7f) virt.emplace_back(c, local(roots[k++]));
  }
  const int m = (int)ids.size();
  if (m == 0) return 0;
  ENG_RC(ensure_tmp_capacity((size_t)m));
  RegionRec* d_recs = (RegionRec*)d_tmp_info;
  ENG_CUDA(cudaMemcpyAsync(d_tmp_ids, ids.data(), sizeof(int) * m, cudaMemcpyHostToDevice, stream));
  gather_rec_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_tmp_ids, m, d_rec, d_recs);
  std::vector<RegionRec> recs(m);
  ENG_CUDA(cudaMemcpyAsync(recs.data(), d_recs, sizeof(RegionRec) * m, cudaMemcpyDeviceToHost, stream));
  ENG_CUDA(cudaStreamSynchronize(stream));
  std::vector<int> par(m);
  for (int i = 0; i < m; ++i) par[i] = i;
  auto find = [&](int x) { while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; } return x; };
  auto dist = [&](const RegionRec& a, const RegionRec& b) {
    const float d1 = a.d0 - b.d0, d2 = a.d1 - b.d1, d3 = a.d2 - b.d2;
    return std::sqrt((d1 * d1 + d2 * d2 + d3 * d3) * (1.0f / 3.0f));   // edge weight 1.0: no force merge
  };
  unsigned long long version = 1;                    // bumped by every change of the walk's state
  std::vector<unsigned long long> settled(m, 0);     // version at which a representative was last found to be a no-op
  auto merge = [&](int a, int b) {                 // MergeRegions(rep_1 = a, rep_2 = b)
    const bool aw = recs[a].sz > recs[b].sz;
    const int mi = aw ? a : b, oi = aw ? b : a;
    RegionRec& M = recs[mi];
    const RegionRec& O = recs[oi];
    const float denom = 1.0f / (float)(O.sz + M.sz);
    const float fa = (float)O.sz * denom, fb = (float)M.sz * denom;
    M.d0 = fa * O.d0 + fb * M.d0; M.d1 = fa * O.d1 + fb * M.d1; M.d2 = fa * O.d2 + fb * M.d2;
    M.sz += O.sz;
    M.con = std::max(recs[a].con, recs[b].con);
    par[oi] = mi;
    ++version;
  };
  std::unordered_map<int, int> con2rep;
  auto visit = [&](int my) {
    // one step of the non-virtual loop (:722-760) for representative `my`
    if (settled[my] == version) return;               // nothing changed since this representative was last a no-op
    auto it = con2rep.find(recs[my].con);
    if (it == con2rep.end()) { con2rep[recs[my].con] = my; ++version; settled[my] = version; return; }
    const int cr = find(it->second);
    if (cr == my) { settled[my] = version; return; }
    const float d = dist(recs[my], recs[cr]);
    if (d > 0.15f) {
      bool changed = false;
      if ((double)recs[my].sz < (double)recs[cr].sz * 0.3) {
        if (recs[my].con != -1) { recs[my].con = -1; changed = true; }
      } else if ((double)recs[cr].sz < (double)recs[my].sz * 0.3) {
        if (recs[cr].con != -1) { recs[cr].con = -1; changed = true; }
        if (it->second != my) { it->second = my; changed = true; }
      } else {
        recs[my].con = -1; recs[cr].con = -1; con2rep.erase(it); changed = true;
      }
      if (changed) ++version; else settled[my] = version;     // unchanged: the same comparison would only repeat itself
    } else {
      merge(my, cr);
    }
  };
  for (unsigned k = 0; k < cnt; ++k) visit(find(visit_rep[k]));
  for (const auto& v : virt) {                        // virtual nodes: always merge (:763-785)
    const int my = find(v.second);
    auto it = con2rep.find(recs[my].con);
    if (it == con2rep.end()) { con2rep[recs[my].con] = my; continue; }
    const int cr = find(it->second);
    if (cr != my) merge(my, cr);
  }
  std::vector<int> parents(m);
  for (int i = 0; i < m; ++i) parents[i] = ids[find(i)];
  ENG_CUDA(cudaMemcpyAsync(d_tmp_ids + tmp_cap, parents.data(), sizeof(int) * m, cudaMemcpyHostToDevice, stream));
  ENG_CUDA(cudaMemcpyAsync(d_recs, recs.data(), sizeof(RegionRec) * m, cudaMemcpyHostToDevice, stream));
  scatter_rec_kernel<<<(m + 255) / 256, 256, 0, stream>>>(d_tmp_ids, d_tmp_ids + tmp_cap, m, d_recs, d_rec, d_parent);
  ENG_CUDA(cudaStreamSynchronize(stream));
  stats[7] += 4;
  return 0;
}

// SegmentAndOutputChunk (dense_segmentation.cpp:330-432) with RunOverSegmentation
// (segmentation.cpp:272-303) inlined.
int vsb200_dense::segment_and_output(bool flush_all, std::vector<std::unique_ptr<FrameOut>>* results) {
  const int slots = buffered;
  const bool constrained_chunk = chunk_id > 0;
  if (!edge_events.empty()) {
    cudaStreamSynchronize(stream);
    for (auto& ev2 : edge_events) { float ms = 0; cudaEventElapsedTime(&ms, ev2.first, ev2.second); edge_ms += ms; edge_launches += 1; cudaEventDestroy(ev2.first); cudaEventDestroy(ev2.second); }
    edge_events.clear();
  }
  struct EventSet {          // destroyed on every exit path
    cudaEvent_t e[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    EventSet() { for (auto& x : e) cudaEventCreate(&x); }
    ~EventSet() { for (auto& x : e) if (x) cudaEventDestroy(x); }
    cudaEvent_t& operator[](int i) { return e[i]; }
  } ev;
  // ---------------- sort ----------------
  const int num_lists = 2 * slots - 1;
  std::vector<const float*> seg(num_lists, nullptr);
  for (int s = 0; s < slots; ++s) {
    if (!(constrained_chunk && s == 0)) seg[2 * s] = d_spatial[s];              // virtual slot: no spatial edges
    if (s >= 1 && !(constrained_chunk && s == 1)) seg[2 * s - 1] = d_temporal[s];   // virtual edges carry no weight
  }
  cudaEventRecord(ev[0], stream);
  ENG_RC(launch_sort_edges(seg.data(), num_lists, w, h, d_codes, d_bstart, d_sort_scratch, sort_scratch_bytes_, stream));
  stats[7] += 5;
  unsigned long long h_bstart[kNumBuckets + 1];
  ENG_CUDA(cudaMemcpyAsync(h_bstart, d_bstart, sizeof(h_bstart), cudaMemcpyDeviceToHost, stream));
  cudaEventRecord(ev[1], stream);
  ENG_CUDA(cudaStreamSynchronize(stream));
  unsigned long long max_bucket = 1;
  for (int b = 0; b < kNumBuckets; ++b) max_bucket = std::max(max_bucket, h_bstart[b + 1] - h_bstart[b]);
  if (max_bucket > live_cap_alloc) {
    if (mp.live_a) cudaFree(mp.live_a);
    if (mp.live_b) cudaFree(mp.live_b);
    live_cap_alloc = max_bucket + max_bucket / 4;
    if (mp.done) cudaFree(mp.done);
    if (mp.live_c) cudaFree(mp.live_c);
    ENG_CUDA(cudaMalloc(&mp.live_a, live_cap_alloc * 16));
    ENG_CUDA(cudaMalloc(&mp.live_b, live_cap_alloc * 16));
    ENG_CUDA(cudaMalloc(&mp.live_c, live_cap_alloc * 16));
    ENG_CUDA(cudaMalloc(&mp.done, live_cap_alloc));
  }
  // ---------------- merge ----------------
  mp.w = w; mp.h = h; mp.slots = slots; mp.min_region_size = min_region_size;
  mp.force_merge_weight = l1 ? 0.002f : 0.001f;
  mp.has_constraints = constrained_chunk ? 1 : 0;
  mp.flows = use_flow ? d_flows : nullptr;
  mp.codes = d_codes; mp.bucket_start = d_bstart; mp.parent = d_parent; mp.rec = d_rec;
  mp.live_cap = live_cap_alloc;
  ENG_CUDA(cudaMemsetAsync(mp.stats, 0, 8 * 8, stream));
  const char* dbg_path = getenv("VSB200_MERGE_DEBUG");       // per-bucket profile of every chunk (development tap)
  if (dbg_path && !mp.debug) ENG_CUDA(cudaMalloc(&mp.debug, (kNumBuckets * 4 + 64) * 8));
  if (mp.debug) ENG_CUDA(cudaMemsetAsync(mp.debug, 0, (kNumBuckets * 4 + 64) * 8, stream));
  ENG_RC(launch_merge(mp, stream));
  if (mp.debug && dbg_path) {
    unsigned long long h_scan = 0;
    cudaMemcpyAsync(&h_scan, mp.stats + 3, 8, cudaMemcpyDeviceToHost, stream);
    std::vector<unsigned long long> dbg(kNumBuckets * 4 + 64);
    ENG_CUDA(cudaMemcpyAsync(dbg.data(), mp.debug, dbg.size() * 8, cudaMemcpyDeviceToHost, stream));
    ENG_CUDA(cudaStreamSynchronize(stream));
    const std::string path = std::string(dbg_path) + ".chunk" + std::to_string(chunk_id);
    if (FILE* f = fopen(path.c_str(), "w")) {
      unsigned long long prev_r = 0, prev_w = 0;
      for (int b = 0; b < kNumBuckets; ++b) {
        if (dbg[b * 4 + 2] == 0) continue;
        fprintf(f, "%d edges %llu windows %llu rounds %llu us %.1f\n", b, dbg[b * 4 + 2], dbg[b * 4 + 3] - prev_w,
                dbg[b * 4 + 1] - prev_r, dbg[b * 4 + 0] / 1000.0);
        prev_r = dbg[b * 4 + 1]; prev_w = dbg[b * 4 + 3];
      }
      const unsigned long long* c = &dbg[kNumBuckets * 4 + 8];
      fprintf(f, "uncertified edge-attempts: hubhub %llu con %llu hubs3 %llu hubless_fin %llu hubless_diam %llu race %llu hub_not_frozen %llu bigmass %llu\n",
              c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
      fprintf(f, "phase ms: raw_prune %.1f certify %.1f ordered_rounds %.1f compaction %.1f hubhub_refresh %.1f exact_scan %.1f serial_mode %.1f; exact scans %llu (%llu edges), serial calls %llu (%llu edges)\n",
              c[12] / 1e6, c[13] / 1e6, c[14] / 1e6, c[15] / 1e6, c[16] / 1e6, c[17] / 1e6, c[18] / 1e6, h_scan, c[19], c[20], c[21]);
      fprintf(f, "certify passes ms: C1 union %.1f C2 records %.1f C3 hubs %.1f C4 count %.1f C5 apply %.1f C6 fold+reset %.1f\n",
              c[32] / 1e6, c[33] / 1e6, c[34] / 1e6, c[35] / 1e6, c[36] / 1e6, c[37] / 1e6);
      fprintf(f, "hub_not_frozen reasons: same_id_or_hubs3 %llu con_conflict %llu open_hubs3 %llu bound %llu\n", c[8], c[9], c[10], c[11]);
      fprintf(f, "scans: event thread Mcycles %.1f, total %.1f (staging+sort %.1f, sub-cluster replay %.1f); staged roots %llu, events %llu of %llu edges: absorb %llu big-big %llu generic+speculated %llu, hub swaps %llu; speculated meetings %llu, scans redone %llu\n", c[42] / 1e6, c[43] / 1e6, c[45] / 1e6, c[46] / 1e6, c[44], c[47], c[48], c[49], c[50], c[51], c[52], c[53], c[54]);
      fclose(f);
    }
  }
  stats[7] += 1;
  if (constrained_chunk) ENG_RC(merge_constrained_regions(slots));
  cudaEventRecord(ev[2], stream);
  // ---------------- labels, N4, RLE ----------------
  const size_t nodes = (size_t)n * slots;
  ENG_RC(launch_flatten(d_parent, nullptr, d_labels, (long long)nodes, stream));
  ENG_CUDA(cudaMemcpyAsync(d_idimg, d_labels, nodes * sizeof(int), cudaMemcpyDeviceToDevice, stream));
  ENG_CUDA(cudaMemcpyAsync(d_roots, d_labels, nodes * sizeof(int), cudaMemcpyDeviceToDevice, stream));
  ENG_CUDA(cudaMemsetAsync(d_size_adjust, 0, (nodes + 1) * sizeof(int), stream));
  std::vector<int> slice_ids;
  for (int s = 0; s < slots; ++s) if (!(constrained_chunk && s == 0)) slice_ids.push_back(s);
  const int ns = (int)slice_ids.size();
  ENG_CUDA(cudaMemcpyAsync(d_slice_ids, slice_ids.data(), ns * sizeof(int), cudaMemcpyHostToDevice, stream));
  if (o.enforce_n4_connectivity) ENG_RC(launch_n4(d_idimg, w, h, ns, d_slice_ids, d_size_adjust, stream));
  ENG_RC(launch_rle_count(d_idimg, w, h, d_slice_ids, ns, d_row_counts, stream));
  ENG_RC(launch_scan_u32(d_row_counts, d_row_offsets, d_total, ns * h, stream));
  unsigned n_runs = 0;
  ENG_CUDA(cudaMemcpyAsync(&n_runs, d_total, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
  ENG_CUDA(cudaStreamSynchronize(stream));
  const bool stage_debug = getenv("VSB200_STAGE_DEBUG") != nullptr;
  const double t_dbg0 = now_ms();
  if (n_runs > runs_cap) {
    if (d_runs) cudaFree(d_runs);
    d_runs = nullptr; runs_cap = 0;
    const size_t cap = (size_t)n_runs + n_runs / 2 + 1024;
    ENG_CUDA(cudaMalloc(&d_runs, cap * sizeof(RunRec)));
    runs_cap = cap;
  }
  ENG_RC(launch_rle_write(d_idimg, w, h, d_slice_ids, ns, d_row_offsets, d_runs, stream));
  // ---------------- K11 + K10 on the device (shape.cu): components of every region in every frame, their moments ----------------
  const int slice0 = slice_ids[0];
  ENG_RC(ensure_shape_capacity(n_runs));
  ENG_RC(launch_run_components(d_runs, n_runs, d_row_offsets, h, slice0, d_cc_parent, d_skeys[0], d_svals[0], stream));
  unsigned *sorted_keys = nullptr, *sorted_vals = nullptr;
  const int comp_bits = bits_for(n_runs);
  ENG_RC(launch_sort_pairs(d_skeys[0], d_svals[0], d_skeys[1], d_svals[1], n_runs, comp_bits, d_shist, d_total, &sorted_keys, &sorted_vals, stream));
  ENG_RC(launch_group_runs(sorted_keys, sorted_vals, n_runs, d_runs, 0, d_tile_counts, d_tile_bases, d_head_pos, d_ngroups, d_groups[0],
                           d_group_of_run, d_intervals[0], stream));
  unsigned n_comps = 0;
  ENG_CUDA(cudaMemcpyAsync(&n_comps, d_ngroups, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
  unsigned long long h_stats[8];
  ENG_CUDA(cudaMemcpyAsync(h_stats, mp.stats, sizeof(h_stats), cudaMemcpyDeviceToHost, stream));
  ENG_CUDA(cudaStreamSynchronize(stream));
  const double t_dbg1 = now_ms();
  ENG_RC(ensure_host_groups(0, n_comps, n_runs));
  ENG_CUDA(cudaMemcpyAsync(h_groups[0], d_groups[0], sizeof(RunGroup) * n_comps, cudaMemcpyDeviceToHost, stream));
  ENG_CUDA(cudaMemcpyAsync(h_intervals[0], d_intervals[0], sizeof(int3) * n_runs, cudaMemcpyDeviceToHost, stream));
  d2h_bytes += (double)sizeof(RunGroup) * n_comps + (double)sizeof(int3) * n_runs + sizeof(h_bstart) + 64;
  cudaEventRecord(ev[3], stream);
  ENG_CUDA(cudaStreamSynchronize(stream));
  stats[7] += 6 + 3 + 3 * ((comp_bits + 7) / 8) + 4;
  stats[8] += (double)h_stats[0];
  const double t_host0 = now_ms();
  if (stage_debug)
    fprintf(stderr, "[vsb200 stage] runs %u components %u: components+sort+moments %.2f ms, copy out %.2f ms\n", n_runs, n_comps,
            t_dbg1 - t_dbg0, t_host0 - t_dbg1);
  // ---------------- regions in first-seen order (ObtainResults, :533-559) ----------------
  // components arrive ascending by their first run, so the first component of a label is the label's first run
  std::vector<std::unique_ptr<Region>> regions;
  std::unordered_map<int, int> label2region;
  label2region.reserve(1 << 14);
  std::vector<int> region_of_group(n_comps);
  for (unsigned g = 0; g < n_comps; ++g) {
    const RunGroup& c = h_groups[0][g];
    auto it = label2region.find(c.tag);
    int ri;
    if (it == label2region.end()) {
      ri = (int)regions.size();
      label2region[c.tag] = ri;
      regions.emplace_back(new Region);
      regions.back()->index = ri;
      regions.back()->label = c.tag;
    } else {
      ri = it->second;
    }
    region_of_group[g] = ri;
    vsbt::Piece piece;
    piece.frame = c.slice; piece.group = (int)g;
    piece.moments.size = (float)c.area; piece.moments.mean_x = c.mean_x; piece.moments.mean_y = c.mean_y;
    piece.moments.xx = c.xx; piece.moments.xy = c.xy; piece.moments.yy = c.yy;
    piece.intervals = h_intervals[0] + c.first; piece.n_intervals = c.count;
    Region& R = *regions[ri];
    R.pieces.push_back(piece);
    if (R.frames.empty() || R.frames.back().first < c.slice) R.frames.emplace_back(c.slice, 0);
    R.frames.back().second += c.area;
  }
  const double t_dbg2 = now_ms();
  // sizes / constraints of the representatives (GetCreateRegionInformation + size_adjust_map)
  {
    const int m = (int)regions.size();
    ENG_RC(ensure_tmp_capacity((size_t)m));
    std::vector<int> ids(m);
    for (int i = 0; i < m; ++i) ids[i] = regions[i]->label;
    std::vector<int2> info(m);
    ENG_CUDA(cudaMemcpyAsync(d_tmp_ids, ids.data(), sizeof(int) * m, cudaMemcpyHostToDevice, stream));
    ENG_RC(launch_gather_region_info(d_tmp_ids, m, d_rec, d_size_adjust, (int2*)d_tmp_info, stream));
    ENG_CUDA(cudaMemcpyAsync(info.data(), d_tmp_info, sizeof(int2) * m, cudaMemcpyDeviceToHost, stream));
    ENG_CUDA(cudaStreamSynchronize(stream));
    for (int i = 0; i < m; ++i) { regions[i]->size = info[i].x; regions[i]->constrained_id = info[i].y; }
    stats[7] += 1;
  }
  const double t_dbg3 = now_ms();
  // ---------------- EnforceSpatialConnectedness (:666-904): tube decisions on the host over the components ----------------
  std::vector<int> label_of_group;                 // fresh label of every component that leaves its region; empty = no split
  if (o.enforce_spatial_connectedness) {
    std::vector<const float*> flows;
    if (use_flow) for (int s = 0; s < slots; ++s) flows.push_back(h_flows[s].empty() ? nullptr : h_flows[s].data());
    int next_label = (int)nodes;
    const int num_regions = (int)regions.size();
    // the per-region tube analysis is independent: host worker threads (the reference runs it serially)
    std::vector<std::vector<vsbt::Tube>> all_tubes(num_regions);
    {
      const int nt = std::max(1, std::min<int>(host_threads, num_regions));
      // regions with the most components first: the one that dominates must not be the last to start
      std::vector<int> by_work(num_regions);
      for (int r = 0; r < num_regions; ++r) by_work[r] = r;
      std::stable_sort(by_work.begin(), by_work.end(), [&](int a, int b) { return regions[a]->pieces.size() > regions[b]->pieces.size(); });
      std::atomic<int> next_region{0};
      auto work = [&]() {
        for (int k = next_region.fetch_add(1); k < num_regions; k = next_region.fetch_add(1)) {
          const int r = by_work[k];
          if (regions[r]->pieces.size() < 2) continue;          // one component in one frame: nothing to split
          all_tubes[r] = vsbt::TubeSplitter(regions[r]->pieces, w, h, use_flow ? &flows : nullptr).run();
        }
      };
      std::vector<std::thread> pool;
      for (int t = 1; t < nt; ++t) pool.emplace_back(work);
      work();
      for (auto& th : pool) th.join();
    }
    for (int r = 0; r < num_regions; ++r) {
      std::vector<vsbt::Tube>& tubes = all_tubes[r];
      if (tubes.empty()) continue;
      int keep = -1, keep_score = 0;                // the largest tube keeps the region's identity
      std::vector<float> areas(tubes.size());
      for (int k = 0; k < (int)tubes.size(); ++k) {
        float area = 0;
        for (const auto& s : tubes[k]) area += s.shape.size;
        areas[k] = area;
        if (area > keep_score) { keep_score = area; keep = k; }
      }
      const std::vector<vsbt::Piece> pieces = regions[r]->pieces;
      for (int k = 0; k < (int)tubes.size(); ++k) {
        Region* target = regions[r].get();
        if (k != keep) {
          regions[r]->size -= areas[k];                 // size_adjust_map[rep] -= area
          regions.emplace_back(new Region);
          target = regions.back().get();
          target->index = (int)regions.size() - 1;
          target->label = next_label++;
          target->size = areas[k];
          target->constrained_id = -1;
          label2region[target->label] = target->index;
          if (label_of_group.empty()) label_of_group.assign(n_comps, -1);
        }
        target->pieces.clear();
        target->frames.clear();
        for (const auto& s : tubes[k]) {
          target->frames.emplace_back(s.frame, s.shape.size);
          for (int pi : s.pieces) {
            target->pieces.push_back(pieces[pi]);
            if (k != keep) { label_of_group[pieces[pi].group] = target->label; region_of_group[pieces[pi].group] = target->index; }
          }
        }
      }
    }
  }
  const double t_dbg4 = now_ms();
  // ---------------- result shaping, part 1 (dense_segmentation.cpp:335-398): which regions, which ids, which order ----------------
  const int overlap_start = slots - (flush_all ? 0 : overlap_frames);
  const int last_output_frame = std::min(slots - 1, overlap_start);
  const int max_result_frame = std::min(slots - 1, last_output_frame + constraint_frames);
  // ConstrainSegmentationToFrameInterval(0, last_output_frame + 1) (segmentation.cpp:392-403)
  for (auto& R : regions)
    if (R->frames.empty() || R->frames.front().first >= last_output_frame + 1 || R->frames.back().first < 0) R->removed = true;
  // AdjustRegionAreaToFrameInterval (segmentation.cpp:424-441)
  for (auto& R : regions)
    for (const auto& fr : R->frames)
      if (fr.first < 0 || fr.first >= last_output_frame + 1) R->size -= fr.second;
  // AssignUniqueRegionIds (segmentation.cpp:549-582)
  const bool use_constraints = constrained_chunk;
  int max_id = -1;
  for (auto& R : regions) {
    R->region_id = (use_constraints && R->constrained_id >= 0) ? R->constrained_id : R->index + max_region_id;
    max_id = std::max(max_id, R->region_id);
  }
  max_region_id = std::max(max_region_id, max_id + 1);
  // order of the regions inside a frame of the result: list order, or ascending id once ids come from constraints
  const unsigned n_ranks = (unsigned)regions.size();
  std::vector<int> region_at_rank(n_ranks);
  for (unsigned k = 0; k < n_ranks; ++k) region_at_rank[k] = (int)k;
  if (use_constraints)
    std::stable_sort(region_at_rank.begin(), region_at_rank.end(),
                     [&](int a, int b) { return regions[a]->region_id < regions[b]->region_id; });
  if ((unsigned long long)ns * n_ranks >= (1ull << 32)) { set_error("%u regions x %d frames overflow the result sort key", n_ranks, ns); return VSB200_ERR_CAPACITY; }
  {
    std::vector<int> rank_of_region(n_ranks);
    for (unsigned k = 0; k < n_ranks; ++k) rank_of_region[region_at_rank[k]] = (int)k;
    std::vector<int> rank_of_group(n_comps);
    for (unsigned g = 0; g < n_comps; ++g) rank_of_group[g] = rank_of_region[region_of_group[g]];
    ENG_CUDA(cudaMemcpyAsync(d_group_tab, rank_of_group.data(), sizeof(int) * n_comps, cudaMemcpyHostToDevice, stream));
    if (!label_of_group.empty())
      ENG_CUDA(cudaMemcpyAsync(d_group_tab + n_comps, label_of_group.data(), sizeof(int) * n_comps, cudaMemcpyHostToDevice, stream));
    ENG_CUDA(cudaStreamSynchronize(stream));        // the tables are stack vectors
  }
  stats[5] += now_ms() - t_host0;
  if (stage_debug)
    fprintf(stderr, "[vsb200 stage] host: regions from components %.2f ms, sizes %.2f ms, tubes %.2f ms (%d regions, %d threads), ids + order + tables %.2f ms\n",
            t_dbg2 - t_host0, t_dbg3 - t_dbg2, t_dbg4 - t_dbg3, (int)regions.size(), host_threads, now_ms() - t_dbg4);
  const double t_dbg5 = now_ms();
  // split-off tubes get their fresh labels in the label volume (:866-893), then every run its place in the result:
  // (frame, rank of its region), stable, so that a region's intervals stay in raster order; K10 over that order gives the
  // ShapeMoments of every region in every frame and the interval arrays of the output, ready to copy
  if (!label_of_group.empty()) {
    ENG_RC(launch_relabel_groups(d_runs, n_runs, d_group_of_run, d_group_tab + n_comps, w, h, d_labels, stream));
    stats[7] += 1;
  }
  const int result_bits = bits_for((unsigned long long)ns * n_ranks);
  ENG_RC(launch_result_keys(d_runs, n_runs, d_group_of_run, d_group_tab, slice0, n_ranks, d_skeys[0], d_svals[0], stream));
  ENG_RC(launch_sort_pairs(d_skeys[0], d_svals[0], d_skeys[1], d_svals[1], n_runs, result_bits, d_shist, d_total, &sorted_keys, &sorted_vals, stream));
  ENG_RC(launch_group_runs(sorted_keys, sorted_vals, n_runs, d_runs, 1, d_tile_counts, d_tile_bases, d_head_pos, d_ngroups, d_groups[1],
                           nullptr, d_intervals[1], stream));
  unsigned n_slices_out = 0;
  ENG_CUDA(cudaMemcpyAsync(&n_slices_out, d_ngroups, sizeof(unsigned), cudaMemcpyDeviceToHost, stream));
  stats[7] += 1 + 3 * ((result_bits + 7) / 8) + 4;
  // ---------------- neighbours (DetermineNeighborIdsImpl, segmentation_graph.h:466-496) ----------------
  unsigned long long n_pairs = 0;
  for (int attempt = 0;; ++attempt) {
    ENG_RC(launch_neighbor_pairs(d_roots, d_labels, w, h, slots, use_flow ? d_flows : nullptr, constrained_chunk ? 1 : 0,
                                 d_pair_table, pair_table_cap, d_pairs, d_pair_count, pairs_cap, stream));
    ENG_CUDA(cudaMemcpyAsync(&n_pairs, d_pair_count, sizeof(n_pairs), cudaMemcpyDeviceToHost, stream));
    ENG_CUDA(cudaStreamSynchronize(stream));
    if (n_pairs <= pairs_cap && n_pairs * 2 <= pair_table_cap) break;
    // more distinct neighbour pairs than the tables hold (small min region size, large frames): the count is exact
    // as long as the hash table itself did not fill up, so size both for it (with head room) and run the pass again
    if (attempt >= 4 || pair_table_cap >= (1u << 30)) {
      set_error("neighbour pair tables cannot hold %llu pairs", n_pairs);
      return VSB200_ERR_CAPACITY;
    }
    unsigned long long want = std::max<unsigned long long>(n_pairs * 2, pairs_cap * 4);
    unsigned cap = 1u << 22;
    while ((unsigned long long)cap < want * 2 && cap < (1u << 30)) cap <<= 1;
    cudaFree(d_pair_table); cudaFree(d_pairs);
    d_pair_table = nullptr; d_pairs = nullptr;
    pair_table_cap = cap; pairs_cap = cap / 2;
    ENG_CUDA(cudaMalloc(&d_pair_table, sizeof(unsigned long long) * pair_table_cap));
    ENG_CUDA(cudaMalloc(&d_pairs, sizeof(unsigned long long) * pairs_cap));
  }
  cudaEventRecord(ev[4], stream);
  std::vector<unsigned long long> pairs(n_pairs);
  if (n_pairs) ENG_CUDA(cudaMemcpy(pairs.data(), d_pairs, sizeof(unsigned long long) * n_pairs, cudaMemcpyDeviceToHost));
  ENG_RC(ensure_host_groups(1, n_slices_out, n_runs));
  ENG_CUDA(cudaMemcpyAsync(h_groups[1], d_groups[1], sizeof(RunGroup) * n_slices_out, cudaMemcpyDeviceToHost, stream));
  ENG_CUDA(cudaMemcpyAsync(h_intervals[1], d_intervals[1], sizeof(int3) * n_runs, cudaMemcpyDeviceToHost, stream));
  ENG_CUDA(cudaStreamSynchronize(stream));
  d2h_bytes += 8.0 * n_pairs + 8.0 * regions.size() + (double)sizeof(RunGroup) * n_slices_out + (double)sizeof(int3) * n_runs;
  stats[7] += 3;
  const double t_host1 = now_ms();
  if (stage_debug) fprintf(stderr, "[vsb200 stage] result order + neighbour pairs on the device, copies out: %.2f ms (%llu pairs)\n", t_host1 - t_dbg5, n_pairs);
  for (unsigned long long key : pairs) {
    const int la = (int)(key >> 32), lb = (int)(key & 0xffffffffu);
    auto ia = label2region.find(la), ib = label2region.find(lb);
    if (ia == label2region.end() || ib == label2region.end()) continue;   // virtual-only representatives: never output
    regions[ia->second]->neighbors.push_back(ib->second);
    regions[ib->second]->neighbors.push_back(ia->second);
  }
  for (auto& R : regions) {
    std::sort(R->neighbors.begin(), R->neighbors.end());
    R->neighbors.erase(std::unique(R->neighbors.begin(), R->neighbors.end()), R->neighbors.end());
  }
  // ---------------- result shaping, part 2: the frames of the result ----------------
  const int chunk_sz = last_output_frame - curr_chunk_start + 1;
  const int hierarchy_frame_idx = num_output_frames;
  const int n_out_frames = max_result_frame - curr_chunk_start + 1;
  std::vector<std::unique_ptr<FrameOut>> frame_outs(std::max(n_out_frames, 0));
  std::vector<std::pair<unsigned, unsigned>> groups_of_slice(ns, std::make_pair(0u, 0u));
  for (unsigned k = 0; k < n_slices_out;) {
    unsigned e = k;
    while (e < n_slices_out && h_groups[1][e].slice == h_groups[1][k].slice) ++e;
    groups_of_slice[h_groups[1][k].slice - slice0] = std::make_pair(k, e);
    k = e;
  }
  auto build_frame = [&](int f) {
    // RetrieveSegmentation3D (segmentation.cpp:458-533)
    std::unique_ptr<FrameOut> out(new FrameOut);
    out->width = w; out->height = h; out->chunk_id = chunk_id;
    out->connectedness = o.enforce_n4_connectivity ? 1 : 2;
    out->chunk_size = chunk_sz; out->overlap_start = chunk_sz; out->hierarchy_frame_idx = hierarchy_frame_idx;
    out->pts = 0;
    // the groups of pass 1 with this frame: one per region present, in result order, intervals contiguous
    const std::pair<unsigned, unsigned> span = groups_of_slice[f - slice0];
    const size_t n_items = span.second - span.first;
    out->region_id.reserve(n_items);
    out->interval_offset.reserve(n_items + 1);
    out->moments.reserve(n_items * 6);
    out->interval_offset.push_back(0);
    if (n_items) {
      const RunGroup& g0 = h_groups[1][span.first];
      const RunGroup& g1 = h_groups[1][span.second - 1];
      const size_t n_iv = (size_t)(g1.first + g1.count - g0.first);
      out->intervals.resize(n_iv * 3);
      memcpy(out->intervals.data(), h_intervals[1] + g0.first, n_iv * sizeof(vsbs::Interval));
      for (unsigned k = span.first; k < span.second; ++k) {
        const RunGroup& g = h_groups[1][k];
        const unsigned rank = (unsigned)g.tag - (unsigned)(f - slice0) * n_ranks;
        out->region_id.push_back(regions[region_at_rank[rank]]->region_id);
        out->interval_offset.push_back((int32_t)(g.first + g.count - g0.first));
        const float mm[6] = {(float)g.area, g.mean_x, g.mean_y, g.xx, g.xy, g.yy};
        out->moments.insert(out->moments.end(), mm, mm + 6);
      }
    }
    out->neighbor_offset.push_back(0);
    if (f == curr_chunk_start) {                         // hierarchy level 0 (segmentation.cpp:702-773)
      struct Comp { int id, size, sf, ef; std::vector<int> nb; };
      std::vector<Comp> comps;
      for (const auto& R : regions) {
        if (R->removed) continue;
        Comp c;
        c.id = R->region_id; c.size = R->size;
        c.sf = R->frames.front().first; c.ef = R->frames.back().first;
        for (int nb : R->neighbors) if (!regions[nb]->removed) c.nb.push_back(regions[nb]->region_id);
        if (use_constraints) std::sort(c.nb.begin(), c.nb.end());
        comps.push_back(std::move(c));
      }
      if (use_constraints) std::sort(comps.begin(), comps.end(), [](const Comp& a, const Comp& b) { return a.id < b.id; });
      for (const Comp& c : comps) {
        out->compound.push_back(c.id); out->compound.push_back(c.size);
        out->compound.push_back(c.sf); out->compound.push_back(c.ef);
        out->neighbor_id.insert(out->neighbor_id.end(), c.nb.begin(), c.nb.end());
        out->neighbor_offset.push_back((int32_t)out->neighbor_id.size());
      }
    }
    if (o.want_id_maps) {
      out->id_map.assign((size_t)n, -1);
      for (size_t k = 0; k < out->region_id.size(); ++k)
        for (int q = out->interval_offset[k]; q < out->interval_offset[k + 1]; ++q) {
          const int y = out->intervals[3 * q], lx = out->intervals[3 * q + 1], rx = out->intervals[3 * q + 2];
          std::fill(out->id_map.begin() + (size_t)y * w + lx, out->id_map.begin() + (size_t)y * w + rx + 1, out->region_id[k]);
        }
    }
    frame_outs[f - curr_chunk_start] = std::move(out);
  };
  {
    const int nt = std::max(1, std::min(host_threads, n_out_frames));
    std::atomic<int> next_frame{curr_chunk_start};
    auto work = [&]() { for (int f = next_frame.fetch_add(1); f <= max_result_frame; f = next_frame.fetch_add(1)) build_frame(f); };
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& th : pool) th.join();
  }
  for (int f = curr_chunk_start; f <= max_result_frame; ++f) {
    std::unique_ptr<FrameOut> out = std::move(frame_outs[f - curr_chunk_start]);
    if (f <= last_output_frame) {
      if (f < last_output_frame) {
        results->push_back(std::move(out));
      } else {
        results->push_back(std::unique_ptr<FrameOut>(new FrameOut(*out)));
      }
      ++num_output_frames;
    }
    if (f >= last_output_frame && out) overlap_out.push_back(std::move(out));
  }
  // feature_buffer_.erase(begin, begin + last_output_frame): the kept frames become slots 0, 1
  if (!flush_all) {
    // slot 1 of the next chunk = frame `last_output_frame + 1`; slot 0 (virtual) needs no pixels
    std::swap(d_frames[1], d_frames[last_output_frame + 1]);
    if (use_flow) {
      ENG_CUDA(cudaMemcpyAsync(d_flows + (size_t)1 * n * 2, d_flows + (size_t)(last_output_frame + 1) * n * 2,
                               (size_t)n * 2 * sizeof(float), cudaMemcpyDeviceToDevice, stream));
      std::swap(h_flows[1], h_flows[last_output_frame + 1]);
      h_flows[0].clear();
    }
    buffered = overlap_frames;
    curr_chunk_start = 1;
  } else {
    buffered = 0;
    curr_chunk_start = 0;
    overlap_out.clear();
  }
  ++chunk_id;
  stats[5] += now_ms() - t_host1;
  if (stage_debug) fprintf(stderr, "[vsb200 stage] host: neighbour lists + frames of the result %.2f ms\n", now_ms() - t_host1);
  float ms;
  cudaEventElapsedTime(&ms, ev[0], ev[1]); stats[2] += ms;
  cudaEventElapsedTime(&ms, ev[1], ev[2]); stats[3] += ms;
  cudaEventElapsedTime(&ms, ev[2], ev[3]); stats[4] += ms;
  cudaEventElapsedTime(&ms, ev[3], ev[4]); stats[6] += ms;
  return 0;
}

// ---------------------------------------------------------------------------
// proto2 wire encoding of segmentation.SegmentationDesc (segment_util/segmentation.proto:55-172)
// ---------------------------------------------------------------------------
namespace {
struct PB {
  std::vector<uint8_t>& b;
  void varint(uint64_t v) { while (v >= 0x80) { b.push_back((uint8_t)(v | 0x80)); v >>= 7; } b.push_back((uint8_t)v); }
  void tag(int field, int wire) { varint((uint64_t)(field << 3 | wire)); }
  void i32(int field, int32_t v) { tag(field, 0); varint((uint64_t)(int64_t)v); }      // int32: sign-extended varint
  void f32(int field, float v) { tag(field, 5); uint32_t u; memcpy(&u, &v, 4); for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(u >> (8 * i))); }
  void bytes(int field, const std::vector<uint8_t>& s) { tag(field, 2); varint(s.size()); b.insert(b.end(), s.begin(), s.end()); }
};

void encode_proto(const FrameOut& f, std::vector<uint8_t>* out) {
  out->clear();
  PB top{*out};
  std::vector<uint8_t> reg, ras, si, sm;
  for (size_t k = 0; k < f.region_id.size(); ++k) {        // repeated Region2D region = 2
    reg.clear(); ras.clear();
    PB r{reg};
    r.i32(1, f.region_id[k]);                               // required int32 id = 1
    PB rs{ras};
    for (int q = f.interval_offset[k]; q < f.interval_offset[k + 1]; ++q) {
      si.clear();
      PB s{si};
      s.i32(1, f.intervals[3 * q]); s.i32(2, f.intervals[3 * q + 1]); s.i32(3, f.intervals[3 * q + 2]);
      rs.bytes(1, si);                                      // repeated ScanInterval scan_inter = 1
    }
    r.bytes(3, ras);                                        // optional Rasterization raster = 3
    sm.clear();
    PB m{sm};
    for (int j = 0; j < 6; ++j) m.f32(j + 1, f.moments[6 * k + j]);
    r.bytes(5, sm);                                         // optional ShapeMoments shape_moments = 5
    top.bytes(2, reg);
  }
  if (!f.compound.empty()) {                                // repeated HierarchyLevel hierarchy = 3
    std::vector<uint8_t> hier, comp;
    PB hl{hier};
    const size_t nc = f.compound.size() / 4;
    for (size_t c = 0; c < nc; ++c) {
      comp.clear();
      PB cr{comp};
      cr.i32(1, f.compound[4 * c]); cr.i32(2, f.compound[4 * c + 1]);
      for (int q = f.neighbor_offset[c]; q < f.neighbor_offset[c + 1]; ++q) cr.i32(3, f.neighbor_id[q]);
      cr.i32(6, f.compound[4 * c + 2]); cr.i32(7, f.compound[4 * c + 3]);
      hl.bytes(2, comp);                                    // repeated CompoundRegion region = 2
    }
    top.bytes(3, hier);
  }
  top.i32(4, f.width); top.i32(5, f.height); top.i32(6, f.chunk_size); top.i32(7, f.overlap_start);
  top.i32(8, f.chunk_id); top.i32(9, f.hierarchy_frame_idx); top.i32(12, f.connectedness);
}
}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

void vsb200_dense_default_opts(vsb200_dense_opts* o) {
  if (!o) return;
  o->presmoothing = 2; o->frac_min_region_size = 0.01f; o->chunk_size = 20; o->chunk_overlap_ratio = 0.2f;
  o->num_constraint_frames = 1; o->two_stage_oversegment = 0; o->thin_structure_suppression = 0;
  o->enforce_n4_connectivity = 1; o->enforce_spatial_connectedness = 1; o->color_distance = 1;
  o->compute_vectorization = 0; o->device = 0; o->want_id_maps = 0;
}

int vsb200_dense_create(const vsb200_dense_opts* o, int width, int height, int use_flow, vsb200_dense** out) {
  if (!o || !out || width < 2 || height < 2) { set_error("create: bad arguments"); return VSB200_ERR_INVALID; }
  if (o->chunk_size < 3) { set_error("Chunk size needs to be at least 3 frames."); return VSB200_ERR_INVALID; }   // dense_segmentation.cpp:54
  int overlap = (int)(o->chunk_overlap_ratio * o->chunk_size + 0.5f);
  overlap = std::min(overlap, 2);                                             // dense_segmentation.cpp:59-62
  if (overlap >= o->chunk_size || overlap < 2) { set_error("Overlap needs to be 2 frames and smaller than chunk_size."); return VSB200_ERR_INVALID; }
  if (o->num_constraint_frames < 1) { set_error("num_constraint_frames must be >= 1"); return VSB200_ERR_INVALID; }
  if (o->presmoothing == 1) { set_error("PRESMOOTH_GAUSSIAN (cv::GaussianBlur) is not built"); return VSB200_ERR_UNSUPPORTED; }
  if (o->two_stage_oversegment || o->thin_structure_suppression || o->compute_vectorization) {
    set_error("two_stage_oversegment / thin_structure_suppression / compute_vectorization are not built");
    return VSB200_ERR_UNSUPPORTED;
  }
  if (vsb200_device_count() <= 0) { set_error("no sm_100 CUDA device available: this path has no CPU fallback"); return VSB200_ERR_NO_DEVICE; }
  std::unique_ptr<vsb200_dense> d(new vsb200_dense);
  d->o = *o; d->w = width; d->h = height; d->use_flow = use_flow != 0; d->l1 = (o->color_distance == 0);
  d->overlap_frames = overlap;
  {
    const unsigned hc = std::thread::hardware_concurrency();
    d->host_threads = (int)std::min(16u, std::max(1u, hc));
    if (const char* e = getenv("VSB200_HOST_THREADS")) d->host_threads = std::max(1, atoi(e));
  }
  d->constraint_frames = std::min(o->num_constraint_frames, overlap - 1);
  if (int rc = d->init()) return rc;
  *out = d.release();
  return VSB200_OK;
}

int vsb200_dense_push(vsb200_dense* d, const uint8_t* bgr, int row_stride_bytes, const float* flow_xy,
                      int flow_row_stride_bytes, int64_t pts, int* n_ready) {
  if (!d) return VSB200_ERR_INVALID;
  return d->push(bgr, row_stride_bytes, flow_xy, flow_row_stride_bytes, pts, n_ready, false);
}

int vsb200_dense_push_device(vsb200_dense* d, const uint8_t* dev_bgr, int row_stride_bytes, int64_t pts, int* n_ready) {
  if (!d) return VSB200_ERR_INVALID;
  if (d->use_flow) { set_error("push_device: flow streams use the host entry point"); return VSB200_ERR_UNSUPPORTED; }
  return d->push(dev_bgr, row_stride_bytes, nullptr, 0, pts, n_ready, true);
}

void vsb200_dense_set_profiling(vsb200_dense* d, int time_edge_kernel) { if (d) d->time_edges = time_edge_kernel; }

void vsb200_dense_io_stats(vsb200_dense* d, double out[4]) {
  if (!d || !out) return;
  out[0] = d->h2d_bytes; out[1] = d->d2h_bytes; out[2] = d->edge_ms; out[3] = d->edge_launches;
}

int vsb200_dense_flush(vsb200_dense* d, int* n_ready) {
  if (!d) return VSB200_ERR_INVALID;
  return d->flush(n_ready);
}

int vsb200_dense_pop(vsb200_dense* d, vsb200_frame_result* out) {
  if (!d || !out) return VSB200_ERR_INVALID;
  if (d->ready.empty()) return VSB200_ERR_EMPTY;
  d->last_popped = std::move(d->ready.front());
  d->ready.pop_front();
  const FrameOut& f = *d->last_popped;
  out->width = f.width; out->height = f.height; out->chunk_id = f.chunk_id; out->chunk_size = f.chunk_size;
  out->overlap_start = f.overlap_start; out->hierarchy_frame_idx = f.hierarchy_frame_idx;
  out->connectedness = f.connectedness;
  out->n_regions = (int32_t)f.region_id.size();
  out->region_id = f.region_id.data(); out->interval_offset = f.interval_offset.data();
  out->intervals = f.intervals.data(); out->shape_moments = f.moments.data();
  out->n_compound = (int32_t)(f.compound.size() / 4);
  out->compound = f.compound.data(); out->neighbor_offset = f.neighbor_offset.data();
  out->neighbor_id = f.neighbor_id.data();
  out->pts = f.pts;
  return VSB200_OK;
}

const int32_t* vsb200_dense_last_id_map(vsb200_dense* d) {
  if (!d || !d->last_popped || d->last_popped->id_map.empty()) return nullptr;
  return d->last_popped->id_map.data();
}

size_t vsb200_dense_last_proto(vsb200_dense* d, uint8_t* buf, size_t cap) {
  if (!d || !d->last_popped) return 0;
  encode_proto(*d->last_popped, &d->proto_buf);
  if (buf && cap) memcpy(buf, d->proto_buf.data(), std::min(cap, d->proto_buf.size()));
  return d->proto_buf.size();
}

// Host only: the same encoder over caller-held arrays (a frame result kept after later pops, or one rebuilt from a
// file).  Bytes are protobuf's canonical serialisation of the message (fields in field-number order, unpacked
// repeated int32), i.e. what SegmentationDesc::SerializeToString writes (segmentation_io.cpp:73-78).
size_t vsb200_encode_frame_proto(const vsb200_frame_result* r, uint8_t* buf, size_t cap) {
  if (!r || r->n_regions < 0 || r->n_compound < 0) return 0;
  FrameOut f;
  f.width = r->width; f.height = r->height; f.chunk_id = r->chunk_id; f.chunk_size = r->chunk_size;
  f.overlap_start = r->overlap_start; f.hierarchy_frame_idx = r->hierarchy_frame_idx; f.connectedness = r->connectedness;
  f.pts = r->pts;
  const int nr = r->n_regions, nc = r->n_compound;
  if (nr > 0) {
    f.region_id.assign(r->region_id, r->region_id + nr);
    f.interval_offset.assign(r->interval_offset, r->interval_offset + nr + 1);
    f.intervals.assign(r->intervals, r->intervals + 3 * (size_t)r->interval_offset[nr]);
    f.moments.assign(r->shape_moments, r->shape_moments + 6 * (size_t)nr);
  }
  if (nc > 0) {
    f.compound.assign(r->compound, r->compound + 4 * (size_t)nc);
    f.neighbor_offset.assign(r->neighbor_offset, r->neighbor_offset + nc + 1);
    f.neighbor_id.assign(r->neighbor_id, r->neighbor_id + r->neighbor_offset[nc]);
  }
  std::vector<uint8_t> out;
  encode_proto(f, &out);
  if (buf && cap) memcpy(buf, out.data(), std::min(cap, out.size()));
  return out.size();
}

void vsb200_dense_stats(vsb200_dense* d, double out[9]) {
  if (d && out) memcpy(out, d->stats, sizeof(double) * 9);
}

void vsb200_dense_destroy(vsb200_dense* d) { delete d; }

int vsb200_dense_export_halo(vsb200_dense* d, int32_t* dev_prev_out, int32_t* dev_last_out, int32_t chain_state[3]) {
  if (!d || !dev_prev_out || !dev_last_out || !chain_state) return VSB200_ERR_INVALID;
  if (d->chunk_id == 0 || d->buffered != d->overlap_frames) {
    set_error("export_halo: only valid right after a chunk boundary");
    return VSB200_ERR_INVALID;
  }
  // the id maps that seed the next chunk (overlap_segmentations_, dense_segmentation.cpp:300-315)
  if (cudaMemcpyAsync(dev_prev_out, d->d_con_ids[0], (size_t)d->n * 4, cudaMemcpyDeviceToDevice, d->stream) != cudaSuccess ||
      cudaMemcpyAsync(dev_last_out, d->d_con_ids[1], (size_t)d->n * 4, cudaMemcpyDeviceToDevice, d->stream) != cudaSuccess ||
      cudaStreamSynchronize(d->stream) != cudaSuccess) {
    set_error("export_halo: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return VSB200_ERR_CUDA;
  }
  chain_state[0] = d->max_region_id;         // max_region_id_ (dense_segmentation.cpp:360-365)
  chain_state[1] = d->chunk_id;              // id of the chunk the maps constrain
  chain_state[2] = d->num_output_frames;     // frames output so far (hierarchy_frame_idx of the next chunk)
  return VSB200_OK;
}

// Successor side of the seam: puts a fresh engine into the state its predecessor is in right after
// a chunk boundary -- constraint maps of the virtual slot and of slot 1, region-id counter, chunk
// and frame counters.  The next push must be the frame of the second map (the predecessor's last
// pushed frame); from there the chain continues exactly as one engine would have run it.
int vsb200_dense_import_halo(vsb200_dense* d, const int32_t* dev_prev, const int32_t* dev_last, const int32_t chain_state[3]) {
  if (!d || !dev_prev || !dev_last || !chain_state) return VSB200_ERR_INVALID;
  if (d->input_frames != 0 || d->import_pending) { set_error("import_halo must precede the first push"); return VSB200_ERR_INVALID; }
  if (chain_state[0] < 0 || chain_state[1] < 1 || chain_state[2] < 0) { set_error("import_halo: bad chain state"); return VSB200_ERR_INVALID; }
  if (cudaSetDevice(d->o.device) != cudaSuccess ||
      cudaMemcpyAsync(d->d_con_ids[0], dev_prev, (size_t)d->n * 4, cudaMemcpyDeviceToDevice, d->stream) != cudaSuccess ||
      cudaMemcpyAsync(d->d_con_ids[1], dev_last, (size_t)d->n * 4, cudaMemcpyDeviceToDevice, d->stream) != cudaSuccess ||
      cudaStreamSynchronize(d->stream) != cudaSuccess) {
    set_error("import_halo: copy failed: %s", cudaGetErrorString(cudaGetLastError()));
    return VSB200_ERR_CUDA;
  }
  d->max_region_id = chain_state[0];
  d->chunk_id = chain_state[1];
  d->num_output_frames = chain_state[2];
  d->buffered = 1;                 // slot 0 = virtual copy of the predecessor's last output frame
  d->curr_chunk_start = 1;
  d->import_pending = true;
  return VSB200_OK;
}

}  // extern "C"
