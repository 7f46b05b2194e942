// capi_kernels.cu -- kernel-level C-ABI entry points (include/vsb200.h) and the one-shot
// whole-chunk segmentation used by the merge parity tests and bench.py.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <atomic>
#include <unistd.h>

#include "../../include/vsb200.h"
#include "common.cuh"
#include "results.cuh"
#include "shape.cuh"

using namespace vsb;

extern "C" {

const char* vsb200_last_error(void) { return vsb::last_error(); }

int vsb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  int ok = 0;
  for (int d = 0; d < n; ++d) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, d) == cudaSuccess && prop.major == 10) ++ok;
  }
  return ok;
}

static int require_device() {
  static int cached = -1;                       // device enumeration is slow (cudaGetDeviceProperties)
  if (cached < 0) cached = vsb200_device_count();
  if (cached <= 0) {
    set_error("no sm_100 CUDA device available: this path has no CPU fallback");
    return VSB200_ERR_NO_DEVICE;
  }
  return 0;
}

size_t vsb200_preprocess_scratch_bytes(void) { return preprocess_scratch_bytes(); }

int vsb200_preprocess(const uint8_t* dev_bgr, int row_stride_bytes, int width, int height, int presmoothing,
                      float* dev_out, void* dev_scratch, void* stream) {
  if (int rc = require_device()) return rc;
  if (!dev_bgr || !dev_out || !dev_scratch || width < 2 || height < 2 || row_stride_bytes < width * 3) {
    set_error("vsb200_preprocess: bad arguments");
    return VSB200_ERR_INVALID;
  }
  return launch_preprocess(dev_bgr, row_stride_bytes, width, height, presmoothing, dev_out, dev_scratch,
                           (cudaStream_t)stream);
}

int vsb200_edge_build(const float* dev_curr, const float* dev_prev, const float* dev_flow, int width, int height,
                      int l1, float* dev_spatial_out, float* dev_temporal_out, void* stream) {
  if (int rc = require_device()) return rc;
  return launch_edge_build(dev_curr, dev_prev, dev_flow, width, height, l1 != 0, dev_spatial_out, dev_temporal_out,
                           (cudaStream_t)stream);
}

int vsb200_bucket_index(float weight) { return bucket_of(weight); }

size_t vsb200_sort_scratch_bytes(int num_lists, int width, int height) {
  return sort_scratch_bytes(num_lists, width, height);
}

int vsb200_sort_edges(const float* const* host_seg_ptrs, int num_lists, int width, int height,
                      uint32_t* dev_codes_out, uint64_t* dev_bucket_start_out, void* dev_scratch,
                      size_t scratch_bytes, void* stream) {
  if (int rc = require_device()) return rc;
  return launch_sort_edges(host_seg_ptrs, num_lists, width, height, dev_codes_out,
                           (unsigned long long*)dev_bucket_start_out, dev_scratch, scratch_bytes,
                           (cudaStream_t)stream);
}

// One chunk, no constraints: frames -> node labels.  Allocates and frees its own workspace
// (a convenience entry point for tests / benchmarks; the streaming engine keeps its workspace).
int vsb200_segment_chunk(const float* dev_frames, int width, int height, int slots, int l1, int min_region_size,
                         int32_t* dev_labels_out, double* stats4, void* stream) {
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const size_t n = (size_t)width * height, nodes = n * slots;
  const int num_lists = 2 * slots - 1;
  std::vector<void*> allocs;
  auto dalloc = [&](size_t bytes) -> void* {
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    allocs.push_back(p);
    return p;
  };
  auto cleanup = [&]() { for (void* p : allocs) cudaFree(p); };
  int rc = 0;
  cudaEvent_t ev[4];
  for (auto& e : ev) cudaEventCreate(&e);
  do {
    std::vector<const float*> seg(num_lists, nullptr);
    size_t total_elems = 0;
    bool oom = false;
    for (int q = 0; q < num_lists; ++q) {
      const size_t e = n * ((q & 1) ? 9 : 4);
      float* p = (float*)dalloc(e * sizeof(float));
      if (!p) { oom = true; break; }
      seg[q] = p;
      total_elems += e;
    }
    if (oom) { set_error("segment_chunk: out of device memory"); rc = VSB200_ERR_CUDA; break; }
    uint32_t* codes = (uint32_t*)dalloc(total_elems * sizeof(uint32_t));
    unsigned long long* bstart = (unsigned long long*)dalloc(sizeof(unsigned long long) * (kNumBuckets + 1));
    const size_t sort_sc = sort_scratch_bytes(num_lists, width, height);
    void* sort_scratch = dalloc(sort_sc);
    int* parent = (int*)dalloc(nodes * sizeof(int));
    RegionRec* rec = (RegionRec*)dalloc(nodes * sizeof(RegionRec));
    if (!codes || !bstart || !sort_scratch || !parent || !rec) { set_error("segment_chunk: out of device memory"); rc = VSB200_ERR_CUDA; break; }
    cudaEventRecord(ev[0], s);
    for (int k = 0; k < slots && rc == 0; ++k) {
      const float* cur = dev_frames + (size_t)k * n * 3;
      rc = launch_init_nodes(cur, nullptr, k, width, height, parent, rec, s);
      if (rc) break;
      rc = launch_edge_build(cur, k ? cur - n * 3 : nullptr, nullptr, width, height, l1 != 0,
                             (float*)seg[2 * k], k ? (float*)seg[2 * k - 1] : nullptr, s);
    }
    if (rc) break;
    cudaEventRecord(ev[1], s);
    rc = launch_sort_edges(seg.data(), num_lists, width, height, codes, bstart, sort_scratch, sort_sc, s);
    if (rc) break;
    cudaEventRecord(ev[2], s);
    unsigned long long h_bstart[kNumBuckets + 1];
    if (cudaMemcpyAsync(h_bstart, bstart, sizeof(h_bstart), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      set_error("segment_chunk: sort failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = VSB200_ERR_CUDA;
      break;
    }
    unsigned long long max_bucket = 1;
    for (int b = 0; b < kNumBuckets; ++b) max_bucket = std::max(max_bucket, h_bstart[b + 1] - h_bstart[b]);
    // the weight lists are no longer needed: free them before the merge workspace is allocated
    for (int q = 0; q < num_lists; ++q) {
      cudaFree((void*)seg[q]);
      allocs.erase(std::find(allocs.begin(), allocs.end(), (void*)seg[q]));
    }
    MergeParams mp;
    mp.w = width; mp.h = height; mp.slots = slots; mp.min_region_size = min_region_size;
    mp.force_merge_weight = l1 ? 0.002f : 0.001f;
    mp.has_constraints = 0;
    mp.flows = nullptr;
    mp.codes = codes;
    mp.bucket_start = bstart;
    mp.parent = parent;
    mp.rec = rec;
    mp.res = (unsigned long long*)dalloc(nodes * 8);
    mp.acc = (unsigned long long*)dalloc(nodes * 32);
    mp.cl = (int*)dalloc(nodes * 4);
    mp.hull = (NodeScratch*)dalloc(nodes * sizeof(NodeScratch));
    mp.live_a = (uint32_t*)dalloc(max_bucket * 16);
    mp.live_b = (uint32_t*)dalloc(max_bucket * 16);
    mp.live_c = (uint32_t*)dalloc(max_bucket * 16);
    mp.done = (unsigned char*)dalloc(max_bucket);
    mp.live_cap = max_bucket;
    mp.counters = (unsigned long long*)dalloc((16 + 4096) * 8);
    mp.scan_queue = (uint32_t*)dalloc(kScanQueueWords * sizeof(uint32_t));
    mp.stats = mp.counters ? mp.counters + 8 : nullptr;
    mp.debug = getenv("VSB200_MERGE_DEBUG") ? (unsigned long long*)dalloc((kNumBuckets * 4 + 64) * 8) : nullptr;
    if (mp.debug) cudaMemsetAsync(mp.debug, 0, (kNumBuckets * 4 + 64) * 8, s);
    mp.trace = nullptr;
    unsigned long long* h_trace = nullptr;
    std::atomic<bool> trace_stop{false};
    std::thread trace_thread;
    if (getenv("VSB200_MERGE_TRACE")) {
      cudaHostAlloc(&h_trace, 64 * 8, cudaHostAllocMapped);
      memset(h_trace, 0, 64 * 8);
      cudaHostGetDevicePointer((void**)&mp.trace, h_trace, 0);
      std::string path = getenv("VSB200_MERGE_TRACE");
      trace_thread = std::thread([h_trace, path, &trace_stop]() {
        while (!trace_stop.load()) {
          FILE* f = fopen(path.c_str(), "a");
          if (f) { fprintf(f, "bucket %llu guard %llu n_live %llu stage %llu serial_rounds %llu wn %llu cursor %llu\n", h_trace[0], h_trace[1], h_trace[2], h_trace[3], h_trace[4], h_trace[5], h_trace[6]); fclose(f); }
          usleep(500000);
        }
      });
    }
    if (!mp.res || !mp.acc || !mp.cl || !mp.hull || !mp.live_a || !mp.live_b || !mp.live_c || !mp.done || !mp.counters || !mp.scan_queue) {
      set_error("segment_chunk: out of device memory (merge workspace)");
      rc = VSB200_ERR_CUDA;
      break;
    }
    cudaMemsetAsync(mp.res, 0xff, nodes * 8, s);
    cudaMemsetAsync(mp.acc, 0, nodes * 32, s);
    cudaMemsetAsync(mp.counters, 0, (16 + 4096) * 8, s);
    if ((rc = launch_init_iota(mp.cl, (long long)nodes, s))) break;
    if ((rc = launch_init_hull(mp.hull, (long long)nodes, s))) break;
    rc = launch_merge(mp, s);
    if (h_trace) { cudaStreamSynchronize(s); trace_stop.store(true); trace_thread.join(); cudaFreeHost(h_trace); }
    if (rc) break;
    cudaEventRecord(ev[3], s);
    if ((rc = launch_flatten(parent, nullptr, dev_labels_out, (long long)nodes, s))) break;
    unsigned long long h_stats[8];
    if (cudaMemcpyAsync(h_stats, mp.stats, sizeof(h_stats), cudaMemcpyDeviceToHost, s) != cudaSuccess ||
        cudaStreamSynchronize(s) != cudaSuccess) {
      set_error("segment_chunk: merge failed: %s", cudaGetErrorString(cudaGetLastError()));
      rc = VSB200_ERR_CUDA;
      break;
    }
    if (mp.debug) {
      std::vector<unsigned long long> dbg(kNumBuckets * 4 + 64);
      cudaMemcpy(dbg.data(), mp.debug, dbg.size() * 8, cudaMemcpyDeviceToHost);
      FILE* f = fopen(getenv("VSB200_MERGE_DEBUG"), "w");
      if (f) {
        unsigned long long prev = 0;
        for (int b = 0; b < kNumBuckets; ++b) {
          if (dbg[b * 4 + 2] == 0) continue;
          fprintf(f, "%d edges %llu pending %llu rounds %llu us %.1f\n", b, dbg[b * 4 + 2], dbg[b * 4 + 3],
                  dbg[b * 4 + 1] - prev, dbg[b * 4 + 0] / 1000.0);
          prev = dbg[b * 4 + 1];
        }
        fprintf(f, "serial phase cycles: refill %llu A %llu B12 %llu prefetch %llu B3 %llu relaxed %llu Ccompact %llu\n", dbg[kNumBuckets * 4], dbg[kNumBuckets * 4 + 1], dbg[kNumBuckets * 4 + 2], dbg[kNumBuckets * 4 + 3], dbg[kNumBuckets * 4 + 4], dbg[kNumBuckets * 4 + 5], dbg[kNumBuckets * 4 + 6]);
        fclose(f);
      }
    }
    if (stats4) {
      float ms;
      cudaEventElapsedTime(&ms, ev[0], ev[1]); stats4[0] = ms;   // init + edge build
      cudaEventElapsedTime(&ms, ev[1], ev[2]); stats4[1] = ms;   // sort
      cudaEventElapsedTime(&ms, ev[2], ev[3]); stats4[2] = ms;   // merge (incl. workspace init)
      stats4[3] = (double)h_stats[0];                            // merge rounds
    }
  } while (0);
  for (auto& e : ev) cudaEventDestroy(e);
  cleanup();
  return rc;
}


// K11 + K10 of csrc/shape.cu on a label volume (kernel-level entry for the parity tests): runs -> N4 components of every
// label in every frame -> one record per component, components numbered in the order of their first scan interval.
int vsb200_label_components(const int32_t* dev_labels, int width, int height, int slices, int32_t* dev_component_out,
                            int32_t* host_records_out, int record_cap, int* n_components_out, void* stream) {
  if (!dev_labels || width <= 0 || height <= 0 || slices <= 0 || !n_components_out) { set_error("label_components: bad arguments"); return VSB200_ERR_INVALID; }
  if (int rc = require_device()) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  const int rows = slices * height;
  int* d_slice_ids = nullptr; unsigned *d_counts = nullptr, *d_offsets = nullptr, *d_total = nullptr, *d_ngroups = nullptr;
  RunRec* d_runs = nullptr; int *d_parent = nullptr, *d_group_of_run = nullptr, *d_iota = nullptr;
  unsigned *d_keys[2] = {nullptr, nullptr}, *d_vals[2] = {nullptr, nullptr}, *d_hist = nullptr, *d_tc = nullptr, *d_tb = nullptr, *d_hp = nullptr;
  RunGroup* d_groups = nullptr; int3* d_iv = nullptr;
  int rc = 0;
  auto cleanup = [&]() {
    for (void* p : {(void*)d_slice_ids, (void*)d_counts, (void*)d_offsets, (void*)d_total, (void*)d_ngroups, (void*)d_runs, (void*)d_parent,
                    (void*)d_group_of_run, (void*)d_iota, (void*)d_keys[0], (void*)d_keys[1], (void*)d_vals[0], (void*)d_vals[1], (void*)d_hist,
                    (void*)d_tc, (void*)d_tb, (void*)d_hp, (void*)d_groups, (void*)d_iv})
      if (p) cudaFree(p);
  };
#define LC_CUDA(expr) do { if ((expr) != cudaSuccess) { set_error("label_components: %s failed: %s", #expr, cudaGetErrorString(cudaGetLastError())); rc = VSB200_ERR_CUDA; goto done; } } while (0)
#define LC_RC(expr) do { if ((rc = (expr))) goto done; } while (0)
  {
    std::vector<int> ids(slices);
    for (int k = 0; k < slices; ++k) ids[k] = k;
    unsigned n_runs = 0, n_groups = 0;
    LC_CUDA(cudaMalloc(&d_slice_ids, sizeof(int) * slices));
    LC_CUDA(cudaMalloc(&d_counts, sizeof(unsigned) * rows));
    LC_CUDA(cudaMalloc(&d_offsets, sizeof(unsigned) * rows));
    LC_CUDA(cudaMalloc(&d_total, sizeof(unsigned)));
    LC_CUDA(cudaMalloc(&d_ngroups, sizeof(unsigned)));
    LC_CUDA(cudaMemcpyAsync(d_slice_ids, ids.data(), sizeof(int) * slices, cudaMemcpyHostToDevice, s));
    LC_RC(launch_rle_count(dev_labels, width, height, d_slice_ids, slices, d_counts, s));
    LC_RC(launch_scan_u32(d_counts, d_offsets, d_total, rows, s));
    LC_CUDA(cudaMemcpyAsync(&n_runs, d_total, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    LC_CUDA(cudaStreamSynchronize(s));
    const size_t cap = (size_t)n_runs + 16;
    LC_CUDA(cudaMalloc(&d_runs, cap * sizeof(RunRec)));
    LC_CUDA(cudaMalloc(&d_parent, cap * sizeof(int)));
    LC_CUDA(cudaMalloc(&d_group_of_run, cap * sizeof(int)));
    LC_CUDA(cudaMalloc(&d_iota, cap * sizeof(int)));
    for (int k = 0; k < 2; ++k) { LC_CUDA(cudaMalloc(&d_keys[k], cap * sizeof(unsigned))); LC_CUDA(cudaMalloc(&d_vals[k], cap * sizeof(unsigned))); }
    LC_CUDA(cudaMalloc(&d_hist, (cap / 2048 + 2) * 512 * sizeof(unsigned)));
    LC_CUDA(cudaMalloc(&d_tc, (cap / 1024 + 2) * sizeof(unsigned)));
    LC_CUDA(cudaMalloc(&d_tb, (cap / 1024 + 2) * sizeof(unsigned)));
    LC_CUDA(cudaMalloc(&d_hp, cap * sizeof(unsigned)));
    LC_CUDA(cudaMalloc(&d_groups, cap * sizeof(RunGroup)));
    LC_CUDA(cudaMalloc(&d_iv, cap * sizeof(int3)));
    LC_RC(launch_rle_write(dev_labels, width, height, d_slice_ids, slices, d_offsets, d_runs, s));
    LC_RC(launch_run_components(d_runs, n_runs, d_offsets, height, 0, d_parent, d_keys[0], d_vals[0], s));
    unsigned *sk = nullptr, *sv = nullptr;
    int bits = 1;
    while ((1ull << bits) < n_runs) ++bits;
    LC_RC(launch_sort_pairs(d_keys[0], d_vals[0], d_keys[1], d_vals[1], n_runs, bits, d_hist, d_total, &sk, &sv, s));
    LC_RC(launch_group_runs(sk, sv, n_runs, d_runs, 0, d_tc, d_tb, d_hp, d_ngroups, d_groups, d_group_of_run, d_iv, s));
    LC_CUDA(cudaMemcpyAsync(&n_groups, d_ngroups, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
    LC_CUDA(cudaStreamSynchronize(s));
    *n_components_out = (int)n_groups;
    if (dev_component_out) {
      LC_RC(launch_init_iota(d_iota, (long long)n_groups, s));
      LC_RC(launch_relabel_groups(d_runs, n_runs, d_group_of_run, d_iota, width, height, dev_component_out, s));
    }
    if (host_records_out && record_cap > 0) {
      static_assert(sizeof(RunGroup) == 40, "record layout");
      const size_t n_copy = std::min<size_t>(n_groups, (size_t)record_cap);
      LC_CUDA(cudaMemcpyAsync(host_records_out, d_groups, n_copy * sizeof(RunGroup), cudaMemcpyDeviceToHost, s));
    }
    LC_CUDA(cudaStreamSynchronize(s));
  }
done:
#undef LC_CUDA
#undef LC_RC
  cleanup();
  return rc;
}

}  // extern "C"
