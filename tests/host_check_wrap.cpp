// TEST INFRASTRUCTURE ONLY.  Exposes the C++ host side (video_segment_b200/host/b200_dense_segmentation.{h,cpp}, the
// class a maintainer puts in place of segmentation::DenseSegmentation) to the Python tests, next to the compiled
// reference (oracle/_ref/libref_results.so):
//   * host_check_desc_vs_reference: the reference's OWN SegmentationDesc objects against the ones
//     FrameResultToSegmentationDesc rebuilds from the flat result arrays -- values and presence bits (CPU test);
//   * ref_io_*: the reference's own SegmentationWriter / SegmentationReader / StripToEssentials
//     (segment_util/segmentation_io.cpp, unmodified) for byte-for-byte checks of csrc/pb_io.cu (CPU test);
//   * b200_dense_*: the call shape of ref_dense_* (oracle/ref_results_wrap.cpp) over B200DenseSegmentation (GPU test).
#include <stdint.h>
#include <string.h>

#include <deque>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

#include "b200_dense_segmentation.h"
#include "ref_flatten.hpp"
#include "segment_util/segmentation_io.h"
#include "segmentation/dense_segmentation.h"

namespace {

using segmentation::SegmentationDesc;

static_assert(sizeof(RefFrameResult) == sizeof(vsb200_frame_result), "frame result layouts differ");

// First difference between two messages, values and has_ bits ("" if none).
std::string DescDifference(const SegmentationDesc& a, const SegmentationDesc& b) {
  // (std::string, not a stream: see the note on FatalStream in oracle/ref_shim/glog/logging.h)
#define SAME(expr) if (!((a.expr) == (b.expr))) return std::string(#expr);
  SAME(has_frame_width()) SAME(frame_width()) SAME(has_frame_height()) SAME(frame_height())
  SAME(has_chunk_size()) SAME(chunk_size()) SAME(has_overlap_start()) SAME(overlap_start())
  SAME(has_chunk_id()) SAME(chunk_id()) SAME(has_hierarchy_frame_idx()) SAME(hierarchy_frame_idx())
  SAME(has_connectedness()) SAME(connectedness()) SAME(has_rasterization_removed()) SAME(has_vector_mesh())
  SAME(features_size()) SAME(region_size()) SAME(hierarchy_size())
#undef SAME
#define SAME(expr) if (!((x.expr) == (y.expr))) return std::string(what) + " " + std::to_string(k) + ": " #expr;
  const char* what = "region";
  for (int k = 0; k < a.region_size(); ++k) {
    const auto& x = a.region(k);
    const auto& y = b.region(k);
    SAME(has_id()) SAME(id()) SAME(has_raster()) SAME(has_shape_moments()) SAME(has_vectorization())
    SAME(raster().scan_inter_size())
    for (int i = 0; i < x.raster().scan_inter_size(); ++i) {
      SAME(raster().scan_inter(i).y()) SAME(raster().scan_inter(i).left_x()) SAME(raster().scan_inter(i).right_x())
      SAME(raster().scan_inter(i).has_y()) SAME(raster().scan_inter(i).has_left_x()) SAME(raster().scan_inter(i).has_right_x())
    }
    SAME(shape_moments().has_size()) SAME(shape_moments().has_mean_x()) SAME(shape_moments().has_mean_y())
    SAME(shape_moments().has_moment_xx()) SAME(shape_moments().has_moment_xy()) SAME(shape_moments().has_moment_yy())
    SAME(shape_moments().size()) SAME(shape_moments().mean_x()) SAME(shape_moments().mean_y())
    SAME(shape_moments().moment_xx()) SAME(shape_moments().moment_xy()) SAME(shape_moments().moment_yy())
  }
  what = "compound";
  for (int l = 0; l < a.hierarchy_size(); ++l) {
    if (a.hierarchy(l).region_size() != b.hierarchy(l).region_size()) return "hierarchy " + std::to_string(l) + " size";
    for (int k = 0; k < a.hierarchy(l).region_size(); ++k) {
      const auto& x = a.hierarchy(l).region(k);
      const auto& y = b.hierarchy(l).region(k);
      SAME(has_id()) SAME(id()) SAME(has_size()) SAME(size()) SAME(has_parent_id()) SAME(parent_id())
      SAME(has_start_frame()) SAME(start_frame()) SAME(has_end_frame()) SAME(end_frame())
      SAME(child_id_size()) SAME(neighbor_id_size())
      for (int i = 0; i < x.neighbor_id_size(); ++i) SAME(neighbor_id(i))
    }
  }
#undef SAME
  return "";
}

struct HostDense {
  std::unique_ptr<segmentation::B200DenseSegmentation> seg;
  int width = 0, height = 0;
  bool use_flow = false;
  std::deque<FlatResult> ready;
  FlatResult current;
};

segmentation::DenseSegmentationOptions MakeOptions(int presmoothing, float frac_min_region_size, int chunk_size, float chunk_overlap_ratio,
                                                   int num_constraint_frames, int enforce_n4, int enforce_connected, int color_distance) {
  segmentation::DenseSegmentationOptions o;
  o.presmoothing = (segmentation::DenseSegmentationOptions::Presmoothing)presmoothing;
  o.frac_min_region_size = frac_min_region_size;
  o.chunk_size = chunk_size;
  o.chunk_overlap_ratio = chunk_overlap_ratio;
  o.num_constraint_frames = num_constraint_frames;
  o.enforce_n4_connectivity = enforce_n4 != 0;
  o.enforce_spatial_connectedness = enforce_connected != 0;
  o.color_distance = (segmentation::DenseSegmentationOptions::ColorDistance)color_distance;
  return o;
}

}  // namespace

extern "C" {

// Runs the REFERENCE's DenseSegmentation over a clip (frames: [t][h][w][3] bytes; flows: [t][h][w][2] floats or null)
// and checks every SegmentationDesc it returns against FrameResultToSegmentationDesc(Flatten(desc)).  Returns the
// number of frames that differ (0 = the host side rebuilds the reference's messages exactly), -1 on a count mismatch;
// msg receives the first difference.
int host_check_desc_vs_reference(const uint8_t* frames, const float* flows, int t, int width, int height, int presmoothing,
                                 float frac_min_region_size, int chunk_size, float chunk_overlap_ratio, int num_constraint_frames,
                                 int enforce_n4, int enforce_connected, int color_distance, char* msg, int msg_cap) {
  segmentation::DenseSegmentation ref(MakeOptions(presmoothing, frac_min_region_size, chunk_size, chunk_overlap_ratio,
                                                  num_constraint_frames, enforce_n4, enforce_connected, color_distance),
                                      width, height);
  int bad = 0, seen = 0;
  std::string first;
  for (int k = 0; k <= t; ++k) {
    std::vector<std::unique_ptr<SegmentationDesc>> results;
    if (k == t) {
      ref.ProcessFrame(true, nullptr, nullptr, &results);
    } else {
      std::vector<cv::Mat> features(1, cv::Mat(height, width, CV_8UC3, (void*)(frames + (size_t)k * height * width * 3), (size_t)width * 3));
      cv::Mat flow;
      if (flows && k > 0) flow = cv::Mat(height, width, CV_32FC2, (void*)(flows + (size_t)k * height * width * 2), (size_t)width * 8);
      ref.ProcessFrame(false, &features, flows ? &flow : nullptr, &results);
    }
    for (const auto& d : results) {
      FlatResult f;
      Flatten(*d, &f);
      RefFrameResult view;
      Expose(f, &view);
      vsb200_frame_result r;
      memcpy(&r, &view, sizeof(r));
      SegmentationDesc rebuilt;
      segmentation::FrameResultToSegmentationDesc(r, &rebuilt);
      const std::string diff = DescDifference(*d, rebuilt);
      if (!diff.empty()) {
        if (first.empty()) first = "frame " + std::to_string(seen) + ": " + diff;
        ++bad;
      }
      ++seen;
    }
  }
  if (msg && msg_cap > 0) {
    strncpy(msg, first.c_str(), msg_cap - 1);
    msg[msg_cap - 1] = 0;
  }
  return seen == t ? bad : -1;
}

// SegmentationWriter over opaque frame payloads: HEAD with `entries`, AddSegmentationDataToChunk per frame, WriteChunk
// after every `chunk_every` frames (0: never), WriteTermHeaderAndClose.
int ref_io_write(const char* filename, const int32_t* entries, int n_entries, const uint8_t* blob, const int64_t* sizes,
                 const int64_t* pts, int n_frames, int chunk_every) {
  segmentation::SegmentationWriter w(filename);
  if (!w.OpenFile(std::vector<int>(entries, entries + n_entries))) return -1;
  size_t pos = 0;
  for (int k = 0; k < n_frames; ++k) {
    w.AddSegmentationDataToChunk(std::string((const char*)blob + pos, (size_t)sizes[k]), pts[k]);
    pos += (size_t)sizes[k];
    if (chunk_every > 0 && (k + 1) % chunk_every == 0) w.WriteChunk();
  }
  w.WriteTermHeaderAndClose();
  return 0;
}

// SegmentationReader: frame count, then per frame pts and payload (concatenated into blob, sizes out).
int ref_io_read(const char* filename, int32_t* flags, int flags_cap, int* n_flags, uint8_t* blob, size_t blob_cap, int64_t* sizes,
                int64_t* pts, int frames_cap) {
  segmentation::SegmentationReader r(filename);
  if (!r.OpenFileAndReadHeaders()) return -1;
  *n_flags = (int)r.GetHeaderFlags().size();
  for (int i = 0; i < *n_flags && i < flags_cap; ++i) flags[i] = r.GetHeaderFlags()[i];
  const int n = r.NumFrames();
  size_t pos = 0;
  for (int k = 0; k < n && k < frames_cap; ++k) {
    std::string data;
    r.SeekToFrame(k);
    if (!r.ReadNextFrameBinary(&data)) return -1;
    if (pos + data.size() > blob_cap) return -1;
    memcpy(blob + pos, data.data(), data.size());
    pos += data.size();
    sizes[k] = (int64_t)data.size();
    pts[k] = r.TimeStamps()[k];
  }
  return n;
}

// StripToEssentials(desc, false, save_shape_moments) on the message FrameResultToSegmentationDesc builds from `r`.
long long ref_io_strip(const RefFrameResult* r, int save_shape_moments, uint8_t* buf, size_t cap) {
  vsb200_frame_result fr;
  memcpy(&fr, r, sizeof(fr));
  SegmentationDesc desc;
  segmentation::FrameResultToSegmentationDesc(fr, &desc);
  std::string out;
  segmentation::StripToEssentials(desc, false, save_shape_moments != 0, &out);
  if (buf && cap) memcpy(buf, out.data(), out.size() < cap ? out.size() : cap);
  return (long long)out.size();
}

void* b200_dense_create(int presmoothing, float frac_min_region_size, int chunk_size, float chunk_overlap_ratio, int num_constraint_frames,
                        int enforce_n4, int enforce_connected, int color_distance, int width, int height, int use_flow) {
  HostDense* h = new HostDense;
  h->width = width;
  h->height = height;
  h->use_flow = use_flow != 0;
  h->seg.reset(new segmentation::B200DenseSegmentation(
      MakeOptions(presmoothing, frac_min_region_size, chunk_size, chunk_overlap_ratio, num_constraint_frames, enforce_n4,
                  enforce_connected, color_distance),
      width, height, 0));
  return h;
}

static int RunHost(HostDense* h, bool flush, const uint8_t* bgr, int stride, const float* flow, int flow_stride) {
  std::vector<std::unique_ptr<SegmentationDesc>> results;
  if (flush) {
    h->seg->ProcessFrame(true, nullptr, nullptr, &results);
  } else {
    std::vector<cv::Mat> features(1, cv::Mat(h->height, h->width, CV_8UC3, (void*)bgr, (size_t)stride));
    cv::Mat flow_mat;
    if (h->use_flow && flow) flow_mat = cv::Mat(h->height, h->width, CV_32FC2, (void*)flow, (size_t)flow_stride);
    h->seg->ProcessFrame(false, &features, h->use_flow ? &flow_mat : nullptr, &results);
  }
  for (const auto& r : results) {
    h->ready.emplace_back();
    Flatten(*r, &h->ready.back());
  }
  return (int)results.size();
}

int b200_dense_push(void* hv, const uint8_t* bgr, int stride, const float* flow, int flow_stride) { return RunHost((HostDense*)hv, false, bgr, stride, flow, flow_stride); }
int b200_dense_flush(void* hv) { return RunHost((HostDense*)hv, true, nullptr, 0, nullptr, 0); }
int b200_dense_pop(void* hv, RefFrameResult* out) {
  HostDense* h = (HostDense*)hv;
  if (h->ready.empty()) return -1;
  h->current = std::move(h->ready.front());
  h->ready.pop_front();
  Expose(h->current, out);
  return 0;
}
long long b200_dense_kernel_launches(void* hv) { return ((HostDense*)hv)->seg->KernelLaunches(); }
void b200_dense_destroy(void* hv) { delete (HostDense*)hv; }

}  // extern "C"
