// TEST INFRASTRUCTURE ONLY.  C wrapper around the REFERENCE's two-stage pipeline: DenseSegmentation (over-segmentation)
// feeding RegionSegmentation (hierarchical stage: region descriptors, RegionAgglomerationGraph, hierarchy levels;
// segmentation/region_segmentation.cpp, region_segmentation_graph.cpp, region_descriptor.cpp, segmentation.cpp), the
// way seg_tree_sample chains DenseSegmentationUnit -> RegionSegmentationUnit (segmentation_unit.cpp:118-178,240-331).
// Compiled from /root/reference by `make -C oracle _ref` with ONE build-time edit (see the Makefile: a default argument
// in region_segmentation_graph.h that GCC 13 cannot use inside the enclosing class).  8-bit BGR->Lab (third party,
// cv::cvtColor) is the oracle's restatement, bit identical to cv2 4.13.  Produces the golden vectors the next
// SURVEY 8f row (N1, hierarchical merge) is built against: tests/golden/make_reference_hierarchy_golden.py.
#include <stdint.h>

#include <deque>
#include <memory>
#include <vector>

#include <opencv2/core/core.hpp>

#include "segmentation/dense_segmentation.h"
#include "segmentation/region_segmentation.h"

extern "C" void vso_bgr2lab(const uint8_t* bgr, int w, int h, int row_stride, uint8_t* lab_out);   // oracle/vso_region.cpp
extern "C" void vso_shim_bgr2lab(const unsigned char* bgr, int w, int h, int row_stride, unsigned char* lab_out) {
  vso_bgr2lab(bgr, w, h, row_stride, lab_out);
}

namespace {

using segmentation::SegmentationDesc;

struct Hier {
  std::unique_ptr<segmentation::DenseSegmentation> dense;
  std::unique_ptr<segmentation::RegionSegmentation> region;
  int width = 0, height = 0;
  bool use_flow = false;
  // frames (and flows) wait here until the dense stage releases their over-segmentation (chunk latency)
  std::deque<std::vector<uint8_t>> frames;
  std::deque<std::vector<float>> flows;
  int region_inputs = 0;
  std::deque<std::unique_ptr<SegmentationDesc>> ready;
  std::vector<int32_t> flat;
};

void FeedRegionStage(Hier* h, std::vector<std::unique_ptr<SegmentationDesc>>* overseg) {
  for (auto& desc : *overseg) {
    std::vector<cv::Mat> features;
    features.push_back(cv::Mat(h->height, h->width, CV_8UC3, h->frames.front().data(), (size_t)h->width * 3));
    if (h->use_flow) {      // RegionSegmentationUnit::ExtractFrameSetFeatures (segmentation_unit.cpp:310-331)
      if (h->region_inputs > 0) features.push_back(cv::Mat(h->height, h->width, CV_32FC2, h->flows.front().data(), (size_t)h->width * 8));
      else features.push_back(cv::Mat());
    }
    std::vector<std::unique_ptr<SegmentationDesc>> results;
    h->region->ProcessFrame(false, desc.get(), &features, &results);
    for (auto& r : results) h->ready.push_back(std::move(r));
    h->frames.pop_front();
    if (h->use_flow) h->flows.pop_front();
    ++h->region_inputs;
  }
}

}  // namespace

extern "C" {

void* ref_hier_create(int width, int height, int use_flow, int dense_chunk_size, int chunk_set_size, int chunk_set_overlap,
                      int min_region_num, float level_cutoff_fraction) {
  Hier* h = new Hier;
  h->width = width;
  h->height = height;
  h->use_flow = use_flow != 0;
  segmentation::DenseSegmentationOptions d;
  d.chunk_size = dense_chunk_size;
  h->dense.reset(new segmentation::DenseSegmentation(d, width, height));
  segmentation::RegionSegmentationOptions r;
  r.chunk_set_size = chunk_set_size;
  r.chunk_set_overlap = chunk_set_overlap;
  r.min_region_num = min_region_num;
  r.level_cutoff_fraction = level_cutoff_fraction;
  r.use_flow = use_flow != 0;           // RegionSegmentationUnit::CreateRegionSegmentation (segmentation_unit.cpp:303-308)
  r.compute_vectorization = false;      // third party (cv::approxPolyDP), SURVEY row N3
  h->region.reset(new segmentation::RegionSegmentation(r, width, height));
  return h;
}

// One frame (flow: interleaved x,y floats or null); returns the number of hierarchical results that became ready.
int ref_hier_push(void* hv, const uint8_t* bgr, const float* flow) {
  Hier* h = (Hier*)hv;
  const size_t before = h->ready.size();
  h->frames.emplace_back(bgr, bgr + (size_t)h->width * h->height * 3);
  if (h->use_flow) {
    if (flow) h->flows.emplace_back(flow, flow + (size_t)h->width * h->height * 2);
    else h->flows.emplace_back();
  }
  std::vector<cv::Mat> features(1, cv::Mat(h->height, h->width, CV_8UC3, h->frames.back().data(), (size_t)h->width * 3));
  cv::Mat flow_mat;
  if (h->use_flow && flow) flow_mat = cv::Mat(h->height, h->width, CV_32FC2, h->flows.back().data(), (size_t)h->width * 8);
  std::vector<std::unique_ptr<SegmentationDesc>> overseg;
  h->dense->ProcessFrame(false, &features, h->use_flow ? &flow_mat : nullptr, &overseg);
  FeedRegionStage(h, &overseg);
  return (int)(h->ready.size() - before);
}

int ref_hier_flush(void* hv) {
  Hier* h = (Hier*)hv;
  const size_t before = h->ready.size();
  std::vector<std::unique_ptr<SegmentationDesc>> overseg;
  h->dense->ProcessFrame(true, nullptr, nullptr, &overseg);
  FeedRegionStage(h, &overseg);
  std::vector<std::unique_ptr<SegmentationDesc>> results;
  h->region->ProcessFrame(true, nullptr, nullptr, &results);
  for (auto& r : results) h->ready.push_back(std::move(r));
  return (int)(h->ready.size() - before);
}

// Pops one result as a flat int32 record (floats as bits):
//   width height chunk_id chunk_size overlap_start hierarchy_frame_idx n_regions n_levels
//   per region: id n_intervals (y lx rx)* 6 x shape-moment bits
//   per level: n_compound, per compound: id size parent_id start_frame end_frame n_neighbors n_children neighbors* children*
// Returns the number of int32 words (0 if nothing is ready); *out points into the handle until the next pop.
long long ref_hier_pop(void* hv, const int32_t** out) {
  Hier* h = (Hier*)hv;
  if (h->ready.empty()) return 0;
  std::unique_ptr<SegmentationDesc> d = std::move(h->ready.front());
  h->ready.pop_front();
  std::vector<int32_t>& f = h->flat;
  f.clear();
  auto bits = [](float v) { int32_t b; memcpy(&b, &v, 4); return b; };
  const int32_t head[8] = {d->frame_width(), d->frame_height(), d->chunk_id(), d->chunk_size(), d->overlap_start(),
                           d->hierarchy_frame_idx(), d->region_size(), d->hierarchy_size()};
  f.insert(f.end(), head, head + 8);
  for (const auto& r : d->region()) {
    f.push_back(r.id());
    f.push_back(r.raster().scan_inter_size());
    for (const auto& s : r.raster().scan_inter()) { f.push_back(s.y()); f.push_back(s.left_x()); f.push_back(s.right_x()); }
    const auto& m = r.shape_moments();
    for (float v : {m.size(), m.mean_x(), m.mean_y(), m.moment_xx(), m.moment_xy(), m.moment_yy()}) f.push_back(bits(v));
  }
  for (const auto& level : d->hierarchy()) {
    f.push_back(level.region_size());
    for (const auto& c : level.region()) {
      f.push_back(c.id()); f.push_back(c.size()); f.push_back(c.parent_id()); f.push_back(c.start_frame()); f.push_back(c.end_frame());
      f.push_back(c.neighbor_id_size()); f.push_back(c.child_id_size());
      for (int k = 0; k < c.neighbor_id_size(); ++k) f.push_back(c.neighbor_id(k));
      for (int k = 0; k < c.child_id_size(); ++k) f.push_back(c.child_id(k));
    }
  }
  *out = f.data();
  return (long long)f.size();
}

void ref_hier_destroy(void* hv) { delete (Hier*)hv; }

}  // extern "C"
