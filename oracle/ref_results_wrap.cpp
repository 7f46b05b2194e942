// TEST INFRASTRUCTURE ONLY.  C wrapper around the REFERENCE's own streaming over-segmentation: DenseSegmentation
// (segmentation/dense_segmentation.cpp), Segmentation (segmentation.cpp), DenseSegmentationGraph
// (dense_segmentation_graph.h/.cpp), FastSegmentationGraph, the pixel distances, BilateralFilter and the
// segmentation_util.cpp helpers -- all compiled UNMODIFIED from /root/reference by `make -C oracle _ref` against the
// stand-ins in ref_shim/ (glog, gflags, cv::Mat, protobuf value classes, boost::disjoint_sets) -- used to pin the whole
// oracle engine (oracle/vso_engine.cpp, vso_graph.cpp, vso_shape.cpp) frame result by frame result.
// Call shape and result layout mirror oracle/vso.h (vso_dense_push / flush / pop, vso_frame_result).
#include <stdint.h>

#include <deque>
#include <memory>
#include <vector>

#include <opencv2/core/core.hpp>

#include "segmentation/dense_segmentation.h"

#include "ref_flatten.hpp"

namespace {

struct RefDense {
  std::unique_ptr<segmentation::DenseSegmentation> seg;
  int width = 0, height = 0;
  bool use_flow = false;
  std::deque<FlatResult> ready;
  FlatResult current;
};

int Run(RefDense* h, bool flush, const uint8_t* bgr, int stride, const float* flow, int flow_stride) {
  std::vector<std::unique_ptr<segmentation::SegmentationDesc>> results;
  if (flush) {
    h->seg->ProcessFrame(true, nullptr, nullptr, &results);
  } else {
    // segmentation_unit.cpp:124-140: a cv::Mat view on the frame bytes, the flow view (empty on the first frame)
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

}  // namespace

extern "C" {

// opts: the first eight fields of vso_dense_opts, same order.
void* ref_dense_create(int presmoothing, float frac_min_region_size, int chunk_size, float chunk_overlap_ratio, int num_constraint_frames,
                       int enforce_n4, int enforce_connected, int color_distance, int width, int height, int use_flow) {
  segmentation::DenseSegmentationOptions o;
  o.presmoothing = (segmentation::DenseSegmentationOptions::Presmoothing)presmoothing;
  o.frac_min_region_size = frac_min_region_size;
  o.chunk_size = chunk_size;
  o.chunk_overlap_ratio = chunk_overlap_ratio;
  o.num_constraint_frames = num_constraint_frames;
  o.enforce_n4_connectivity = enforce_n4 != 0;
  o.enforce_spatial_connectedness = enforce_connected != 0;
  o.color_distance = (segmentation::DenseSegmentationOptions::ColorDistance)color_distance;
  RefDense* h = new RefDense;
  h->width = width;
  h->height = height;
  h->use_flow = use_flow != 0;
  h->seg.reset(new segmentation::DenseSegmentation(o, width, height));
  return h;
}

int ref_dense_push(void* hv, const uint8_t* bgr, int stride, const float* flow, int flow_stride) { return Run((RefDense*)hv, false, bgr, stride, flow, flow_stride); }
int ref_dense_flush(void* hv) { return Run((RefDense*)hv, true, nullptr, 0, nullptr, 0); }

int ref_dense_pop(void* hv, RefFrameResult* out) {
  RefDense* h = (RefDense*)hv;
  if (h->ready.empty()) return -1;
  h->current = std::move(h->ready.front());
  h->ready.pop_front();
  Expose(h->current, out);
  return 0;
}

void ref_dense_destroy(void* hv) { delete (RefDense*)hv; }

}  // extern "C"
