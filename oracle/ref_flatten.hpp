// TEST INFRASTRUCTURE ONLY.  SegmentationDesc -> flat arrays in the layout of vso_frame_result / vsb200_frame_result,
// shared by the wrappers that expose reference-side objects to the Python tests (ref_results_wrap.cpp,
// tests/host_check_wrap.cpp).
#ifndef VSO_REF_FLATTEN_HPP_
#define VSO_REF_FLATTEN_HPP_
#include <stdint.h>

#include <vector>

#include "segment_util/segmentation.pb.h"

namespace {

struct FlatResult {
  int32_t head[8];  // width height chunk_id chunk_size overlap_start hierarchy_frame_idx connectedness n_regions
  std::vector<int32_t> region_id, interval_offset, intervals, compound, neighbor_offset, neighbor_id;
  std::vector<float> shape_moments;
};

struct RefFrameResult {  // == vso_frame_result
  int32_t width, height, chunk_id, chunk_size, overlap_start, hierarchy_frame_idx, connectedness, n_regions;
  const int32_t* region_id;
  const int32_t* interval_offset;
  const int32_t* intervals;
  const float* shape_moments;
  int32_t n_compound;
  const int32_t* compound;
  const int32_t* neighbor_offset;
  const int32_t* neighbor_id;
  int64_t pts;
};

inline void Flatten(const segmentation::SegmentationDesc& d, FlatResult* f) {
  f->head[0] = d.frame_width();
  f->head[1] = d.frame_height();
  f->head[2] = d.chunk_id();
  f->head[3] = d.chunk_size();
  f->head[4] = d.overlap_start();
  f->head[5] = d.hierarchy_frame_idx();
  f->head[6] = (int)d.connectedness();
  f->head[7] = d.region_size();
  f->interval_offset.push_back(0);
  for (const auto& r : d.region()) {
    f->region_id.push_back(r.id());
    for (const auto& s : r.raster().scan_inter()) {
      f->intervals.push_back(s.y());
      f->intervals.push_back(s.left_x());
      f->intervals.push_back(s.right_x());
    }
    f->interval_offset.push_back((int32_t)(f->intervals.size() / 3));
    const auto& m = r.shape_moments();
    const float v[6] = {m.size(), m.mean_x(), m.mean_y(), m.moment_xx(), m.moment_xy(), m.moment_yy()};
    f->shape_moments.insert(f->shape_moments.end(), v, v + 6);
  }
  f->neighbor_offset.push_back(0);
  if (d.hierarchy_size() > 0) {
    for (const auto& c : d.hierarchy(0).region()) {
      const int32_t v[4] = {c.id(), c.size(), c.start_frame(), c.end_frame()};
      f->compound.insert(f->compound.end(), v, v + 4);
      for (int k = 0; k < c.neighbor_id_size(); ++k) f->neighbor_id.push_back(c.neighbor_id(k));
      f->neighbor_offset.push_back((int32_t)f->neighbor_id.size());
    }
  }
}


inline void Expose(const FlatResult& f, RefFrameResult* out) {
  out->width = f.head[0]; out->height = f.head[1]; out->chunk_id = f.head[2]; out->chunk_size = f.head[3];
  out->overlap_start = f.head[4]; out->hierarchy_frame_idx = f.head[5]; out->connectedness = f.head[6]; out->n_regions = f.head[7];
  out->region_id = f.region_id.data();
  out->interval_offset = f.interval_offset.data();
  out->intervals = f.intervals.data();
  out->shape_moments = f.shape_moments.data();
  out->n_compound = (int32_t)(f.compound.size() / 4);
  out->compound = f.compound.data();
  out->neighbor_offset = f.neighbor_offset.data();
  out->neighbor_id = f.neighbor_id.data();
  out->pts = 0;
}

}  // namespace
#endif
