// b200_dense_segmentation.cpp -- see the header.  Host glue only: every frame's work happens in libvsb200.so.
#include "b200_dense_segmentation.h"

namespace segmentation {

void FrameResultToSegmentationDesc(const vsb200_frame_result& r, SegmentationDesc* desc) {
  CHECK_NOTNULL(desc);
  // Segmentation::RetrieveSegmentation3D (segmentation.cpp:473-478)
  desc->set_frame_width(r.width);
  desc->set_frame_height(r.height);
  desc->set_chunk_id(r.chunk_id);
  desc->set_connectedness(r.connectedness == 1 ? SegmentationDesc::N4_CONNECT : SegmentationDesc::N8_CONNECT);

  // AddRegion2DToSegmentationDesc (segmentation.cpp:671-697): id, rasterization, shape moments, already in id order
  for (int k = 0; k < r.n_regions; ++k) {
    SegmentationDesc::Region2D* region = desc->add_region();
    region->set_id(r.region_id[k]);
    SegmentationDesc::Rasterization* raster = region->mutable_raster();
    for (int i = r.interval_offset[k]; i < r.interval_offset[k + 1]; ++i) {
      SegmentationDesc::Rasterization::ScanInterval* scan = raster->add_scan_inter();
      scan->set_y(r.intervals[3 * i]);
      scan->set_left_x(r.intervals[3 * i + 1]);
      scan->set_right_x(r.intervals[3 * i + 2]);
    }
    const float* m = r.shape_moments + 6 * k;
    SegmentationDesc::ShapeMoments* moments = region->mutable_shape_moments();
    moments->set_size(m[0]);
    moments->set_mean_x(m[1]);
    moments->set_mean_y(m[2]);
    moments->set_moment_xx(m[3]);
    moments->set_moment_xy(m[4]);
    moments->set_moment_yy(m[5]);
  }

  // AddCompoundRegionToSegmentationDesc (segmentation.cpp:699-773), level 0 of a one-level hierarchy: no parent, no
  // children; only the chunk's first output frame carries it (dense_segmentation.cpp:378-381).
  if (r.n_compound > 0) {
    SegmentationDesc::HierarchyLevel* level = desc->add_hierarchy();
    for (int k = 0; k < r.n_compound; ++k) {
      SegmentationDesc::CompoundRegion* c = level->add_region();
      c->set_id(r.compound[4 * k]);
      c->set_size(r.compound[4 * k + 1]);
      for (int i = r.neighbor_offset[k]; i < r.neighbor_offset[k + 1]; ++i) {
        c->add_neighbor_id(r.neighbor_id[i]);
      }
      c->set_start_frame(r.compound[4 * k + 2]);
      c->set_end_frame(r.compound[4 * k + 3]);
    }
  }

  // DenseSegmentation::SegmentAndOutputChunk (dense_segmentation.cpp:385-387)
  desc->set_chunk_size(r.chunk_size);
  desc->set_overlap_start(r.overlap_start);
  desc->set_hierarchy_frame_idx(r.hierarchy_frame_idx);
}

B200DenseSegmentation::B200DenseSegmentation(const DenseSegmentationOptions& options, int frame_width, int frame_height,
                                             int cuda_device)
    : options_(options), frame_width_(frame_width), frame_height_(frame_height), device_(cuda_device) {
  CHECK_GE(options_.chunk_size, 3) << "Chunk size needs to be at least 3 frames.";   // dense_segmentation.cpp:54
}

B200DenseSegmentation::~B200DenseSegmentation() {
  if (handle_) vsb200_dense_destroy(handle_);
}

long long B200DenseSegmentation::KernelLaunches() const {
  if (!handle_) return 0;
  double s[9];
  vsb200_dense_stats(handle_, s);
  return (long long)s[7];
}

void B200DenseSegmentation::Pop(int n_ready, std::vector<std::unique_ptr<SegmentationDesc>>* results) {
  for (int k = 0; k < n_ready; ++k) {
    vsb200_frame_result r;
    CHECK_EQ(VSB200_OK, vsb200_dense_pop(handle_, &r)) << vsb200_last_error();
    std::unique_ptr<SegmentationDesc> desc(new SegmentationDesc());
    FrameResultToSegmentationDesc(r, desc.get());
    results->push_back(std::move(desc));
  }
}

int B200DenseSegmentation::ProcessFrame(bool flush, const std::vector<cv::Mat>* features, const cv::Mat* flow,
                                        std::vector<std::unique_ptr<SegmentationDesc>>* results) {
  CHECK_NOTNULL(results);
  results->clear();
  if (handle_ == nullptr) {
    // The engine is created on the first call, like the reference's Segmentation object (dense_segmentation.cpp:113-118):
    // whether a flow stream exists is only known from the first call's `flow` argument.
    vsb200_dense_opts o;
    vsb200_dense_default_opts(&o);
    o.presmoothing = options_.presmoothing == DenseSegmentationOptions::PRESMOOTH_NONE ? 0
                   : options_.presmoothing == DenseSegmentationOptions::PRESMOOTH_GAUSSIAN ? 1 : 2;
    o.frac_min_region_size = options_.frac_min_region_size;
    o.chunk_size = options_.chunk_size;
    o.chunk_overlap_ratio = options_.chunk_overlap_ratio;
    o.num_constraint_frames = options_.num_constraint_frames;
    o.two_stage_oversegment = options_.two_stage_oversegment;
    o.thin_structure_suppression = options_.thin_structure_suppression;
    o.enforce_n4_connectivity = options_.enforce_n4_connectivity;
    o.enforce_spatial_connectedness = options_.enforce_spatial_connectedness;
    o.color_distance = options_.color_distance == DenseSegmentationOptions::COLOR_DISTANCE_L1 ? 0 : 1;
    o.compute_vectorization = options_.compute_vectorization;
    o.device = device_;
    o.want_id_maps = 0;
    CHECK_EQ(VSB200_OK, vsb200_dense_create(&o, frame_width_, frame_height_, flow != nullptr, &handle_))
        << "B200 dense segmentation: " << vsb200_last_error();
  }

  int n_ready = 0;
  if (flush) {
    CHECK_EQ(VSB200_OK, vsb200_dense_flush(handle_, &n_ready)) << vsb200_last_error();
  } else {
    CHECK_NOTNULL(features);
    CHECK_EQ(features->size(), 1) << "Only appearance supported by default DenseSegmentation.";   // :203-204
    const cv::Mat& frame = (*features)[0];
    CHECK_EQ(frame.rows, frame_height_);      // dense_segmentation.cpp:175-176
    CHECK_EQ(frame.cols, frame_width_);
    CHECK_EQ(frame.type(), CV_8UC3) << "Expecting 8-bit BGR frames.";
    const float* flow_ptr = nullptr;
    int flow_step = 0;
    if (flow != nullptr && input_frames_ > 0) {     // the first frame's flow is an empty cv::Mat (:129-131)
      CHECK_EQ(frame_height_, flow->rows);
      CHECK_EQ(frame_width_, flow->cols);
      flow_ptr = flow->ptr<float>(0);
      flow_step = (int)flow->step[0];
    }
    CHECK_EQ(VSB200_OK, vsb200_dense_push(handle_, frame.ptr<uint8_t>(0), (int)frame.step[0], flow_ptr, flow_step,
                                          input_frames_, &n_ready))
        << vsb200_last_error();
    ++input_frames_;
  }
  Pop(n_ready, results);
  return (int)results->size();
}

}  // namespace segmentation
