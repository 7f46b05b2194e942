// b200_segmentation_units.cpp -- see b200_segmentation_units.h.
#include "b200_segmentation_units.h"

#include <gflags/gflags.h>
#include <glog/logging.h>

// The command-line overrides of the reference's DenseSegmentation / RegionSegmentation constructors
// (dense_segmentation.cpp:39-46,55-101; region_segmentation.cpp:36-47,84-95) are honoured here, by reading the flags the
// reference defines (they live in the reference's own translation units when this file is linked into its tree).
DECLARE_string(dense_smoothing);
DECLARE_string(dense_color_dist);
DECLARE_double(dense_min_region_size);
DECLARE_int32(chunk_size);
DECLARE_int32(min_region_num);
DECLARE_int32(max_region_num);
DECLARE_double(level_cutoff_fraction);
DECLARE_double(small_region_penalizer);
DECLARE_int32(chunk_set_size);

namespace segmentation {

using video_framework::DataStream;
using video_framework::DenseFlowFrame;
using video_framework::Frame;
using video_framework::FrameSetPtr;
using video_framework::PointerFrame;
using video_framework::SegmentationStream;
using video_framework::StreamSet;
using video_framework::VideoFrame;
using video_framework::VideoStream;

namespace {

DenseSegmentationOptions WithDenseFlags(DenseSegmentationOptions o) {      // dense_segmentation.cpp:55-101
  if (FLAGS_chunk_size >= 3) o.chunk_size = FLAGS_chunk_size;
  if (!FLAGS_dense_smoothing.empty()) {
    if (FLAGS_dense_smoothing == "bilateral") o.presmoothing = DenseSegmentationOptions::PRESMOOTH_BILATERAL;
    else if (FLAGS_dense_smoothing == "gaussian") o.presmoothing = DenseSegmentationOptions::PRESMOOTH_GAUSSIAN;
    else LOG(ERROR) << "Undefined smoothing mode specified. Ignoring.";
  }
  if (!FLAGS_dense_color_dist.empty()) {
    if (FLAGS_dense_color_dist == "l1") o.color_distance = DenseSegmentationOptions::COLOR_DISTANCE_L1;
    else if (FLAGS_dense_color_dist == "l2") o.color_distance = DenseSegmentationOptions::COLOR_DISTANCE_L2;
    else LOG(ERROR) << "Undefined color distance specified. Ignoring.";
  }
  if (FLAGS_dense_min_region_size >= 1e-3) o.frac_min_region_size = FLAGS_dense_min_region_size;
  return o;
}

RegionSegmentationOptions WithRegionFlags(RegionSegmentationOptions o) {   // region_segmentation.cpp:52-95
  if (FLAGS_chunk_set_size >= 2) o.chunk_set_size = FLAGS_chunk_set_size;
  if (FLAGS_min_region_num > 0) o.min_region_num = FLAGS_min_region_num;
  if (FLAGS_max_region_num > 0) o.max_region_num = FLAGS_max_region_num;
  if (FLAGS_level_cutoff_fraction > 0) o.level_cutoff_fraction = std::max(0.95, FLAGS_level_cutoff_fraction);   // as written in the reference
  if (FLAGS_small_region_penalizer >= 0) o.small_region_penalizer = FLAGS_small_region_penalizer;
  return o;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// B200RegionSegmentation
// ---------------------------------------------------------------------------------------------
B200RegionSegmentation::B200RegionSegmentation(const RegionSegmentationOptions& options, int frame_width, int frame_height, int cuda_device)
    : options_(WithRegionFlags(options)), frame_width_(frame_width), frame_height_(frame_height), device_(cuda_device) {
  vsb200_region_opts o;
  vsb200_region_default_opts(&o);
  o.min_region_num = options_.min_region_num; o.max_region_num = options_.max_region_num;
  o.level_cutoff_fraction = options_.level_cutoff_fraction; o.small_region_penalizer = options_.small_region_penalizer;
  o.luminance_bins = options_.luminance_bins; o.color_bins = options_.color_bins; o.flow_bins = options_.flow_bins;
  o.chunk_set_size = options_.chunk_set_size; o.chunk_set_overlap = options_.chunk_set_overlap; o.constraint_chunks = options_.constraint_chunks;
  o.save_descriptors = options_.save_descriptors; o.use_appearance = options_.use_appearance; o.use_flow = options_.use_flow;
  o.use_size_penalizer = options_.use_size_penalizer; o.compute_vectorization = options_.compute_vectorization;
  o.device = device_;
  CHECK_EQ(VSB200_OK, vsb200_region_create(&o, frame_width_, frame_height_, &handle_)) << "B200 region segmentation: " << vsb200_last_error();
}

B200RegionSegmentation::~B200RegionSegmentation() {
  if (handle_) vsb200_region_destroy(handle_);
}

long long B200RegionSegmentation::KernelLaunches() const {
  double s[2] = {0, 0};
  vsb200_region_stats(handle_, s);
  return (long long)s[0];
}

void RegionRecordToSegmentationDesc(const int32_t* rec, long long words, SegmentationDesc* desc) {
  auto f32 = [](int32_t b) { float v; memcpy(&v, &b, 4); return v; };
  desc->set_frame_width(rec[0]);
  desc->set_frame_height(rec[1]);
  desc->set_chunk_id(rec[2]);
  desc->set_connectedness(SegmentationDesc::N4_CONNECT);     // Segmentation's default options in the region stage (segmentation.cpp:477-480)
  const int n_regions = rec[6], n_levels = rec[7];
  long long p = 8;
  for (int k = 0; k < n_regions; ++k) {
    SegmentationDesc::Region2D* r = desc->add_region();
    r->set_id(rec[p]);
    const int n = rec[p + 1];
    p += 2;
    SegmentationDesc::Rasterization* raster = r->mutable_raster();
    for (int i = 0; i < n; ++i, p += 3) {
      SegmentationDesc::Rasterization::ScanInterval* s = raster->add_scan_inter();
      s->set_y(rec[p]); s->set_left_x(rec[p + 1]); s->set_right_x(rec[p + 2]);
    }
    SegmentationDesc::ShapeMoments* m = r->mutable_shape_moments();
    m->set_size(f32(rec[p])); m->set_mean_x(f32(rec[p + 1])); m->set_mean_y(f32(rec[p + 2]));
    m->set_moment_xx(f32(rec[p + 3])); m->set_moment_xy(f32(rec[p + 4])); m->set_moment_yy(f32(rec[p + 5]));
    p += 6;
  }
  for (int l = 0; l < n_levels; ++l) {
    SegmentationDesc::HierarchyLevel* level = desc->add_hierarchy();
    const int nc = rec[p++];
    for (int c = 0; c < nc; ++c) {
      SegmentationDesc::CompoundRegion* cr = level->add_region();
      cr->set_id(rec[p]); cr->set_size(rec[p + 1]);
      if (l < n_levels - 1) cr->set_parent_id(rec[p + 2]);     // the top level has no parent (segmentation.cpp:728-730)
      const int nn = rec[p + 5], nch = rec[p + 6];
      const int sf = rec[p + 3], ef = rec[p + 4];
      p += 7;
      for (int i = 0; i < nn; ++i) cr->add_neighbor_id(rec[p++]);
      for (int i = 0; i < nch; ++i) cr->add_child_id(rec[p++]);
      cr->set_start_frame(sf); cr->set_end_frame(ef);
    }
  }
  desc->set_chunk_size(rec[3]);
  desc->set_overlap_start(rec[4]);
  desc->set_hierarchy_frame_idx(rec[5]);
  CHECK_EQ(p, words) << "malformed region record";
}

void B200RegionSegmentation::Pop(int n_ready, std::vector<std::unique_ptr<SegmentationDesc>>* results) {
  for (int k = 0; k < n_ready; ++k) {
    const int32_t* rec = nullptr;
    const long long words = vsb200_region_pop(handle_, &rec);
    CHECK_GT(words, 0) << vsb200_last_error();
    std::unique_ptr<SegmentationDesc> desc(new SegmentationDesc());
    RegionRecordToSegmentationDesc(rec, words, desc.get());
    results->push_back(std::move(desc));
  }
}

int B200RegionSegmentation::ProcessFrame(bool flush, const SegmentationDesc* desc, const std::vector<cv::Mat>* features,
                                         std::vector<std::unique_ptr<SegmentationDesc>>* results) {
  CHECK_NOTNULL(results);
  int n_ready = 0;
  if (desc == nullptr || features == nullptr) {
    CHECK(desc == nullptr && features == nullptr) << "Requring both segmentation and features to be either set or null.";   // :102-106
  } else {
    // SegmentationDesc -> the arrays of a vsb200_frame_result
    std::vector<int32_t> region_id, interval_offset(1, 0), intervals, compound, neighbor_offset(1, 0), neighbor_id;
    std::vector<float> moments;
    for (const auto& r : desc->region()) {
      region_id.push_back(r.id());
      for (const auto& s : r.raster().scan_inter()) { intervals.push_back(s.y()); intervals.push_back(s.left_x()); intervals.push_back(s.right_x()); }
      interval_offset.push_back((int32_t)(intervals.size() / 3));
      const auto& m = r.shape_moments();
      for (float v : {m.size(), m.mean_x(), m.mean_y(), m.moment_xx(), m.moment_xy(), m.moment_yy()}) moments.push_back(v);
    }
    if (desc->hierarchy_size() > 0) {
      for (const auto& c : desc->hierarchy(0).region()) {
        compound.push_back(c.id()); compound.push_back(c.size()); compound.push_back(c.start_frame()); compound.push_back(c.end_frame());
        for (int i = 0; i < c.neighbor_id_size(); ++i) neighbor_id.push_back(c.neighbor_id(i));
        neighbor_offset.push_back((int32_t)neighbor_id.size());
      }
    }
    vsb200_frame_result r;
    memset(&r, 0, sizeof(r));
    r.width = desc->frame_width(); r.height = desc->frame_height(); r.chunk_id = desc->chunk_id();
    r.chunk_size = desc->chunk_size(); r.overlap_start = desc->overlap_start(); r.hierarchy_frame_idx = desc->hierarchy_frame_idx();
    r.connectedness = (int32_t)desc->connectedness();
    r.n_regions = (int32_t)region_id.size();
    r.region_id = region_id.data(); r.interval_offset = interval_offset.data(); r.intervals = intervals.data(); r.shape_moments = moments.data();
    r.n_compound = (int32_t)(compound.size() / 4);
    r.compound = compound.data(); r.neighbor_offset = neighbor_offset.data(); r.neighbor_id = neighbor_id.data();
    CHECK_GE(features->size(), 1u);
    const cv::Mat& frame = (*features)[0];
    CHECK_EQ(frame.rows, frame_height_);
    CHECK_EQ(frame.cols, frame_width_);
    const float* flow_ptr = nullptr;
    int flow_step = 0;
    if (options_.use_flow && features->size() > 1 && !(*features)[1].empty()) {
      flow_ptr = (*features)[1].ptr<float>(0);
      flow_step = (int)(*features)[1].step[0];
    }
    CHECK_EQ(VSB200_OK, vsb200_region_push(handle_, &r, frame.ptr<uint8_t>(0), (int)frame.step[0], flow_ptr, flow_step, &n_ready))
        << vsb200_last_error();
    Pop(n_ready, results);
  }
  if (flush) {
    CHECK_EQ(VSB200_OK, vsb200_region_flush(handle_, &n_ready)) << vsb200_last_error();
    Pop(n_ready, results);
  }
  return (int)results->size();
}

// ---------------------------------------------------------------------------------------------
// B200DenseSegmentationUnit  <->  DenseSegmentationUnit (segmentation_unit.cpp:48-178)
// ---------------------------------------------------------------------------------------------
B200DenseSegmentationUnit::B200DenseSegmentationUnit(const DenseSegmentationUnitOptions& options,
                                                     const DenseSegmentationOptions* dense_seg_options, int cuda_device)
    : options_(options), device_(cuda_device) {
  if (dense_seg_options) dense_seg_options_ = *dense_seg_options;
}

bool B200DenseSegmentationUnit::OpenStreams(StreamSet* set) {
  video_stream_idx_ = FindStreamIdx(options_.video_stream_name, set);
  if (video_stream_idx_ < 0) { LOG(ERROR) << "Could not find video stream!\n"; return false; }
  const VideoStream& vid_stream = set->at(video_stream_idx_)->As<VideoStream>();
  frame_width_ = vid_stream.frame_width();
  frame_height_ = vid_stream.frame_height();
  if (vid_stream.pixel_format() != video_framework::PIXEL_FORMAT_BGR24) { LOG(ERROR) << "Expecting video format to be BGR24.\n"; return false; }
  if (!options_.flow_stream_name.empty()) {
    flow_stream_idx_ = FindStreamIdx(options_.flow_stream_name, set);
    if (flow_stream_idx_ < 0) { LOG(ERROR) << "Flow stream specified but not present"; return false; }
  } else {
    flow_stream_idx_ = -1;
  }
  set->push_back(std::shared_ptr<DataStream>(new SegmentationStream(frame_width_, frame_height_, options_.segment_stream_name)));
  if (vsb200_device_count() <= 0) {
    LOG(ERROR) << "B200 dense segmentation: no sm_100 CUDA device available (this path has no CPU fallback)";
    return false;
  }
  dense_seg_.reset(new B200DenseSegmentation(WithDenseFlags(dense_seg_options_), frame_width_, frame_height_, device_));
  SetRateBufferSize(dense_seg_->ChunkSize() * 3);
  return true;
}

void B200DenseSegmentationUnit::ProcessFrame(FrameSetPtr input, std::list<FrameSetPtr>* output) {
  VLOG(1) << "Processing frame #" << input_frames_;
  std::vector<cv::Mat> features;
  const VideoFrame& video_frame = input->at(video_stream_idx_)->As<VideoFrame>();
  cv::Mat mat_view;
  video_frame.MatView(&mat_view);
  features.push_back(mat_view);
  cv::Mat flow;
  if (input_frames_ > 0 && flow_stream_idx_ >= 0) {
    const DenseFlowFrame& flow_frame = input->at(flow_stream_idx_)->As<DenseFlowFrame>();
    flow = flow_frame.MatViewInterleaved();
  }
  frame_set_buffer_.push_back(input);
  ++input_frames_;
  std::vector<std::unique_ptr<SegmentationDesc>> results;
  if (dense_seg_->ProcessFrame(false, &features, flow_stream_idx_ >= 0 ? &flow : nullptr, &results) > 0) OutputSegmentation(&results, output);
}

bool B200DenseSegmentationUnit::PostProcess(std::list<FrameSetPtr>* append) {
  std::vector<std::unique_ptr<SegmentationDesc>> results;
  if (dense_seg_->ProcessFrame(true, nullptr, nullptr, &results) > 0) OutputSegmentation(&results, append);
  return false;
}

void B200DenseSegmentationUnit::OutputSegmentation(std::vector<std::unique_ptr<SegmentationDesc>>* results, std::list<FrameSetPtr>* output) {
  for (size_t k = 0; k < results->size(); ++k) {
    FrameSetPtr frame_set = frame_set_buffer_.front();
    frame_set_buffer_.pop_front();
    const int64_t pts = frame_set->at(video_stream_idx_)->pts();
    frame_set->push_back(std::shared_ptr<Frame>(new PointerFrame<SegmentationDesc>(std::move((*results)[k]), pts)));
    output->push_back(frame_set);
    ++output_frames_;
  }
  LOG(INFO) << "__STREAMING_SIZE__: " << output_frames_ << "\n";        // progress marker, kept byte-identical (:177)
}

// ---------------------------------------------------------------------------------------------
// B200RegionSegmentationUnit  <->  RegionSegmentationUnit (segmentation_unit.cpp:180-331)
// ---------------------------------------------------------------------------------------------
B200RegionSegmentationUnit::B200RegionSegmentationUnit(const RegionSegmentationUnitOptions& options,
                                                       const RegionSegmentationOptions* region_options, int cuda_device)
    : options_(options), device_(cuda_device) {
  if (region_options) region_options_ = *region_options;
  SetRateBufferSize(300);
}

bool B200RegionSegmentationUnit::OpenStreams(StreamSet* set) {
  video_stream_idx_ = FindStreamIdx(options_.video_stream_name, set);
  if (video_stream_idx_ < 0) { LOG(ERROR) << "Could not find video stream!\n"; return false; }
  const VideoStream& vid_stream = set->at(video_stream_idx_)->As<VideoStream>();
  frame_width_ = vid_stream.frame_width();
  frame_height_ = vid_stream.frame_height();
  if (vid_stream.pixel_format() != video_framework::PIXEL_FORMAT_BGR24) { LOG(ERROR) << "Expecting video format to be BGR24.\n"; return false; }
  if (!options_.flow_stream_name.empty()) {
    flow_stream_idx_ = FindStreamIdx(options_.flow_stream_name, set);
    if (flow_stream_idx_ < 0) { LOG(ERROR) << "Flow stream specified but not present"; return false; }
  } else {
    flow_stream_idx_ = -1;
  }
  seg_stream_idx_ = FindStreamIdx(options_.segment_stream_name, set);
  if (seg_stream_idx_ < 0) { LOG(ERROR) << "Could not find Segmentation stream!\n"; return false; }
  if (vsb200_device_count() <= 0) {
    LOG(ERROR) << "B200 region segmentation: no sm_100 CUDA device available (this path has no CPU fallback)";
    return false;
  }
  region_options_.use_flow = flow_stream_idx_ >= 0;                      // CreateRegionSegmentation (:303-308)
  region_seg_.reset(new B200RegionSegmentation(region_options_, frame_width_, frame_height_, device_));
  return true;
}

void B200RegionSegmentationUnit::ProcessFrame(FrameSetPtr input, std::list<FrameSetPtr>* output) {
  PointerFrame<SegmentationDesc>* seg_frame = input->at(seg_stream_idx_)->AsMutablePtr<PointerFrame<SegmentationDesc>>();
  const SegmentationDesc* desc = seg_frame->Ptr();
  std::vector<cv::Mat> features;                                         // ExtractFrameSetFeatures (:310-331)
  const VideoFrame& frame = input->at(video_stream_idx_)->As<VideoFrame>();
  cv::Mat image_view;
  frame.MatView(&image_view);
  features.push_back(image_view);
  if (flow_stream_idx_ >= 0) {
    if (num_input_frames_ > 0) features.push_back(input->at(flow_stream_idx_)->As<DenseFlowFrame>().MatViewInterleaved());
    else features.push_back(cv::Mat());
  }
  frame_set_buffer_.push_back(input);
  std::vector<std::unique_ptr<SegmentationDesc>> results;
  region_seg_->ProcessFrame(false, desc, &features, &results);
  seg_frame->release();
  if (options_.free_video_frames) input->at(video_stream_idx_).reset();
  if (flow_stream_idx_ >= 0 && options_.free_flow_frames) input->at(flow_stream_idx_).reset();
  if (!results.empty()) OutputSegmentation(&results, output);
  ++num_input_frames_;
}

bool B200RegionSegmentationUnit::PostProcess(std::list<FrameSetPtr>* append) {
  std::vector<std::unique_ptr<SegmentationDesc>> results;
  if (region_seg_->ProcessFrame(true, nullptr, nullptr, &results) > 0) OutputSegmentation(&results, append);
  return false;
}

void B200RegionSegmentationUnit::OutputSegmentation(std::vector<std::unique_ptr<SegmentationDesc>>* results, std::list<FrameSetPtr>* output) {
  for (size_t k = 0; k < results->size(); ++k) {
    FrameSetPtr frame_set = frame_set_buffer_.front();
    frame_set_buffer_.pop_front();
    const int64_t pts = frame_set->at(seg_stream_idx_)->pts();
    frame_set->at(seg_stream_idx_).reset(new PointerFrame<SegmentationDesc>(std::move((*results)[k]), pts));
    output->push_back(frame_set);
  }
}

}  // namespace segmentation
