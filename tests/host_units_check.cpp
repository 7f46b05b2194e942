// TEST INFRASTRUCTURE: runs the product's two VideoUnits (video_segment_b200/host/b200_segmentation_units.cpp) inside the
// REFERENCE's own video_framework (video_unit.cpp compiled unmodified): memory source -> B200DenseSegmentationUnit ->
// B200RegionSegmentationUnit -> sink, the tree seg_tree_sample builds (seg_tree_sample/seg_tree.cpp:194-240), and writes
// the hierarchical results as flat int32 records (layout of vsb200_region_pop / oracle/ref_hier_wrap.cpp).
// usage: b200_units_check in.bgr out.bin [chunk_size_flag]   (in.bgr: int32 w, h, t, has_flow; t frames BGR24; then, with
// flow, t x [h][w][2] float32)
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include <fstream>
#include <vector>

#include <gflags/gflags.h>
#include <glog/logging.h>

#include "b200_segmentation_units.h"

// the reference defines these in dense_segmentation.cpp / region_segmentation.cpp, which this binary does not link
DEFINE_string(dense_smoothing, "", "");
DEFINE_string(dense_color_dist, "", "");
DEFINE_double(dense_min_region_size, 0, "");
DEFINE_int32(chunk_size, 0, "");
DEFINE_int32(min_region_num, 0, "");
DEFINE_int32(max_region_num, 0, "");
DEFINE_double(level_cutoff_fraction, 0.0, "");
DEFINE_double(small_region_penalizer, -1, "");
DEFINE_int32(chunk_set_size, 0, "");

// DenseFlowFrame's three members used on this path.  The reference defines them in video_framework/flow_reader.cpp next to
// its OpenCV optical-flow unit (cv::DenseOpticalFlow, not available here), so that file cannot be compiled; these are
// the same one-liners (flow_reader.cpp:38-61): a DataFrame of 2 * w * h floats viewed as a CV_32FC2 matrix.
namespace video_framework {
DenseFlowFrame::DenseFlowFrame(int width, int height, bool backward_flow, int64_t pts)
    : DataFrame(&typeid(DenseFlowFrame), 2 * width * height * sizeof(float), pts), width_(width), height_(height), backward_flow_(backward_flow) {}
cv::Mat DenseFlowFrame::MatViewInterleaved() { return cv::Mat(height_, width_, CV_32FC2, mutable_data(), width_ * 2 * sizeof(float)); }
const cv::Mat DenseFlowFrame::MatViewInterleaved() const { return cv::Mat(height_, width_, CV_32FC2, (void*)data(), width_ * 2 * sizeof(float)); }
}  // namespace video_framework

using namespace video_framework;
using segmentation::SegmentationDesc;

namespace {

class MemorySource : public VideoUnit {
 public:
  MemorySource(int w, int h, int t, const uint8_t* frames, const float* flows) : w_(w), h_(h), t_(t), frames_(frames), flows_(flows) {}
  bool OpenStreams(StreamSet* set) override {
    set->push_back(std::shared_ptr<DataStream>(new VideoStream(w_, h_, w_ * 3, 25.f, PIXEL_FORMAT_BGR24, "VideoStream")));
    if (flows_) set->push_back(std::shared_ptr<DataStream>(new DataStream("BackwardFlowStream")));
    return true;
  }
  bool PostProcess(std::list<FrameSetPtr>* append) override {        // a root unit emits its frames here (VideoUnit::NextFrame)
    if (next_ >= t_) return false;
    FrameSetPtr fs(new FrameSet);
    VideoFrame* vf = new VideoFrame(w_, h_, 3, w_ * 3, (int64_t)next_ * 40000);
    memcpy(vf->mutable_data(), frames_ + (size_t)next_ * w_ * h_ * 3, (size_t)w_ * h_ * 3);
    fs->push_back(std::shared_ptr<Frame>(vf));
    if (flows_) {
      DenseFlowFrame* ff = new DenseFlowFrame(w_, h_, true, (int64_t)next_ * 40000);
      memcpy(ff->mutable_data(), flows_ + (size_t)next_ * w_ * h_ * 2, (size_t)w_ * h_ * 8);
      fs->push_back(std::shared_ptr<Frame>(ff));
    }
    append->push_back(fs);
    ++next_;
    return true;
  }
 private:
  int w_, h_, t_, next_ = 0;
  const uint8_t* frames_;
  const float* flows_;
};

class Sink : public VideoUnit {
 public:
  explicit Sink(std::vector<std::vector<int32_t>>* out) : out_(out) {}
  bool OpenStreams(StreamSet* set) override { seg_idx_ = FindStreamIdx("SegmentationStream", set); return seg_idx_ >= 0; }
  void ProcessFrame(FrameSetPtr input, std::list<FrameSetPtr>* output) override {
    const SegmentationDesc& d = input->at(seg_idx_)->As<PointerFrame<SegmentationDesc>>().Ref();
    std::vector<int32_t> f;
    auto bits = [](float v) { int32_t b; memcpy(&b, &v, 4); return b; };
    const int32_t head[8] = {d.frame_width(), d.frame_height(), d.chunk_id(), d.chunk_size(), d.overlap_start(),
                             d.hierarchy_frame_idx(), d.region_size(), d.hierarchy_size()};
    f.insert(f.end(), head, head + 8);
    for (const auto& r : d.region()) {
      f.push_back(r.id());
      f.push_back(r.raster().scan_inter_size());
      for (const auto& s : r.raster().scan_inter()) { f.push_back(s.y()); f.push_back(s.left_x()); f.push_back(s.right_x()); }
      const auto& m = r.shape_moments();
      for (float v : {m.size(), m.mean_x(), m.mean_y(), m.moment_xx(), m.moment_xy(), m.moment_yy()}) f.push_back(bits(v));
    }
    for (const auto& level : d.hierarchy()) {
      f.push_back(level.region_size());
      for (const auto& c : level.region()) {
        f.push_back(c.id()); f.push_back(c.size()); f.push_back(c.parent_id()); f.push_back(c.start_frame()); f.push_back(c.end_frame());
        f.push_back(c.neighbor_id_size()); f.push_back(c.child_id_size());
        for (int k = 0; k < c.neighbor_id_size(); ++k) f.push_back(c.neighbor_id(k));
        for (int k = 0; k < c.child_id_size(); ++k) f.push_back(c.child_id(k));
      }
    }
    out_->push_back(std::move(f));
    output->push_back(input);
  }
 private:
  std::vector<std::vector<int32_t>>* out_;
  int seg_idx_ = -1;
};

}  // namespace

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: %s in.bgr out.bin [chunk_size_flag]\n", argv[0]); return 2; }
  if (argc > 3) FLAGS_chunk_size = atoi(argv[3]);
  std::ifstream in(argv[1], std::ios::binary);
  int32_t hdr[4];
  in.read((char*)hdr, 16);
  const int w = hdr[0], h = hdr[1], t = hdr[2], has_flow = hdr[3];
  std::vector<uint8_t> frames((size_t)w * h * 3 * t);
  in.read((char*)frames.data(), frames.size());
  std::vector<float> flows;
  if (has_flow) { flows.resize((size_t)w * h * 2 * t); in.read((char*)flows.data(), flows.size() * 4); }
  if (!in) { fprintf(stderr, "short input file\n"); return 2; }

  MemorySource source(w, h, t, frames.data(), has_flow ? flows.data() : nullptr);
  segmentation::DenseSegmentationUnitOptions dense_unit_options;
  segmentation::RegionSegmentationUnitOptions region_unit_options;
  if (!has_flow) { dense_unit_options.flow_stream_name.clear(); region_unit_options.flow_stream_name.clear(); }
  segmentation::DenseSegmentationOptions dense_options;
  segmentation::RegionSegmentationOptions region_options;
  region_options.compute_vectorization = false;                     // not built
  segmentation::B200DenseSegmentationUnit dense(dense_unit_options, &dense_options);
  segmentation::B200RegionSegmentationUnit region(region_unit_options, &region_options);
  std::vector<std::vector<int32_t>> records;
  Sink sink(&records);
  dense.AttachTo(&source);
  region.AttachTo(&dense);
  sink.AttachTo(&region);
  if (!source.PrepareAndRun()) { fprintf(stderr, "pipeline could not be opened (no CPU fallback: an sm_100 device is required)\n"); return 1; }
  FILE* out = fopen(argv[2], "wb");
  if (!out) return 2;
  const int64_t n = (int64_t)records.size();
  fwrite(&n, 8, 1, out);
  for (const auto& r : records) { const int64_t len = (int64_t)r.size(); fwrite(&len, 8, 1, out); fwrite(r.data(), 4, r.size(), out); }
  fclose(out);
  printf("b200_units_check: %lld hierarchical frame results\n", (long long)n);
  return 0;
}
