// b200_segmentation_units.h -- the two plug points named by the reference's pipeline, on the B200 path:
//   segmentation::B200DenseSegmentationUnit  in place of DenseSegmentationUnit   (segmentation/segmentation_unit.h:63-124)
//   segmentation::B200RegionSegmentationUnit in place of RegionSegmentationUnit  (segmentation/segmentation_unit.h:126-190)
// plus segmentation::B200RegionSegmentation, a class with RegionSegmentation's constructor and ProcessFrame signature
// (segmentation/region_segmentation.h:100-146) over vsb200_region_*.  Both units are video_framework::VideoUnit
// subclasses with the reference units' option structs, stream contract (video [+ flow] in, SegmentationStream added /
// replaced), frame-set buffering and log markers, so that seg_tree_sample/seg_tree.cpp:194-240 switches over by
// changing two type names.
//
// Written against the reference's own headers (video_framework/video_unit.h, video_framework/flow_reader.h,
// segmentation/segmentation_unit.h for the option structs).  In this repository they are compiled and run by
// `make -C oracle _ref` -> oracle/_ref/b200_units_check against the reference tree (video_unit.cpp, flow_reader.cpp
// compiled unmodified) plus stand-ins for the libraries the image lacks (oracle/ref_shim: boost::circular_buffer /
// mutex / posix_time, glog, OpenCV core, protobuf accessors).
#ifndef VSB200_HOST_B200_SEGMENTATION_UNITS_H_
#define VSB200_HOST_B200_SEGMENTATION_UNITS_H_

#include <list>
#include <memory>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

#include "b200_dense_segmentation.h"
#include "segmentation/segmentation_unit.h"      // DenseSegmentationUnitOptions, RegionSegmentationUnitOptions (reference)
#include "video_framework/flow_reader.h"          // DenseFlowFrame
#include "video_framework/video_unit.h"
#include "vsb200.h"

namespace segmentation {

// RegionSegmentation::ProcessFrame(flush, desc, features, results) on the device path.
class B200RegionSegmentation {
 public:
  B200RegionSegmentation(const RegionSegmentationOptions& options, int frame_width, int frame_height, int cuda_device = 0);
  ~B200RegionSegmentation();
  B200RegionSegmentation(const B200RegionSegmentation&) = delete;
  B200RegionSegmentation& operator=(const B200RegionSegmentation&) = delete;

  // features: [BGR frame] or [BGR frame, flow (empty Mat on the first frame)], as RegionSegmentationUnit::
  // ExtractFrameSetFeatures builds them (segmentation_unit.cpp:310-331).
  int ProcessFrame(bool flush, const SegmentationDesc* desc, const std::vector<cv::Mat>* features,
                   std::vector<std::unique_ptr<SegmentationDesc>>* results);
  long long KernelLaunches() const;

 private:
  void Pop(int n_ready, std::vector<std::unique_ptr<SegmentationDesc>>* results);
  RegionSegmentationOptions options_;
  int frame_width_, frame_height_, device_;
  vsb200_region* handle_ = nullptr;
};

// One hierarchical result record of vsb200_region_pop -> SegmentationDesc (field order of the record = the message's).
void RegionRecordToSegmentationDesc(const int32_t* record, long long words, SegmentationDesc* desc);

class B200DenseSegmentationUnit : public video_framework::VideoUnit {
 public:
  B200DenseSegmentationUnit(const DenseSegmentationUnitOptions& options, const DenseSegmentationOptions* dense_seg_options,
                            int cuda_device = 0);
  ~B200DenseSegmentationUnit() override = default;
  bool OpenStreams(video_framework::StreamSet* set) override;
  void ProcessFrame(video_framework::FrameSetPtr input, std::list<video_framework::FrameSetPtr>* output) override;
  bool PostProcess(std::list<video_framework::FrameSetPtr>* append) override;

 private:
  void OutputSegmentation(std::vector<std::unique_ptr<SegmentationDesc>>* results, std::list<video_framework::FrameSetPtr>* output);
  DenseSegmentationUnitOptions options_;
  DenseSegmentationOptions dense_seg_options_;
  std::unique_ptr<B200DenseSegmentation> dense_seg_;
  int device_ = 0, video_stream_idx_ = -1, flow_stream_idx_ = -1;
  int frame_width_ = 0, frame_height_ = 0, input_frames_ = 0, output_frames_ = 0;
  std::list<video_framework::FrameSetPtr> frame_set_buffer_;
};

class B200RegionSegmentationUnit : public video_framework::VideoUnit {
 public:
  B200RegionSegmentationUnit(const RegionSegmentationUnitOptions& options, const RegionSegmentationOptions* region_options,
                             int cuda_device = 0);
  ~B200RegionSegmentationUnit() override = default;
  bool OpenStreams(video_framework::StreamSet* set) override;
  void ProcessFrame(video_framework::FrameSetPtr input, std::list<video_framework::FrameSetPtr>* output) override;
  bool PostProcess(std::list<video_framework::FrameSetPtr>* append) override;

 private:
  void OutputSegmentation(std::vector<std::unique_ptr<SegmentationDesc>>* results, std::list<video_framework::FrameSetPtr>* output);
  RegionSegmentationUnitOptions options_;
  RegionSegmentationOptions region_options_;
  std::unique_ptr<B200RegionSegmentation> region_seg_;
  int device_ = 0, video_stream_idx_ = -1, flow_stream_idx_ = -1, seg_stream_idx_ = -1;
  int frame_width_ = 0, frame_height_ = 0, num_input_frames_ = 0;
  std::list<video_framework::FrameSetPtr> frame_set_buffer_;
};

}  // namespace segmentation

#endif  // VSB200_HOST_B200_SEGMENTATION_UNITS_H_
