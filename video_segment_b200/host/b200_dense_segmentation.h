// b200_dense_segmentation.h -- C++ host side above the C ABI (include/vsb200.h), written against the REFERENCE's own
// headers: a class with the constructor and ProcessFrame signature of segmentation::DenseSegmentation
// (segmentation/dense_segmentation.h:98-131) that runs the over-segmentation on a B200 and hands back the same
// std::unique_ptr<SegmentationDesc> objects.  DenseSegmentationUnit (segmentation/segmentation_unit.cpp:58-178) holds
// its engine as `std::unique_ptr<DenseSegmentation> dense_seg_` and only ever calls ProcessFrame / ChunkSize on it, so
// a maintainer swaps the member's type (or makes ProcessFrame virtual and derives): see INTEGRATION.md section 2.
//
// Build: this file needs the reference tree on the include path ("segmentation/dense_segmentation.h",
// "segment_util/segmentation.pb.h", <opencv2/core/core.hpp>) and links libvsb200.so.  In this repository it is compiled
// by `__graft_entry__.build()` where /root/reference is mounted, against the library stand-ins of oracle/ref_shim
// (OpenCV / glog / protobuf are not installed in the image); it uses nothing but cv::Mat's data / step / rows / cols
// and the protoc-generated accessors, so it builds unchanged against the real libraries.
// There is no CPU fallback: without an sm_100 device the constructor's first ProcessFrame fails loudly.
#ifndef VSB200_HOST_B200_DENSE_SEGMENTATION_H_
#define VSB200_HOST_B200_DENSE_SEGMENTATION_H_

#include <memory>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

#include "segmentation/dense_segmentation.h"   // DenseSegmentationOptions, SegmentationDesc (reference headers)
#include "vsb200.h"

namespace segmentation {

// Fills a SegmentationDesc from the arrays of one frame result: the fields Segmentation::RetrieveSegmentation3D
// (segmentation.cpp:458-533), AddRegion2DToSegmentationDesc / AddCompoundRegionToSegmentationDesc (:671-773) and
// DenseSegmentation::SegmentAndOutputChunk (dense_segmentation.cpp:385-387) set, with the same presence bits.
void FrameResultToSegmentationDesc(const vsb200_frame_result& r, SegmentationDesc* desc);

class B200DenseSegmentation {
 public:
  // Same arguments as DenseSegmentation(options, frame_width, frame_height) plus the CUDA device ordinal.
  B200DenseSegmentation(const DenseSegmentationOptions& options, int frame_width, int frame_height, int cuda_device = 0);
  ~B200DenseSegmentation();
  B200DenseSegmentation(const B200DenseSegmentation&) = delete;
  B200DenseSegmentation& operator=(const B200DenseSegmentation&) = delete;

  // Same contract as DenseSegmentation::ProcessFrame (dense_segmentation.cpp:108-162): features->at(0) is the 8-bit
  // BGR frame, `flow` is non-null on every call iff a flow stream exists (an empty cv::Mat for the first frame);
  // flush = true (features / flow null) drains the stream.  Results arrive in input order; returns results->size().
  // Setup and device errors abort through CHECK like the reference's own precondition failures, with
  // vsb200_last_error() in the message.
  int ProcessFrame(bool flush, const std::vector<cv::Mat>* features, const cv::Mat* flow,
                   std::vector<std::unique_ptr<SegmentationDesc>>* results);

  int ChunkSize() const { return options_.chunk_size; }

  // Not part of the reference interface: kernel launches so far, for callers that want to assert the GPU ran.
  long long KernelLaunches() const;

 private:
  void Pop(int n_ready, std::vector<std::unique_ptr<SegmentationDesc>>* results);

  DenseSegmentationOptions options_;
  int frame_width_ = 0;
  int frame_height_ = 0;
  int device_ = 0;
  int input_frames_ = 0;
  vsb200_dense* handle_ = nullptr;
};

}  // namespace segmentation

#endif  // VSB200_HOST_B200_DENSE_SEGMENTATION_H_
