// TEST INFRASTRUCTURE ONLY.  Shadows the reference's segment_util/segmentation_util.h (protobuf + OpenCV) with the
// few type names segmentation/segmentation_common.h needs to DECLARE RegionInformation, so that the reference's
// segmentation/segmentation_graph.h (FastSegmentationGraph) compiles unmodified into oracle/_ref.
#ifndef VSO_REF_SHIM_SEGMENTATION_UTIL_H_
#define VSO_REF_SHIM_SEGMENTATION_UTIL_H_
#include <memory>
#include <utility>
#include <vector>
namespace segmentation {
struct Rasterization {};
struct Rasterization3D {};
struct RegionFeatures {};
struct SegmentationDesc {};
struct ShapeMoments {};
}  // namespace segmentation
#endif
