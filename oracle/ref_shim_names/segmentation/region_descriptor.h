// TEST INFRASTRUCTURE ONLY.  Shadows the reference's segmentation/region_descriptor.h (OpenCV) with the type names
// segmentation/segmentation_common.h needs to declare RegionInformation (see segment_util/segmentation_util.h here).
#ifndef VSO_REF_SHIM_REGION_DESCRIPTOR_H_
#define VSO_REF_SHIM_REGION_DESCRIPTOR_H_
#include <memory>
#include <vector>
namespace segmentation {
class RegionDescriptor { public: virtual ~RegionDescriptor() {} };
class RegionDescriptorUpdater { public: virtual ~RegionDescriptorUpdater() {} };
typedef std::vector<std::shared_ptr<RegionDescriptorUpdater>> DescriptorUpdaterList;
}  // namespace segmentation
#endif
