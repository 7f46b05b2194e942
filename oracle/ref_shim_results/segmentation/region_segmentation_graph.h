// TEST INFRASTRUCTURE ONLY.  Shadows the reference's segmentation/region_segmentation_graph.h for the translation unit
// segmentation/segmentation.cpp when it is compiled into oracle/_ref/libref_results.so.  The real header does not
// compile with GCC 13 (its nested EdgeHasher is used as an unordered_map hasher inside the still-incomplete enclosing
// class; libstdc++ then sees no default constructor), and the hierarchical stage is not what this library pins:
// only the over-segmentation half of segmentation.cpp is exercised (RunOverSegmentation, AssignUniqueRegionIds,
// RetrieveSegmentation3D ...).  This declares the interface segmentation.cpp calls; every method aborts.
#ifndef VSO_REF_SHIM_REGION_SEGMENTATION_GRAPH_H_
#define VSO_REF_SHIM_REGION_SEGMENTATION_GRAPH_H_
#include <cstdlib>
#include <unordered_map>
#include <vector>
#include "base/base.h"
#include "segmentation/segmentation_common.h"
#include "segmentation/segmentation_graph.h"
namespace segmentation {
typedef std::unordered_map<int, std::vector<int>> Skeleton;
class RegionAgglomerationGraph {
 public:
  typedef std::unordered_map<long long, float> EdgeWeightMap;
  RegionAgglomerationGraph(float, int, const RegionDistance*) { std::abort(); }
  void AddRegionEdges(const RegionInfoList&, const EdgeWeightMap*) { std::abort(); }
  void AddRegionEdgesConstrained(const RegionInfoList&, const EdgeWeightMap*, const std::vector<int>&, const Skeleton&) { std::abort(); }
  int SegmentGraph(bool, float) { std::abort(); }
  void ObtainSegmentationResult(RegionInfoList*, RegionInfoList*, EdgeWeightMap*) { std::abort(); }
};
}  // namespace segmentation
#endif
