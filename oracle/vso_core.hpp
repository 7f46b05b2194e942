// vso_core.hpp -- internal types of the CPU oracle (test infrastructure only).
// See vso.h for the scope statement.  Citations are reference file:line.
#ifndef VSO_CORE_HPP_
#define VSO_CORE_HPP_

#include <array>
#include <cstdint>
#include <memory>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

namespace vso {

// ---------------------------------------------------------------------------
// Result containers (stand-ins for the protobuf messages, segmentation.proto:55-172)
// ---------------------------------------------------------------------------
struct ScanInterval { int y, left_x, right_x; };
typedef std::vector<ScanInterval> Rasterization;
// segment_util/segmentation_util.h:230
typedef std::vector<std::pair<int, std::shared_ptr<Rasterization>>> Rasterization3D;

struct ShapeMoments { float size = 0, mean_x = 0, mean_y = 0, moment_xx = 0, moment_xy = 0, moment_yy = 0; };

struct Point2f { float x = 0, y = 0; };

// segment_util/segmentation_util.h:137-150
struct ShapeDescriptor {
  Point2f center;
  float mag_major = 0;
  float mag_minor = 0;
  Point2f dir_major{1.0f, 0.0f};
  Point2f dir_minor{0.0f, 1.0f};
  int size = 0;
};

struct Region2D { int id = -1; Rasterization raster; ShapeMoments moments; };
struct CompoundRegion { int id = -1, size = 0; std::vector<int> neighbor_id; int start_frame = 0, end_frame = 0; };

struct SegDesc {
  int frame_width = 0, frame_height = 0, chunk_id = -1, connectedness = 1;
  int chunk_size = 0, overlap_start = 0, hierarchy_frame_idx = 0;
  std::vector<Region2D> region;
  bool has_hierarchy = false;
  std::vector<CompoundRegion> hierarchy0;
};

// segmentation/segmentation_common.h:39-116 (over-segmentation subset)
struct RegionInformation {
  int index = -1;
  int size = 0;
  bool flagged_for_removal = false;
  std::vector<int> neighbor_idx;
  std::unique_ptr<Rasterization3D> raster;
  int constrained_id = -1;
  int region_id = -1;
};
typedef std::unordered_map<int, RegionInformation*> RegionInfoPtrMap;
typedef std::vector<std::unique_ptr<RegionInformation>> RegionInfoList;

// ---------------------------------------------------------------------------
// raster / shape helpers (segment_util/segmentation_util.cpp)
// ---------------------------------------------------------------------------
void MergeRasterization(const Rasterization& lhs, const Rasterization& rhs, Rasterization* out);
int RasterizationArea(const Rasterization& r);
void ShapeMomentsFromRasterization(const Rasterization& r, ShapeMoments* m);
bool GetShapeDescriptorFromShapeMoment(const ShapeMoments& m, ShapeDescriptor* d);
int ConnectedComponentsN4(const Rasterization& raster, std::vector<Rasterization>* components);
void SegDescToIdImage(const SegDesc& seg, int width, int* id_image /* [h*w] */);
template <class T> bool InsertSortedUniquely(const T& t, std::vector<T>* array);

// ---------------------------------------------------------------------------
// image preprocessing (dense_segmentation.cpp:164-198, imagefilter/image_filter.cpp)
// ---------------------------------------------------------------------------
void ConvertU8ToF32(const uint8_t* bgr, int w, int h, int row_stride, float* out);
void BilateralFilter(const float* in, int w, int h, float sigma_space, float sigma_color,
                     float* out, int num_threads, float* lut_out, float* scale_out);

// ---------------------------------------------------------------------------
// pixel distances (segmentation/pixel_distance.h:141-157)
// ---------------------------------------------------------------------------
float ColorDiff3L1(const float* a, const float* b);
float ColorDiff3L2(const float* a, const float* b);

// ---------------------------------------------------------------------------
// The dense graph = DenseSegmentationGraph<Distance, ColorMeanDescriptorTraits>
// on top of FastSegmentationGraph (segmentation_graph.h, dense_segmentation_graph.h)
// ---------------------------------------------------------------------------
class DenseGraph {
 public:
  DenseGraph(int w, int h, int max_frames, bool l1, bool parallel_build);
  ~DenseGraph();

  void AddNodesAndSpatialEdges(const float* img);                         // :906-916
  void AddNodesAndSpatialEdgesConstrained(const float* img, const SegDesc& d);  // :918-930
  void AddVirtualNodesConstrained(const SegDesc& d);                      // :327-367
  void AddTemporalEdges(const float* curr, const float* prev);            // :932-941
  void AddTemporalFlowEdges(const float* curr, const float* prev, const float* flow);
  void AddTemporalVirtualEdges();                                         // :369-380
  void AddTemporalFlowVirtualEdges(const float* flow);                    // :382-395
  void FinishBuildingGraph();                                             // :397-404
  void SegmentFullGraph(int min_region_size, bool force_constraints);     // :421-423
  void ObtainResults(RegionInfoList* list, RegionInfoPtrMap* map,
                     const std::vector<const float*>* flows,
                     bool enforce_n4, bool enforce_spatial_connectedness); // :467-579
  void DetermineNeighborIds(RegionInfoList* list, RegionInfoPtrMap* map); // segmentation_graph.h:466-496

  int num_frames() const { return num_frames_; }
  // test hook: a graph of `frames` slices whose union-find is the given label volume (no edges), ready for ObtainResults
  void TestLoadLabels(const int32_t* labels, int frames);
  // debug taps
  std::vector<int32_t> node_labels_after_flatten;   // [num_frames * N]
  std::vector<int32_t> id_images_after_n4;          // [num_frames * N], -1 on virtual slices
  int64_t merge_stats[3] = {0, 0, 0};

 private:
  struct Edge { int region_1, region_2; };
  typedef std::vector<Edge> EdgeList;
  struct Region {                                     // segmentation_graph.h:249-266
    int my_id = -1;
    int sz = 0;
    int constraint_id = -1;
    bool region_finalized = false;
    float descriptor[3];
  };

  // --- FastSegmentationGraph part
  inline void AddEdge(int r1, int r2, float weight, int bucket_list);
  Region* GetRegion(int id);
  Region* MergeRegions(Region* rep_1, Region* rep_2);
  float DescriptorDistance(const float* lhs, const float* rhs, float edge_distance) const;
  void SegmentGraph(int min_region_size, bool force_constraints);
  void SimulateStage0(int64_t* stats);   // simulation of the product's force-bucket shortcut (tests only), see vso_graph.cpp
  void MergeConstrainedRegions();
  void FlattenUnionFind(bool separate_representatives);
  RegionInformation* GetCreateRegionInformation(const Region& rep, RegionInfoList* list,
                                                RegionInfoPtrMap* map);

  // --- DenseSegmentationGraph part
  void AddNodesWithDescriptors(const float* img, const int* constraint_ids);
  void AddSpatialEdgesImpl(const float* img, int frame_idx);
  void AddTemporalEdgesImpl(const float* curr, const float* prev, const float* flow,
                            bool constant, int frame_idx);
  void AddIntervalToRasterization(int frame, int y, int lx, int rx, int region_id,
                                  RegionInfoList* list, RegionInfoPtrMap* map);
  void EnforceN4Connectivity(int* id_image_with_border, std::unordered_map<int, int>* adj);
  void EnforceSpatialConnectedness(RegionInfoList* list, RegionInfoPtrMap* map,
                                   const std::vector<const float*>* flows,
                                   std::unordered_map<int, int>* adj);
  float PixelDistance(const float* a, const float* b) const {
    return l1_ ? ColorDiff3L1(a, b) : ColorDiff3L2(a, b);
  }

  int w_, h_, max_frames_;
  bool l1_, parallel_build_;
  int num_frames_ = 0;
  int num_buckets_ = 2048;
  float scale_ = 1.0f;
  float force_merge_weight_;
  std::vector<Region> regions_;
  std::vector<std::pair<int, int>> virtual_nodes_;
  std::vector<std::vector<EdgeList>> bucket_lists_;   // [list][bucket]
  bool flattened_ = false;
  int max_region_id_ = 0;
  std::vector<int> region_ids_;       // (h+2) x (w+2) id image with border
  std::vector<int> virtual_slices_;
  std::vector<std::thread> add_edges_tasks_;
};

}  // namespace vso
#endif
