// TEST INFRASTRUCTURE ONLY.  Drives the REFERENCE's own FastSegmentationGraph (segmentation/segmentation_graph.h:
// AddEdge bucketing, SegmentGraph, GetRegion, MergeRegions, MergeConstrainedRegions), ColorMeanDescriptorTraits and
// the Spatial / TemporalCvMatDistance walkers with ColorDiff3L2 / L1 (segmentation/pixel_distance.h), all compiled
// unmodified from /root/reference (`make -C oracle _ref`; glog / cv::Mat / two type-only headers are stand-ins under
// oracle/ref_shim).  The graph CONSTRUCTION below restates DenseSegmentationGraph (dense_segmentation_graph.h, which
// needs protobuf and cannot be compiled): AddNodesWithDescriptors :1180-1199, AddSpatialEdgesImpl :956-1000,
// GetLocalEdges + AddTemporalEdgesImpl :1002-1097, SegmentFullGraph :418-423 -- same calls in the same order.
// Used to pin the oracle's restatement of the merge (oracle/vso_graph.cpp) against the reference's algorithm.
#include <stdint.h>

#include <vector>

#include "segmentation/pixel_distance.h"
#include "segmentation/segmentation_graph.h"

namespace {
using namespace segmentation;

class RefGraph : public FastSegmentationGraph<ColorMeanDescriptorTraits> {
 public:
  typedef FastSegmentationGraph<ColorMeanDescriptorTraits> Base;
  RefGraph(float force_merge_weight, int max_frames)
      : Base(1.0, 2048, ColorMeanDescriptorTraits(force_merge_weight), 2 * max_frames - 1) {}
  using Base::AddEdge;
  using Base::AddRegionWithDescriptor;
  using Base::SegmentGraph;
  int Root(int node) { return GetRegion(node)->my_id; }
  int NumNodes() const { return (int)regions_.size(); }
};

template <class Spatial, class Temporal>
void build_and_segment(const float* frames, int w, int h, int t, float force_merge_weight, int min_region_size,
                       int force_constraints, int32_t* labels_out, float* spatial_w_out, float* temporal_w_out) {
  RefGraph g(force_merge_weight, t);
  const int n = w * h;
  for (int f = 0; f < t; ++f) {
    const cv::Mat curr(h, w, CV_32FC3, (void*)(frames + (size_t)f * n * 3), (size_t)w * 3 * sizeof(float));
    {
      Spatial distance(curr);                                  // AddNodesWithDescriptors
      const int base_idx = f * n;
      for (int i = 0; i < h; ++i) {
        const int row_idx = base_idx + i * w;
        distance.MoveAnchorTo(0, i);
        float descriptor[3];
        for (int j = 0; j < w; ++j, distance.IncrementAnchor()) {
          distance.SetPixelDescriptor(descriptor);
          g.AddRegionWithDescriptor(row_idx + j, 1, -1, descriptor);
        }
      }
    }
    {
      Spatial distance(curr);                                  // AddSpatialEdgesImpl
      const int base_idx = f * n, bucket_list_idx = 2 * f;
      float* wout = spatial_w_out ? spatial_w_out + (size_t)f * n * 4 : nullptr;
      for (int i = 0, end_y = h - 1, cur_idx = base_idx; i <= end_y; ++i) {
        distance.MoveAnchorTo(0, i);
        distance.MoveTestAnchorTo(0, i);
        for (int j = 0, end_x = w - 1; j <= end_x; ++j, ++cur_idx, distance.IncrementAnchor(), distance.IncrementTestAnchor()) {
          float d[4] = {-1.f, -1.f, -1.f, -1.f};
          if (j < end_x) g.AddEdge(cur_idx, cur_idx + 1, d[0] = distance.PixelDistance(1, 0), bucket_list_idx);
          if (i < end_y) {
            g.AddEdge(cur_idx, cur_idx + w, d[1] = distance.PixelDistance(0, 1), bucket_list_idx);
            if (j > 0) g.AddEdge(cur_idx, cur_idx + w - 1, d[2] = distance.PixelDistance(-1, 1), bucket_list_idx);
            if (j < end_x) g.AddEdge(cur_idx, cur_idx + w + 1, d[3] = distance.PixelDistance(1, 1), bucket_list_idx);
          }
          if (wout) for (int k = 0; k < 4; ++k) wout[(size_t)(cur_idx - base_idx) * 4 + k] = d[k];
        }
      }
    }
    if (f > 0) {
      const cv::Mat prev(h, w, CV_32FC3, (void*)(frames + (size_t)(f - 1) * n * 3), (size_t)w * 3 * sizeof(float));
      Temporal distance(curr, prev);                           // AddTemporalEdgesImpl (frame_idx = f + 1)
      const int base_idx = f * n, bucket_list_idx = 2 * f - 1;
      float* wout = temporal_w_out ? temporal_w_out + (size_t)f * n * 9 : nullptr;
      int curr_idx = base_idx;
      for (int i = 0; i < h; ++i) {
        distance.MoveAnchorTo(0, i);
        distance.MoveTestAnchorTo(0, i);
        for (int j = 0; j < w; ++j, ++curr_idx, distance.IncrementAnchor(), distance.IncrementTestAnchor()) {
          const int prev_idx = curr_idx - n;
          float d[9];
          for (int k = 0; k < 9; ++k) d[k] = -1.f;
          for (int dy = -1; dy <= 1; ++dy) {                   // GetLocalEdges order: TL,T,TR,L,C,R,BL,B,BR
            if (i + dy < 0 || i + dy >= h) continue;
            for (int dx = -1; dx <= 1; ++dx) {
              if (j + dx < 0 || j + dx >= w) continue;
              const float wgt = distance.PixelDistance(dx, dy);
              d[(dy + 1) * 3 + dx + 1] = wgt;
              g.AddEdge(curr_idx, prev_idx + dy * w + dx, wgt, bucket_list_idx);
            }
          }
          if (wout) for (int k = 0; k < 9; ++k) wout[(size_t)(curr_idx - base_idx) * 9 + k] = d[k];
        }
      }
    }
  }
  g.SegmentGraph(min_region_size, force_constraints != 0, nullptr);      // SegmentFullGraph
  for (int i = 0; i < n * t; ++i) labels_out[i] = g.Root(i);
}
}  // namespace

extern "C" {

// One unconstrained chunk: smoothed frames [t][h][w][3] -> representative node id per voxel; optionally the edge
// weights as the reference's distance walkers produce them ([t][h][w][4] spatial R,B,BL,BR; [t][h][w][9] temporal,
// -1 = no edge; frame 0 has no temporal edges).  force_merge_weight as in dense_segmentation.cpp:259-264.
int ref_segment_chunk_labels(const float* frames, int w, int h, int t, int l1, int min_region_size, int force_constraints,
                             int32_t* labels_out, float* spatial_w_out, float* temporal_w_out) {
  if (l1) build_and_segment<SpatialCvMatDistance3L1, TemporalCvMatDistance3L1>(frames, w, h, t, 0.002f, min_region_size, force_constraints,
                                                                              labels_out, spatial_w_out, temporal_w_out);
  else build_and_segment<SpatialCvMatDistance3L2, TemporalCvMatDistance3L2>(frames, w, h, t, 0.001f, min_region_size, force_constraints,
                                                                           labels_out, spatial_w_out, temporal_w_out);
  return 0;
}

// FastSegmentationGraph::AddEdge bucketing (segmentation_graph.h:158-162) as the constructor scales it (:336)
int ref_bucket_index(float weight) {
  const float scale = 2048 / (1.0f + 1e-6f);
  return (int)(std::min<float>(2048, weight * scale));
}

}  // extern "C"
