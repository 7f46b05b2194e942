// vso_graph.cpp -- CPU ORACLE (test infrastructure): the chunk graph.
// Restates FastSegmentationGraph<ColorMeanDescriptorTraits>
// (segmentation/segmentation_graph.h) and DenseSegmentationGraph
// (segmentation/dense_segmentation_graph.h, .cpp) for the default distances
// (segmentation/pixel_distance.h:141-157,469-521).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>
#include <numeric>
#include <unordered_map>

#include "vso_core.hpp"

namespace vso {

// pixel_distance.h:141-148.  The reference writes unqualified fabs() after including only <cmath>: with GCC /
// libstdc++ that is ::fabs(double), so the three absolute differences are summed and scaled in DOUBLE and rounded to
// float once on return (found by compiling the reference's own header into oracle/_ref; float arithmetic differs in
// the last bit for about one weight in five).  ColorDiff3L2's unqualified sqrt() is ::sqrt(double) in the same way,
// which rounds like sqrtf.
float ColorDiff3L1(const float* p1, const float* p2) {
  const float diff_1 = p1[0] - p2[0];
  const float diff_2 = p1[1] - p2[1];
  const float diff_3 = p1[2] - p2[2];
  return (float)((std::fabs((double)diff_1) + std::fabs((double)diff_2) + std::fabs((double)diff_3)) * (double)(1.0f / 3.0f));
}

// pixel_distance.h:150-157
float ColorDiff3L2(const float* p1, const float* p2) {
  const float diff_1 = p1[0] - p2[0];
  const float diff_2 = p1[1] - p2[1];
  const float diff_3 = p1[2] - p2[2];
  return std::sqrt((diff_1 * diff_1 + diff_2 * diff_2 + diff_3 * diff_3) * (1.0f / 3.0f));
}

// dense_segmentation_graph.h:290-311 + segmentation_graph.h:322-337;
// traits: dense_segmentation.cpp:253-266 (force merge weight 0.001f for L2, 0.002f for L1).
DenseGraph::DenseGraph(int w, int h, int max_frames, bool l1, bool parallel_build)
    : w_(w), h_(h), max_frames_(max_frames), l1_(l1), parallel_build_(parallel_build) {
  force_merge_weight_ = l1 ? 0.002f : 0.001f;
  const float max_weight = 1.0;
  scale_ = num_buckets_ / (max_weight + 1e-6f);
  bucket_lists_.resize(2 * max_frames - 1);
  for (auto& bl : bucket_lists_) bl.resize(num_buckets_ + 1);
  region_ids_.assign((size_t)(h + 2) * (w + 2), 0);
  // InitializeGraph (:313-324): reserve only.
  regions_.reserve((size_t)((double)w * h * max_frames * 1.02));
}

DenseGraph::~DenseGraph() { FinishBuildingGraph(); }

// segmentation_graph.h:158-162
inline void DenseGraph::AddEdge(int r1, int r2, float weight, int bucket_list) {
  const int bucket_index = (int)(std::min<float>(num_buckets_, weight * scale_));
  bucket_lists_[bucket_list][bucket_index].push_back(Edge{r1, r2});
}

// segmentation_graph.h:651-669 (recursive find with full path compression; written
// iteratively -- identical resulting parent pointers).
DenseGraph::Region* DenseGraph::GetRegion(int id) {
  int root = id;
  while (regions_[root].my_id != root) root = regions_[root].my_id;
  int cur = id;
  while (regions_[cur].my_id != root) {
    const int next = regions_[cur].my_id;
    regions_[cur].my_id = root;
    cur = next;
  }
  return &regions_[root];
}

// segmentation_graph.h:671-701 with ColorMeanDescriptorTraits::MergeDescriptor
// (pixel_distance.h:494-504).
DenseGraph::Region* DenseGraph::MergeRegions(Region* rep_1, Region* rep_2) {
  Region* merged;
  Region* other;
  if (rep_1->sz > rep_2->sz) { merged = rep_1; other = rep_2; }
  else                       { merged = rep_2; other = rep_1; }
  {
    const int lhs_size = other->sz, rhs_size = merged->sz;
    const float denom = 1.0f / (lhs_size + rhs_size);
    const float a = lhs_size * denom;
    const float b = rhs_size * denom;
    merged->descriptor[0] = a * other->descriptor[0] + b * merged->descriptor[0];
    merged->descriptor[1] = a * other->descriptor[1] + b * merged->descriptor[1];
    merged->descriptor[2] = a * other->descriptor[2] + b * merged->descriptor[2];
  }
  merged->sz += other->sz;
  merged->constraint_id = std::max(rep_1->constraint_id, rep_2->constraint_id);
  other->my_id = merged->my_id;
  return merged;
}

// pixel_distance.h:478-491
float DenseGraph::DescriptorDistance(const float* lhs, const float* rhs, float edge_distance) const {
  const float diff_1 = lhs[0] - rhs[0];
  const float diff_2 = lhs[1] - rhs[1];
  const float diff_3 = lhs[2] - rhs[2];
  const float dist = std::sqrt((diff_1 * diff_1 + diff_2 * diff_2 + diff_3 * diff_3) * (1.0f / 3.0f));
  if (edge_distance < force_merge_weight_ && dist < 0.2) return 0.0f;
  return dist;
}

// ---------------------------------------------------------------------------------------------
// NOT part of the restatement: a CPU simulation of a candidate "stage 0" shortcut for the product's merge
// (development only, switched on with VSO_SIM_STAGE0=1 by tools/sim_stage0.py; measured at 1080p: one
// percolating component holds 32 M of the 41 M nodes, so the shortcut was not built).  Claim under test: in the force-merge buckets (bucket * inv_scale < force_merge_weight)
// an edge between two different un-finalised regions always merges while the two means are closer
// than 0.2; every region mean is a convex combination of the pixel colours of its connected
// component (edges of the force buckets), so a component whose pixel-colour box has a diagonal
// below 0.2 and that carries at most one constraint id merges completely whatever the order.
// Such "safe" components are merged up front (exact size-weighted mean); all the others are left
// untouched for the ordered scan.
// ---------------------------------------------------------------------------------------------
void DenseGraph::SimulateStage0(int64_t* stats) {
  const float inv_scale = 1.0 / scale_;
  const int n = (int)regions_.size();
  std::vector<int> cc(n);
  std::iota(cc.begin(), cc.end(), 0);
  auto find = [&](int x) { while (cc[x] != x) { cc[x] = cc[cc[x]]; x = cc[x]; } return x; };
  int n_force = 0;
  while (n_force < num_buckets_ && n_force * inv_scale < force_merge_weight_) ++n_force;
  std::vector<char> touched(n, 0);
  for (int b = 0; b < n_force; ++b)
    for (auto& bl : bucket_lists_)
      for (const auto& e : bl[b]) {
        touched[e.region_1] = touched[e.region_2] = 1;
        const int a = find(e.region_1), c = find(e.region_2);
        if (a != c) cc[std::max(a, c)] = std::min(a, c);
      }
  struct Acc { float mn[3], mx[3]; double s[3]; int64_t sz; int con; bool multi; };
  std::unordered_map<int, Acc> acc;
  for (int i = 0; i < n; ++i) {
    if (!touched[i]) continue;
    const int r = find(i);
    auto it = acc.find(r);
    if (it == acc.end()) {
      Acc a;
      for (int k = 0; k < 3; ++k) { a.mn[k] = 1e30f; a.mx[k] = -1e30f; a.s[k] = 0; }
      a.sz = 0; a.con = -1; a.multi = false;
      it = acc.insert(std::make_pair(r, a)).first;
    }
    Acc& a = it->second;
    const Region& R = regions_[i];
    for (int k = 0; k < 3; ++k) {
      a.mn[k] = std::min(a.mn[k], R.descriptor[k]);
      a.mx[k] = std::max(a.mx[k], R.descriptor[k]);
      a.s[k] += (double)R.descriptor[k] * R.sz;
    }
    a.sz += R.sz;
    if (R.constraint_id >= 0) {
      if (a.con >= 0 && a.con != R.constraint_id) a.multi = true;
      a.con = std::max(a.con, R.constraint_id);
    }
  }
  const float thr = 0.2f * 0.999f - 2e-5f;
  int64_t safe_nodes = 0, unsafe_nodes = 0, unsafe_comps = 0, safe_comps = 0;
  for (int i = 0; i < n; ++i) {
    if (!touched[i]) continue;
    const int r = find(i);
    const Acc& a = acc[r];
    const float dx = a.mx[0] - a.mn[0], dy = a.mx[1] - a.mn[1], dz = a.mx[2] - a.mn[2];
    const float diam = std::sqrt((dx * dx + dy * dy + dz * dz) * (1.0f / 3.0f));
    const bool safe = !a.multi && diam < thr;
    if (i == r) { if (safe) ++safe_comps; else ++unsafe_comps; }
    if (!safe) { ++unsafe_nodes; continue; }
    ++safe_nodes;
    Region& R = regions_[i];
    R.my_id = r;
    if (i == r) {
      R.sz = (int)a.sz;
      R.constraint_id = a.con;
      for (int k = 0; k < 3; ++k) R.descriptor[k] = (float)(a.s[k] / (double)a.sz);
    }
  }
  if (stats) { stats[0] = safe_nodes; stats[1] = unsafe_nodes; stats[2] = safe_comps; stats[3] = unsafe_comps; }
  if (getenv("VSO_SIM_STAGE0_VERBOSE"))
    fprintf(stderr, "stage0: nodes %d safe %lld unsafe %lld comps safe %lld unsafe %lld\n", n, (long long)safe_nodes,
            (long long)unsafe_nodes, (long long)safe_comps, (long long)unsafe_comps);
}

// segmentation_graph.h:339-463
void DenseGraph::SegmentGraph(int min_region_size, bool force_constraints) {
  if (getenv("VSO_SIM_STAGE0")) SimulateStage0(nullptr);
  const float inv_scale = 1.0 / scale_;
  int64_t num_forced_merges = 0, num_regular_merges = 0, num_small_region_merges = 0;
  const float merge_distance_threshold = 0.05f;   // pixel_distance.h:471
  const float split_distance_threshold = 0.15f;   // pixel_distance.h:472
  const int num_lists = (int)bucket_lists_.size();
  const bool trace = getenv("VSO_TRACE_MERGE") != nullptr;
  std::vector<int64_t> trace_stats(64 * 16, 0);
  for (int bucket_idx = 0; bucket_idx < num_buckets_; ++bucket_idx) {
    const float weight = bucket_idx * inv_scale;
    for (int bucket_list_idx = 0; bucket_list_idx < num_lists; ++bucket_list_idx) {
      EdgeList remaining_edges;
      for (const auto& e : bucket_lists_[bucket_list_idx][bucket_idx]) {
        Region* rep_1 = GetRegion(e.region_1);
        Region* rep_2 = GetRegion(e.region_2);
        if (rep_1 == rep_2) continue;
        if (trace) {   // development statistics only (VSO_TRACE_MERGE), no effect on the scan
          const bool b1 = rep_1->sz >= min_region_size, b2 = rep_2->sz >= min_region_size;
          int64_t* t = &trace_stats[(size_t)std::min(bucket_idx, 63) * 16];
          ++t[0];
          t[1 + (b1 ? 1 : 0) + (b2 ? 1 : 0)]++;                       // ss / sh / hh live edges
          if (!rep_1->region_finalized && !rep_2->region_finalized && (rep_1->constraint_id < 0 || rep_2->constraint_id < 0)) {
            const float d = DescriptorDistance(rep_1->descriptor, rep_2->descriptor, 1.0f);
            const float thr = (weight < force_merge_weight_) ? 0.2f : 0.05f;
            ++t[4];
            if (d >= thr) ++t[5 + (b1 ? 1 : 0) + (b2 ? 1 : 0)];      // failed tests ss / sh / hh
            else if (d >= 0.9f * thr) ++t[8 + (b1 ? 1 : 0) + (b2 ? 1 : 0)];   // near passes
          }
          if ((rep_1->region_finalized || rep_2->region_finalized) && (b1 != b2)) ++t[11];   // small into finalised/any big w/o test
          if ((rep_1->region_finalized || rep_2->region_finalized) && b1 && b2) ++t[12];   // inert
          if (rep_1->constraint_id >= 0 && rep_2->constraint_id >= 0) {
            if (rep_1->constraint_id != rep_2->constraint_id) ++t[13];   // passive (different ids)
            else {
              ++t[14];                                                   // same-id tests
              if (DescriptorDistance(rep_1->descriptor, rep_2->descriptor, weight) > 0.15f) ++t[15];   // splits
            }
          }
        }
        if (rep_1->constraint_id < 0 || rep_2->constraint_id < 0) {
          if (!rep_1->region_finalized && !rep_2->region_finalized) {
            const float desc_distance =
                DescriptorDistance(rep_1->descriptor, rep_2->descriptor, weight);
            if (desc_distance < merge_distance_threshold) {
              MergeRegions(rep_1, rep_2);
              ++num_regular_merges;
            } else {
              rep_1->region_finalized = true;
              rep_2->region_finalized = true;
            }
          }
          if (rep_1->region_finalized || rep_2->region_finalized) {
            if (rep_1->sz < min_region_size || rep_2->sz < min_region_size) {
              MergeRegions(rep_1, rep_2);
              ++num_small_region_merges;
            } else {
              remaining_edges.push_back(e);
            }
          }
        } else if (rep_1->constraint_id == rep_2->constraint_id) {
          const float desc_distance =
              DescriptorDistance(rep_1->descriptor, rep_2->descriptor, weight);
          if (desc_distance > split_distance_threshold) {
            if (rep_1->sz < rep_2->sz * 0.3) {
              rep_1->constraint_id = -1;
            } else if (rep_2->sz < rep_1->sz * 0.3) {
              rep_2->constraint_id = -1;
            } else {
              rep_1->constraint_id = -1;
              rep_2->constraint_id = -1;
            }
            remaining_edges.push_back(e);
          } else {
            MergeRegions(rep_1, rep_2);
            ++num_forced_merges;
          }
        } else {
          remaining_edges.push_back(e);
        }
      }
      bucket_lists_[bucket_list_idx][bucket_idx].swap(remaining_edges);
    }
  }
  if (trace) {
    fprintf(stderr, "bucket live ss sh hh | tests fail_ss fail_sh fail_hh near_ss near_sh near_hh | notest_abs inert | passive sameid splits\n");
    for (int b = 0; b < 64; ++b) {
      const int64_t* t = &trace_stats[(size_t)b * 16];
      if (!t[0]) continue;
      fprintf(stderr, "%2d %9lld %9lld %9lld %7lld | %9lld %6lld %6lld %6lld %6lld %6lld %6lld | %8lld %8lld | %8lld %8lld %6lld\n", b, (long long)t[0], (long long)t[1],
              (long long)t[2], (long long)t[3], (long long)t[4], (long long)t[5], (long long)t[6], (long long)t[7], (long long)t[8],
              (long long)t[9], (long long)t[10], (long long)t[11], (long long)t[12], (long long)t[13], (long long)t[14], (long long)t[15]);
    }
  }
  if (force_constraints) MergeConstrainedRegions();
  merge_stats[0] = num_regular_merges;
  merge_stats[1] = num_small_region_merges;
  merge_stats[2] = num_forced_merges;
}

// segmentation_graph.h:703-786
void DenseGraph::MergeConstrainedRegions() {
  std::unordered_map<int, int> constraint_to_region_map;
  std::vector<std::pair<int, int>> virtual_nodes(virtual_nodes_);
  virtual_nodes.push_back(std::make_pair(0, 0));
  virtual_nodes.push_back(std::make_pair((int)regions_.size(), (int)regions_.size()));
  std::sort(virtual_nodes.begin(), virtual_nodes.end());
  const float split_distance_threshold = 0.15f;
  // development switch (not part of the restatement): VSO_SIM_MCR=1 visits every representative once, at its first
  // node, plus one immediate re-visit when it loses its constraint -- the product's round-1 walk, kept to measure
  // what that shortcut costs against the real loop below.
  const bool sim_first_only = getenv("VSO_SIM_MCR") != nullptr;
  std::vector<char> seen_rep(sim_first_only ? regions_.size() : 0, 0);
  for (size_t k = 1; k < virtual_nodes.size(); ++k) {
    for (int idx = virtual_nodes[k - 1].second, end_idx = virtual_nodes[k].first; idx < end_idx; ++idx) {
      if (regions_[idx].constraint_id < 0) continue;
      if (sim_first_only) {
        Region* r = GetRegion(regions_[idx].my_id);
        if (seen_rep[r->my_id] >= 1) { if (!(seen_rep[r->my_id] == 2)) continue; seen_rep[r->my_id] = 1; }
        else seen_rep[r->my_id] = 1;
        if (idx >= 2 * w_ * h_) continue;     // slot 1 only
      }
      Region* my_rep = GetRegion(regions_[idx].my_id);
      auto pos = constraint_to_region_map.find(my_rep->constraint_id);
      if (pos == constraint_to_region_map.end()) {
        constraint_to_region_map.insert(std::make_pair(my_rep->constraint_id, my_rep->my_id));
      } else {
        Region* constraint_rep = GetRegion(pos->second);
        if (constraint_rep != my_rep) {
          const float distance = DescriptorDistance(my_rep->descriptor, constraint_rep->descriptor, 1.0f);
          if (distance > split_distance_threshold) {
            const int before = my_rep->constraint_id;
            if (my_rep->sz < constraint_rep->sz * 0.3) {
              my_rep->constraint_id = -1;
            } else if (constraint_rep->sz < my_rep->sz * 0.3) {
              constraint_rep->constraint_id = -1;
              pos->second = my_rep->my_id;
            } else {
              my_rep->constraint_id = -1;
              constraint_rep->constraint_id = -1;
              constraint_to_region_map.erase(pos);
            }
            if (sim_first_only && before >= 0 && my_rep->constraint_id < 0) seen_rep[my_rep->my_id] = 2;
          } else {
            Region* m = MergeRegions(my_rep, constraint_rep);
            if (sim_first_only) seen_rep[m->my_id] = 1;
          }
        }
      }
    }
  }
  for (size_t k = 0; k < virtual_nodes.size(); ++k) {
    for (int idx = virtual_nodes[k].first, end_idx = virtual_nodes[k].second; idx < end_idx; ++idx) {
      Region* my_rep = GetRegion(regions_[idx].my_id);
      auto pos = constraint_to_region_map.find(my_rep->constraint_id);
      if (pos == constraint_to_region_map.end()) {
        constraint_to_region_map.insert(std::make_pair(my_rep->constraint_id, my_rep->my_id));
      } else {
        Region* constraint_rep = GetRegion(pos->second);
        if (constraint_rep != my_rep) MergeRegions(my_rep, constraint_rep);
      }
    }
  }
}

// segmentation_graph.h:596-629
void DenseGraph::FlattenUnionFind(bool separate_representatives) {
  if (flattened_) return;
  flattened_ = true;
  const int region_offset = (int)regions_.size();
  int new_region_id = region_offset;
  if (separate_representatives) {
    for (int i = 0; i < region_offset; ++i) {
      Region* r = GetRegion(i);
      int flattened_id = r->my_id;
      if (flattened_id < region_offset) {
        r->my_id = new_region_id;
        flattened_id = new_region_id;
        Region nr;
        nr.my_id = new_region_id++;
        nr.sz = r->sz;
        nr.constraint_id = r->constraint_id;
        nr.descriptor[0] = nr.descriptor[1] = nr.descriptor[2] = 0;
        regions_.push_back(nr);
      }
      regions_[i].my_id = flattened_id;
    }
  } else {
    for (auto& region : regions_) region.my_id = GetRegion(region.my_id)->my_id;
  }
}

// segmentation_graph.h:498-522
RegionInformation* DenseGraph::GetCreateRegionInformation(const Region& region, RegionInfoList* list,
                                                          RegionInfoPtrMap* map) {
  auto it = map->find(region.my_id);
  if (it != map->end()) return it->second;
  RegionInformation* ri = new RegionInformation;
  ri->index = max_region_id_++;
  ri->size = region.sz;
  ri->constrained_id = region.constraint_id;
  list->emplace_back(ri);
  map->insert(std::make_pair(region.my_id, ri));
  return ri;
}

// segmentation_graph.h:466-496
void DenseGraph::DetermineNeighborIds(RegionInfoList* list, RegionInfoPtrMap* map) {
  for (int bucket_idx = 0; bucket_idx <= num_buckets_; ++bucket_idx) {
    for (size_t bl = 0; bl < bucket_lists_.size(); ++bl) {
      for (const auto& e : bucket_lists_[bl][bucket_idx]) {
        const Region* r1 = GetRegion(e.region_1);
        const Region* r2 = GetRegion(e.region_2);
        const int r1_id = r1->my_id, r2_id = r2->my_id;
        if (r1_id == r2_id) continue;
        // copy: GetCreateRegionInformation never touches regions_
        RegionInformation* r1_info = GetCreateRegionInformation(*r1, list, map);
        RegionInformation* r2_info = GetCreateRegionInformation(*r2, list, map);
        InsertSortedUniquely(r2_info->index, &r1_info->neighbor_idx);
        InsertSortedUniquely(r1_info->index, &r2_info->neighbor_idx);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// DenseSegmentationGraph
// ---------------------------------------------------------------------------

// dense_segmentation_graph.h:1180-1228 (constraint_ids == nullptr: unconstrained)
void DenseGraph::AddNodesWithDescriptors(const float* img, const int* constraint_ids) {
  const int base_idx = num_frames_ * h_ * w_;
  if ((int)regions_.size() != base_idx) { std::fprintf(stderr, "vso: node count mismatch\n"); std::abort(); }
  for (int i = 0; i < h_; ++i) {
    for (int j = 0; j < w_; ++j) {
      Region r;
      r.my_id = base_idx + i * w_ + j;
      r.sz = 1;
      r.constraint_id = constraint_ids ? constraint_ids[i * w_ + j] : -1;
      const float* p = img + ((size_t)i * w_ + j) * 3;
      r.descriptor[0] = p[0]; r.descriptor[1] = p[1]; r.descriptor[2] = p[2];
      regions_.push_back(r);
    }
  }
}

void DenseGraph::AddNodesAndSpatialEdges(const float* img) {
  AddNodesWithDescriptors(img, nullptr);
  const int frame_idx = num_frames_;
  if (parallel_build_) add_edges_tasks_.emplace_back([this, img, frame_idx] { AddSpatialEdgesImpl(img, frame_idx); });
  else AddSpatialEdgesImpl(img, frame_idx);
  ++num_frames_;
}

void DenseGraph::AddNodesAndSpatialEdgesConstrained(const float* img, const SegDesc& desc) {
  std::vector<int> ids((size_t)w_ * h_, 0);   // region_ids_ view is reused uncleared in the
  // reference; a desc covers every pixel so stale content is always overwritten.
  SegDescToIdImage(desc, w_, ids.data());
  AddNodesWithDescriptors(img, ids.data());
  const int frame_idx = num_frames_;
  if (parallel_build_) add_edges_tasks_.emplace_back([this, img, frame_idx] { AddSpatialEdgesImpl(img, frame_idx); });
  else AddSpatialEdgesImpl(img, frame_idx);
  ++num_frames_;
}

// dense_segmentation_graph.h:327-367
void DenseGraph::AddVirtualNodesConstrained(const SegDesc& desc) {
  const int base_idx = num_frames_ * h_ * w_;
  virtual_slices_.push_back(num_frames_);
  std::vector<int> ids((size_t)w_ * h_, 0);
  SegDescToIdImage(desc, w_, ids.data());
  std::unordered_map<int, int> constraint_to_rep;
  for (int i = 0, region_idx = base_idx; i < h_; ++i) {
    for (int j = 0; j < w_; ++j, ++region_idx) {
      Region r;
      r.my_id = region_idx; r.sz = 0; r.constraint_id = ids[i * w_ + j];
      r.descriptor[0] = r.descriptor[1] = r.descriptor[2] = 0;   // Region(id, sz, constraint): uninitialised in the reference
      regions_.push_back(r);
      auto pos = constraint_to_rep.find(r.constraint_id);
      if (pos == constraint_to_rep.end()) constraint_to_rep.insert(std::make_pair(r.constraint_id, region_idx));
      else regions_[region_idx].my_id = regions_[pos->second].my_id;
    }
  }
  virtual_nodes_.push_back(std::make_pair(base_idx, base_idx + h_ * w_));
  ++num_frames_;
}

// dense_segmentation_graph.h:956-1000
void DenseGraph::AddSpatialEdgesImpl(const float* img, int frame_idx) {
  const int base_idx = frame_idx * h_ * w_;
  const int bucket_list_idx = 2 * frame_idx;
  for (int i = 0, end_y = h_ - 1, cur_idx = base_idx; i <= end_y; ++i) {
    const float* row0 = img + (size_t)i * w_ * 3;
    const float* row1 = row0 + (size_t)w_ * 3;
    for (int j = 0, end_x = w_ - 1; j <= end_x; ++j, ++cur_idx) {
      const float* a = row0 + 3 * j;
      if (j < end_x) AddEdge(cur_idx, cur_idx + 1, PixelDistance(a, row0 + 3 * (j + 1)), bucket_list_idx);
      if (i < end_y) {
        AddEdge(cur_idx, cur_idx + w_, PixelDistance(a, row1 + 3 * j), bucket_list_idx);
        if (j > 0) AddEdge(cur_idx, cur_idx + w_ - 1, PixelDistance(a, row1 + 3 * (j - 1)), bucket_list_idx);
        if (j < end_x) AddEdge(cur_idx, cur_idx + w_ + 1, PixelDistance(a, row1 + 3 * (j + 1)), bucket_list_idx);
      }
    }
  }
}

// dense_segmentation_graph.h:1002-1142: GetLocalEdges + AddTemporalEdgesImpl /
// AddTemporalFlowEdgesImpl; `constant` = ConstantPixelDistance(1e10) (virtual edges).
// frame_idx is num_frames_ AFTER the spatial call incremented it (:933-941).
void DenseGraph::AddTemporalEdgesImpl(const float* curr, const float* prev, const float* flow,
                                      bool constant, int frame_idx) {
  const int base_idx = (frame_idx - 1) * w_ * h_;
  const int base_diff = w_ * h_;
  const int bucket_list_idx = 2 * (frame_idx - 1) - 1;
  int curr_idx = base_idx;
  for (int i = 0; i < h_; ++i) {
    for (int j = 0; j < w_; ++j, ++curr_idx) {
      int x = j, y = i;
      if (flow) {
        const float* flow_ptr = flow + ((size_t)i * w_ + j) * 2;
        int prev_x = j + flow_ptr[0];
        int prev_y = i + flow_ptr[1];
        x = std::max(0, std::min(w_ - 1, prev_x));
        y = std::max(0, std::min(h_ - 1, prev_y));
      }
      const int prev_idx = base_idx - base_diff + y * w_ + x;
      const float* a = constant ? nullptr : curr + ((size_t)i * w_ + j) * 3;
      auto dist = [&](int dx, int dy) -> float {
        if (constant) return 1e10f;
        return PixelDistance(a, prev + ((size_t)(y + dy) * w_ + (x + dx)) * 3);
      };
      if (y > 0) {
        const int lp = prev_idx - w_;
        if (x > 0) AddEdge(curr_idx, lp - 1, dist(-1, -1), bucket_list_idx);
        AddEdge(curr_idx, lp, dist(0, -1), bucket_list_idx);
        if (x + 1 < w_) AddEdge(curr_idx, lp + 1, dist(1, -1), bucket_list_idx);
      }
      if (x > 0) AddEdge(curr_idx, prev_idx - 1, dist(-1, 0), bucket_list_idx);
      AddEdge(curr_idx, prev_idx, dist(0, 0), bucket_list_idx);
      if (x + 1 < w_) AddEdge(curr_idx, prev_idx + 1, dist(1, 0), bucket_list_idx);
      if (y + 1 < h_) {
        const int lp = prev_idx + w_;
        if (x > 0) AddEdge(curr_idx, lp - 1, dist(-1, 1), bucket_list_idx);
        AddEdge(curr_idx, lp, dist(0, 1), bucket_list_idx);
        if (x + 1 < w_) AddEdge(curr_idx, lp + 1, dist(1, 1), bucket_list_idx);
      }
    }
  }
}

void DenseGraph::AddTemporalEdges(const float* curr, const float* prev) {
  const int frame_idx = num_frames_;
  if (parallel_build_) add_edges_tasks_.emplace_back([=] { AddTemporalEdgesImpl(curr, prev, nullptr, false, frame_idx); });
  else AddTemporalEdgesImpl(curr, prev, nullptr, false, frame_idx);
}
void DenseGraph::AddTemporalFlowEdges(const float* curr, const float* prev, const float* flow) {
  const int frame_idx = num_frames_;
  if (parallel_build_) add_edges_tasks_.emplace_back([=] { AddTemporalEdgesImpl(curr, prev, flow, false, frame_idx); });
  else AddTemporalEdgesImpl(curr, prev, flow, false, frame_idx);
}
void DenseGraph::AddTemporalVirtualEdges() {
  const int frame_idx = num_frames_;
  if (parallel_build_) add_edges_tasks_.emplace_back([=] { AddTemporalEdgesImpl(nullptr, nullptr, nullptr, true, frame_idx); });
  else AddTemporalEdgesImpl(nullptr, nullptr, nullptr, true, frame_idx);
}
void DenseGraph::AddTemporalFlowVirtualEdges(const float* flow) {
  const int frame_idx = num_frames_;
  if (parallel_build_) add_edges_tasks_.emplace_back([=] { AddTemporalEdgesImpl(nullptr, nullptr, flow, true, frame_idx); });
  else AddTemporalEdgesImpl(nullptr, nullptr, flow, true, frame_idx);
}

void DenseGraph::FinishBuildingGraph() {
  for (auto& t : add_edges_tasks_) t.join();
  add_edges_tasks_.clear();
}

void DenseGraph::SegmentFullGraph(int min_region_size, bool force_constraints) {
  SegmentGraph(min_region_size, force_constraints);
}

// dense_segmentation_graph.h:432-465
void DenseGraph::AddIntervalToRasterization(int frame, int y, int lx, int rx, int region_id,
                                            RegionInfoList* list, RegionInfoPtrMap* map) {
  RegionInformation* ri = GetCreateRegionInformation(regions_[region_id], list, map);
  if (ri->raster == nullptr) ri->raster.reset(new Rasterization3D);
  if (ri->raster->empty() || ri->raster->back().first < frame) {
    ri->raster->push_back(std::make_pair(frame, std::shared_ptr<Rasterization>(new Rasterization())));
  }
  ri->raster->back().second->push_back(ScanInterval{y, lx, rx});
}

// dense_segmentation_graph.h:1303-1337.  id points at pixel (0,0) of the bordered image.
void DenseGraph::EnforceN4Connectivity(int* id, std::unordered_map<int, int>* size_adjust_map) {
  const int lda = w_ + 2;
  for (int i = 0; i < h_ - 1; ++i) {
    int* region_ptr = id + (size_t)i * lda;
    for (int j = 0; j < w_; ++j, ++region_ptr) {
      const int region_id = *region_ptr;
      if (region_ptr[lda - 1] == region_id && region_ptr[-1] != region_id && region_ptr[lda] != region_id) {
        --(*size_adjust_map)[region_ptr[lda]];
        ++(*size_adjust_map)[region_id];
        region_ptr[lda] = region_id;
      }
      if (region_ptr[lda + 1] == region_id && region_ptr[1] != region_id && region_ptr[lda] != region_id) {
        --(*size_adjust_map)[region_ptr[lda]];
        ++(*size_adjust_map)[region_id];
        region_ptr[lda] = region_id;
      }
    }
  }
}

// dense_segmentation_graph.h:467-579
void DenseGraph::ObtainResults(RegionInfoList* region_list, RegionInfoPtrMap* region_map,
                               const std::vector<const float*>* flows,
                               bool enforce_n4, bool enforce_spatial_connectedness) {
  if (enforce_spatial_connectedness) FlattenUnionFind(true);
  const int lda = w_ + 2;
  std::fill(region_ids_.begin(), region_ids_.end(), -1);      // border = -1 (:483-486); interior overwritten
  int* id_view = region_ids_.data() + lda + 1;
  std::unordered_map<int, int> size_adjust_map;
  const int N = w_ * h_;
  node_labels_after_flatten.assign((size_t)num_frames_ * N, -1);
  id_images_after_n4.assign((size_t)num_frames_ * N, -1);
  for (int idx = 0; idx < num_frames_ * N; ++idx) node_labels_after_flatten[idx] = GetRegion(idx)->my_id;

  for (int t = 0; t < num_frames_; ++t) {
    const int base_idx = N * t;
    if (std::binary_search(virtual_slices_.begin(), virtual_slices_.end(), t)) continue;
    for (int i = 0, idx = base_idx; i < h_; ++i) {
      int* region_ptr = id_view + (size_t)i * lda;
      for (int j = 0; j < w_; ++j, ++idx) region_ptr[j] = GetRegion(idx)->my_id;
    }
    // constrained_slices_ is never filled on the live path (only the dead AddNodesConstrained
    // appends, :1158-1162), so the N4 pass also runs on constrained slices.
    if (enforce_n4) EnforceN4Connectivity(id_view, &size_adjust_map);
    for (int i = 0; i < h_; ++i)
      for (int j = 0; j < w_; ++j) id_images_after_n4[(size_t)base_idx + i * w_ + j] = id_view[(size_t)i * lda + j];

    for (int i = 0; i < h_; ++i) {                             // :533-559
      const int* region_ptr = id_view + (size_t)i * lda;
      int prev_id = region_ptr[0];
      int left_x = 0;
      for (int j = 1; j < w_; ++j) {
        const int curr_id = region_ptr[j];
        if (prev_id != curr_id) {
          AddIntervalToRasterization(t, i, left_x, j - 1, prev_id, region_list, region_map);
          left_x = j;
          prev_id = curr_id;
        }
        if (j + 1 == w_) AddIntervalToRasterization(t, i, left_x, j, prev_id, region_list, region_map);
      }
    }
  }

  if (enforce_spatial_connectedness) EnforceSpatialConnectedness(region_list, region_map, flows, &size_adjust_map);

  for (const auto& entry : size_adjust_map) {                  // :565-578
    const auto pos = region_map->find(entry.first);
    if (pos == region_map->end()) {
      regions_[entry.first].sz = 0;
      continue;
    }
    pos->second->size += entry.second;
  }
}

// ---------------------------------------------------------------------------
// Tube helpers (dense_segmentation_graph.h:581-664, dense_segmentation_graph.cpp:35-209)
// ---------------------------------------------------------------------------
namespace {
struct TubeSlice {
  int frame = -1;
  Rasterization raster;
  ShapeDescriptor shape;
  void ComputeShapeDescriptor() {
    ShapeMoments m;
    ShapeMomentsFromRasterization(raster, &m);
    GetShapeDescriptorFromShapeMoment(m, &shape);
  }
  void MergeFrom(const TubeSlice& other) {
    MergeRasterization(raster, other.raster, &raster);
    ComputeShapeDescriptor();
  }
};
typedef std::vector<TubeSlice> Tube3D;

// dense_segmentation_graph.h:601-628
std::pair<int, float> FindPreviousTube(const TubeSlice& s, const std::vector<Tube3D>& slices, int frame,
                                       const float* flow, int w) {
  Point2f prev_center = s.shape.center;
  if (flow) {
    const float* flow_ptr = flow + ((size_t)(int)prev_center.y * w) * 2 + 2 * (int)prev_center.x;
    prev_center.x += flow_ptr[0];
    prev_center.y += flow_ptr[1];
  }
  float closest_dist = std::numeric_limits<float>::max();
  float closest_idx = -1;
  for (int k = 0; k < (int)slices.size(); ++k) {
    if (slices[k].empty() || slices[k].back().frame >= frame) continue;
    const float dx = slices[k].back().shape.center.x - prev_center.x;
    const float dy = slices[k].back().shape.center.y - prev_center.y;
    const float dist = std::hypot(dy, dx);
    if (dist < closest_dist) { closest_dist = dist; closest_idx = k; }
  }
  return std::make_pair((int)closest_idx, closest_dist);
}

float AverageTubeSliceSize(const Tube3D& ts) {
  if (ts.empty()) return 0;
  float area_sum = 0;
  for (const auto& s : ts) area_sum += s.shape.size;
  return area_sum / ts.size();
}

void MergeTube3D(const Tube3D& lhs, const Tube3D& rhs, Tube3D* result) {
  size_t li = 0, ri = 0;
  if (lhs.empty()) { *result = rhs; return; }
  if (rhs.empty()) { *result = lhs; return; }
  while (li < lhs.size() && ri < rhs.size()) {
    if (lhs[li].frame < rhs[ri].frame) result->push_back(lhs[li++]);
    else if (lhs[li].frame > rhs[ri].frame) result->push_back(rhs[ri++]);
    else {
      TubeSlice merged = lhs[li];
      merged.MergeFrom(rhs[ri]);
      result->push_back(merged);
      ++li; ++ri;
    }
  }
  while (li < lhs.size()) result->push_back(lhs[li++]);
  while (ri < rhs.size()) result->push_back(rhs[ri++]);
}

bool AreTubesTemporalNeighbors(const Tube3D& lhs, const Tube3D& rhs) {
  if (lhs.empty() || rhs.empty()) return false;
  ShapeDescriptor a, b;
  if (lhs[0].frame - 1 == rhs.back().frame) { a = lhs[0].shape; b = rhs.back().shape; }
  else if (lhs.back().frame + 1 == rhs[0].frame) { a = lhs.back().shape; b = rhs[0].shape; }
  else return false;
  const float size_ratio = std::min(a.size, b.size) * (1.0f / std::max(a.size, b.size));
  const float dx = a.center.x - b.center.x, dy = a.center.y - b.center.y;
  return size_ratio > 0.9 && std::hypot(dy, dx) < 20;
}

float AverageTubeDistance(const Tube3D& lhs, const Tube3D& rhs) {
  if (lhs.empty() || rhs.empty()) return std::numeric_limits<float>::max();
  const int start_frame = std::max(lhs[0].frame, rhs[0].frame);
  const int end_frame = std::min(lhs.back().frame, rhs.back().frame);
  int li = 0, ri = 0, weight = 0;
  float diff_sum = 0;
  for (int f = start_frame; f <= end_frame; ++f) {
    while (lhs[li].frame < f) ++li;
    while (rhs[ri].frame < f) ++ri;
    if (lhs[li].frame != f || rhs[ri].frame != f) continue;
    const float dx = lhs[li].shape.center.x - rhs[ri].shape.center.x;
    const float dy = lhs[li].shape.center.y - rhs[ri].shape.center.y;
    diff_sum += std::hypot(dy, dx);
    ++weight;
  }
  return weight > 0 ? diff_sum / weight : std::numeric_limits<float>::max();
}

// segment_util/segmentation_util.cpp:364-379
void ShapeDescriptorBox(const ShapeDescriptor& s, float border, Point2f c[4]) {
  const float ma = s.mag_major * 1.65f + border, mi = s.mag_minor * 1.65f + border;
  const Point2f major{s.dir_major.x * ma, s.dir_major.y * ma};
  const Point2f minor{s.dir_minor.x * mi, s.dir_minor.y * mi};
  const Point2f ctr = s.center;
  c[0] = Point2f{ctr.x - major.x + minor.x, ctr.y - major.y + minor.y};
  c[1] = Point2f{ctr.x - major.x - minor.x, ctr.y - major.y - minor.y};
  c[2] = Point2f{ctr.x + major.x - minor.x, ctr.y + major.y - minor.y};
  c[3] = Point2f{ctr.x + major.x + minor.x, ctr.y + major.y + minor.y};
}

// segment_util/segmentation_util.cpp:381-410
bool ShapeDescriptorBoxesIntersect(const Point2f lhs[4], const Point2f rhs[4]) {
  for (int k = 0; k < 4; ++k) {
    const double lhs_dx = (float)(lhs[(k + 1) % 4].x - lhs[k].x), lhs_dy = (float)(lhs[(k + 1) % 4].y - lhs[k].y);
    for (int l = 0; l < 4; ++l) {
      const double rhs_dx = (float)(rhs[(l + 1) % 4].x - rhs[l].x), rhs_dy = (float)(rhs[(l + 1) % 4].y - rhs[l].y);
      const double delta_x = (float)(rhs[l].x - lhs[k].x), delta_y = (float)(rhs[l].y - lhs[k].y);
      const double kross = lhs_dx * rhs_dy - lhs_dy * rhs_dx;
      if (std::fabs(kross) < 1e-6) continue;
      const float inv_kross = 1.0f / kross;
      const double t = (delta_x * rhs_dy - delta_y * rhs_dx) * inv_kross;
      const double s = (delta_x * lhs_dy - delta_y * lhs_dx) * inv_kross;
      if (t > -1e-6f && t < 1.0f + 1e-6f && s > -1e-6f && s < 1.0f + 1e-6f) return true;
    }
  }
  return false;
}

float Tube3DIntersection(const Tube3D& lhs, const Tube3D& rhs) {
  if (lhs.empty() || rhs.empty()) return std::numeric_limits<float>::max();
  const int start_frame = std::max(lhs[0].frame, rhs[0].frame);
  const int end_frame = std::min(lhs.back().frame, rhs.back().frame);
  int li = 0, ri = 0, intersect_count = 0, weight = 0;
  for (int f = start_frame; f <= end_frame; ++f) {
    while (lhs[li].frame < f) ++li;
    while (rhs[ri].frame < f) ++ri;
    if (lhs[li].frame != f || rhs[ri].frame != f) continue;
    Point2f lb[4], rb[4];
    ShapeDescriptorBox(lhs[li].shape, 10, lb);
    ShapeDescriptorBox(rhs[ri].shape, 10, rb);
    if (ShapeDescriptorBoxesIntersect(lb, rb)) ++intersect_count;
    ++weight;
  }
  return weight > 0 ? intersect_count * (1.0f / weight) : std::numeric_limits<float>::max();
}

int GetClosestTube3D(const Tube3D& tube, const std::vector<Tube3D>& tubes, int ignore_index) {
  float min_dist = std::numeric_limits<float>::max();
  int min_idx = -1;
  for (int k = 0; k < (int)tubes.size(); ++k) {
    if (k == ignore_index) continue;
    const float d = AverageTubeDistance(tube, tubes[k]);
    if (d < min_dist) { min_dist = d; min_idx = k; }
  }
  return min_idx;
}
}  // namespace

// dense_segmentation_graph.h:666-904
void DenseGraph::EnforceSpatialConnectedness(RegionInfoList* region_list, RegionInfoPtrMap* region_map,
                                             const std::vector<const float*>* flows,
                                             std::unordered_map<int, int>* size_adjust_map) {
  const int num_regions = (int)region_list->size();
  for (int r = 0; r < num_regions; ++r) {
    RegionInformation& ri = *(*region_list)[r];
    if (ri.raster == nullptr) continue;
    Rasterization3D& raster = *ri.raster;
    std::vector<Tube3D> result_tubes;
    std::vector<Tube3D> active_tubes;
    const float inv_frame_diam = 1.0f / std::hypot((float)w_, (float)h_);

    for (const auto& raster_slice : raster) {
      const int frame = raster_slice.first;
      std::vector<Rasterization> components;
      ConnectedComponentsN4(*raster_slice.second, &components);
      std::vector<TubeSlice> slices;
      slices.reserve(components.size());
      for (auto& comp : components) {
        TubeSlice slice;
        slice.frame = frame;
        slice.raster.swap(comp);
        slice.ComputeShapeDescriptor();
        slices.push_back(std::move(slice));
      }
      components.clear();

      if (active_tubes.empty()) {
        for (auto& slice : slices) active_tubes.push_back(Tube3D{std::move(slice)});
      } else {
        std::vector<Tube3D> new_active_tubes;
        std::vector<int> used_indices(active_tubes.size(), 0);
        for (auto& slice : slices) {
          const auto match = FindPreviousTube(slice, active_tubes, frame,
                                              flows ? (*flows)[frame] : nullptr, w_);
          const int prev_idx = match.first;
          if (prev_idx < 0) {
            new_active_tubes.push_back(Tube3D{std::move(slice)});
            continue;
          }
          const float diff_dist = match.second;
          const float area_ratio =
              std::min(active_tubes[prev_idx].back().shape.size, slice.shape.size) /
              (std::max(active_tubes[prev_idx].back().shape.size, slice.shape.size) + 1e-6);
          if (area_ratio > 0.75 && diff_dist * inv_frame_diam < 0.04f) {
            ++used_indices[prev_idx];
            active_tubes[prev_idx].push_back(std::move(slice));
            new_active_tubes.push_back(Tube3D());
            new_active_tubes.back().swap(active_tubes[prev_idx]);
          } else {
            new_active_tubes.push_back(Tube3D{std::move(slice)});
          }
        }
        for (size_t k = 0; k < active_tubes.size(); ++k) {
          if (used_indices[k] == 0) result_tubes.push_back(std::move(active_tubes[k]));
        }
        new_active_tubes.swap(active_tubes);
      }
    }
    for (auto& t : active_tubes) result_tubes.push_back(std::move(t));
    if (result_tubes.size() <= 1) continue;

    auto merge_with_closest_tube = [&result_tubes](int k) -> bool {
      const int idx = GetClosestTube3D(result_tubes[k], result_tubes, k);
      if (idx < 0) return false;
      Tube3D merged;
      MergeTube3D(result_tubes[idx], result_tubes[k], &merged);
      result_tubes[idx].swap(merged);
      result_tubes.erase(result_tubes.begin() + k);
      return true;
    };

    for (int k = 0; k < (int)result_tubes.size();) {
      bool merge = AverageTubeSliceSize(result_tubes[k]) < 20;
      if (!merge) {
        for (int l = 0; l < (int)result_tubes.size(); ++l) {
          if (l == k) continue;
          if (Tube3DIntersection(result_tubes[k], result_tubes[l]) > 0.8) { merge = true; break; }
        }
      }
      if (merge && merge_with_closest_tube(k)) {
      } else {
        ++k;
      }
    }

    for (int k = 0; k < (int)result_tubes.size();) {
      bool is_merged = false;
      for (int l = 0; l < (int)result_tubes.size(); ++l) {
        if (l == k) continue;
        if (AreTubesTemporalNeighbors(result_tubes[k], result_tubes[l])) {
          Tube3D merged;
          MergeTube3D(result_tubes[k], result_tubes[l], &merged);
          result_tubes[l].swap(merged);
          result_tubes.erase(result_tubes.begin() + k);
          is_merged = true;
          break;
        }
      }
      if (!is_merged) ++k;
    }

    int tube_to_keep = -1;
    int tube_to_keep_score = 0;
    std::vector<float> tube_areas(result_tubes.size());
    for (int k = 0; k < (int)result_tubes.size(); ++k) {
      float area = 0;
      for (const auto& slice : result_tubes[k]) area += slice.shape.size;
      tube_areas[k] = area;
      const float tube_score = area;
      if (tube_score > tube_to_keep_score) { tube_to_keep_score = tube_score; tube_to_keep = k; }
    }

    for (int k = 0; k < (int)result_tubes.size(); ++k) {
      int first_idx = result_tubes[k][0].frame * w_ * h_;
      const auto& first_scanline = result_tubes[k][0].raster[0];
      first_idx += first_scanline.y * w_ + first_scanline.left_x;
      Region* rep = GetRegion(first_idx);
      if (k != tube_to_keep) {
        (*size_adjust_map)[rep->my_id] -= tube_areas[k];
        Region nr;
        nr.my_id = (int)regions_.size();
        nr.sz = tube_areas[k];
        nr.constraint_id = -1;
        nr.descriptor[0] = nr.descriptor[1] = nr.descriptor[2] = 0;
        regions_.push_back(nr);
        rep = &regions_.back();
        const int region_id = rep->my_id;
        for (const auto& slice : result_tubes[k]) {
          const int base_idx = slice.frame * w_ * h_;
          for (const auto& s : slice.raster) {
            const int row_idx = base_idx + s.y * w_;
            for (int x = s.left_x; x <= s.right_x; ++x) regions_[row_idx + x].my_id = region_id;
          }
        }
      }
      RegionInformation* info = GetCreateRegionInformation(*rep, region_list, region_map);
      info->raster.reset(new Rasterization3D);
      for (auto& slice : result_tubes[k]) {
        std::shared_ptr<Rasterization> nr(new Rasterization());
        nr->swap(slice.raster);
        info->raster->push_back(std::make_pair(slice.frame, nr));
      }
    }
  }
}

// Test hook (tests/host_tubes_check.cpp): the state SegmentFullGraph would leave behind for a given label volume -- every
// voxel's parent is the first voxel of its label -- so that ObtainResults / EnforceSpatialConnectedness can be run on
// label volumes made up by a test.
void DenseGraph::TestLoadLabels(const int32_t* labels, int frames) {
  std::vector<float> blank((size_t)w_ * h_ * 3, 0.f);
  std::unordered_map<int, int> first_of_label;
  for (int t = 0; t < frames; ++t) {
    AddNodesWithDescriptors(blank.data(), nullptr);
    ++num_frames_;
  }
  const int n = w_ * h_ * frames;
  for (int i = 0; i < n; ++i) {
    auto it = first_of_label.find(labels[i]);
    if (it == first_of_label.end()) { first_of_label[labels[i]] = i; continue; }
    regions_[i].my_id = it->second;
    regions_[it->second].sz += 1;
  }
}

}  // namespace vso

// region index (position in the region list after EnforceSpatialConnectedness) of every voxel of a label volume;
// flows: [frames][h][w][2] backward flow or NULL.  Returns the number of regions.
extern "C" int vso_test_spatial_connectedness(const int32_t* labels, int w, int h, int frames, const float* flows,
                                              int32_t* region_index_out) {
  vso::DenseGraph g(w, h, frames, false, false);
  g.TestLoadLabels(labels, frames);
  vso::RegionInfoList list;
  vso::RegionInfoPtrMap map;
  std::vector<const float*> fl;
  if (flows) for (int t = 0; t < frames; ++t) fl.push_back(t == 0 ? nullptr : flows + (size_t)t * w * h * 2);
  g.ObtainResults(&list, &map, flows ? &fl : nullptr, false, true);
  std::fill(region_index_out, region_index_out + (size_t)w * h * frames, -1);
  for (const auto& ri : list) {
    if (!ri->raster) continue;
    for (const auto& slice : *ri->raster)
      for (const auto& s : *slice.second)
        for (int x = s.left_x; x <= s.right_x; ++x) region_index_out[((size_t)slice.first * h + s.y) * w + x] = ri->index;
  }
  return (int)list.size();
}
