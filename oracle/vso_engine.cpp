// vso_engine.cpp -- CPU ORACLE (test infrastructure): the streaming driver and C ABI.
// Restates Segmentation (segmentation/segmentation.cpp:58-78,272-303,392-582,671-773)
// and DenseSegmentation (segmentation/dense_segmentation.cpp:50-162,268-432).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <unordered_map>
#include <vector>

#include "vso.h"
#include "vso_core.hpp"

namespace vso {

static double NowSec() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// segmentation/segmentation.h:46-95 (over-segmentation subset)
struct SegmentationOptions {
  int min_region_size = 200;
  bool enforce_n4_connectivity = true;
  bool enforce_spatial_connectedness = true;
};

// ---------------------------------------------------------------------------
// Segmentation: per-chunk engine (over-segmentation + result halves)
// ---------------------------------------------------------------------------
class Segmentation {
 public:
  Segmentation(const SegmentationOptions& o, int w, int h, int chunk_id)
      : options_(o), w_(w), h_(h), chunk_id_(chunk_id) {}

  // segmentation.cpp:58-64
  void InitializeOverSegmentation(int chunk_size, bool l1, bool parallel) {
    graph_.reset(new DenseGraph(w_, h_, chunk_size, l1, parallel));
  }
  DenseGraph* graph() { return graph_.get(); }

  // segmentation.cpp:272-303
  void RunOverSegmentation(const std::vector<const float*>* flows, double* t_segment, double* t_results,
                           std::vector<int32_t>* node_labels, std::vector<int32_t>* id_images,
                           int* slots, int64_t stats[3]) {
    std::unique_ptr<RegionInfoList> region_list(new RegionInfoList());
    graph_->FinishBuildingGraph();
    double t0 = NowSec();
    graph_->SegmentFullGraph(options_.min_region_size, true);
    double t1 = NowSec();
    RegionInfoPtrMap map;
    graph_->ObtainResults(region_list.get(), &map, flows, options_.enforce_n4_connectivity,
                          options_.enforce_spatial_connectedness);
    graph_->DetermineNeighborIds(region_list.get(), &map);
    double t2 = NowSec();
    *t_segment += t1 - t0;
    *t_results += t2 - t1;
    node_labels->swap(graph_->node_labels_after_flatten);
    id_images->swap(graph_->id_images_after_n4);
    *slots = graph_->num_frames();
    std::memcpy(stats, graph_->merge_stats, sizeof(int64_t) * 3);
    graph_.reset();
    region_infos_ = std::move(region_list);
  }

  // segmentation.cpp:392-403 (level 0 only)
  void ConstrainSegmentationToFrameInterval(int lhs, int rhs) {
    for (auto& r : *region_infos_) {
      if (r->raster == nullptr || r->raster->empty() || r->raster->front().first >= rhs ||
          r->raster->back().first < lhs) {
        r->flagged_for_removal = true;
      }
    }
  }

  // segmentation.cpp:424-441 (level 0 only)
  void AdjustRegionAreaToFrameInterval(int lhs, int rhs) {
    for (auto& r : *region_infos_) {
      int size_increment = 0;
      if (r->raster == nullptr) continue;
      for (const auto& slice : *r->raster) {
        if (slice.first < lhs || slice.first >= rhs) size_increment -= RasterizationArea(*slice.second);
      }
      r->size += size_increment;
    }
  }

  // segmentation.cpp:537-582 (one level)
  void AssignUniqueRegionIds(bool use_constrained_ids, int region_id_offset, int* max_region_id) {
    assigned_constrained_ids_ = use_constrained_ids;
    int max_id = -1;
    for (auto& r : *region_infos_) {
      r->region_id = (use_constrained_ids && r->constrained_id >= 0) ? r->constrained_id
                                                                      : r->index + region_id_offset;
      max_id = std::max(max_id, r->region_id);
    }
    if (max_region_id) *max_region_id = std::max(region_id_offset, max_id + 1);
  }

  // segmentation.cpp:458-533 + 671-773 (level-0 hierarchy, no vectorisation)
  void RetrieveSegmentation3D(int frame_number, bool output_hierarchy, SegDesc* desc) {
    desc->frame_width = w_;
    desc->frame_height = h_;
    desc->chunk_id = chunk_id_;
    desc->connectedness = options_.enforce_n4_connectivity ? 1 : 2;
    for (const auto& rp : *region_infos_) {
      const RegionInformation& ri = *rp;
      if (ri.raster == nullptr) continue;
      auto it = std::lower_bound(ri.raster->begin(), ri.raster->end(), frame_number,
                                 [](const std::pair<int, std::shared_ptr<Rasterization>>& a, int f) {
                                   return a.first < f;
                                 });
      if (it == ri.raster->end() || it->first != frame_number) continue;
      Region2D r;
      r.id = ri.region_id;
      r.raster = *it->second;
      ShapeMomentsFromRasterization(r.raster, &r.moments);
      desc->region.push_back(std::move(r));
    }
    if (assigned_constrained_ids_) {
      std::sort(desc->region.begin(), desc->region.end(),
                [](const Region2D& a, const Region2D& b) { return a.id < b.id; });
    }
    if (output_hierarchy) {
      desc->has_hierarchy = true;
      for (const auto& rp : *region_infos_) {
        const RegionInformation& ri = *rp;
        if (ri.flagged_for_removal) continue;
        CompoundRegion c;
        c.id = ri.region_id;
        c.size = ri.size;
        for (int n : ri.neighbor_idx) {
          if ((*region_infos_)[n]->flagged_for_removal) continue;
          c.neighbor_id.push_back((*region_infos_)[n]->region_id);
        }
        if (assigned_constrained_ids_) std::sort(c.neighbor_id.begin(), c.neighbor_id.end());
        c.start_frame = ri.raster->front().first;
        c.end_frame = ri.raster->back().first;
        desc->hierarchy0.push_back(std::move(c));
      }
      if (assigned_constrained_ids_) {
        std::sort(desc->hierarchy0.begin(), desc->hierarchy0.end(),
                  [](const CompoundRegion& a, const CompoundRegion& b) { return a.id < b.id; });
      }
    }
  }

 private:
  SegmentationOptions options_;
  int w_, h_, chunk_id_;
  std::unique_ptr<DenseGraph> graph_;
  std::unique_ptr<RegionInfoList> region_infos_;
  bool assigned_constrained_ids_ = false;
};

// ---------------------------------------------------------------------------
// DenseSegmentation: the streaming chunker
// ---------------------------------------------------------------------------
class DenseSegmentation {
 public:
  DenseSegmentation(const vso_dense_opts& o, int w, int h, bool use_flow)
      : options_(o), w_(w), h_(h), use_flow_(use_flow) {
    // dense_segmentation.cpp:55-75
    overlap_frames_ = options_.chunk_overlap_ratio * options_.chunk_size + 0.5f;
    overlap_frames_ = std::min(overlap_frames_, 2);
    constraint_frames_ = std::min(options_.num_constraint_frames, overlap_frames_ - 1);
  }

  // dense_segmentation.cpp:268-279
  void GetSegmentationOptions(SegmentationOptions* so) const {
    so->min_region_size = options_.frac_min_region_size * w_ * options_.frac_min_region_size * h_ *
                          options_.chunk_size;
    so->enforce_n4_connectivity = options_.enforce_n4_connectivity != 0;
    so->enforce_spatial_connectedness = options_.enforce_spatial_connectedness != 0;
  }

  void NewSegmentation(int chunk_size) {
    SegmentationOptions so;
    GetSegmentationOptions(&so);
    seg_.reset(new Segmentation(so, w_, h_, chunk_id_));
    seg_->InitializeOverSegmentation(chunk_size, options_.color_distance == 0, options_.num_threads > 1);
  }

  // dense_segmentation.cpp:108-162
  int ProcessFrame(bool flush, const uint8_t* bgr, int row_stride, const float* flow, int flow_stride,
                   std::vector<std::unique_ptr<SegDesc>>* results) {
    if (seg_ == nullptr) NewSegmentation(options_.chunk_size);
    if (bgr) {
      double t0 = NowSec();
      std::shared_ptr<std::vector<float>> feat(new std::vector<float>((size_t)w_ * h_ * 3));
      Preprocess(bgr, row_stride, feat->data());
      stage_sec[0] += NowSec() - t0;
      feature_buffer_.push_back(feat);
      if (use_flow_) {
        if (input_frames_ == 0) {
          flow_buffer_.push_back(nullptr);
        } else {
          std::shared_ptr<std::vector<float>> fl(new std::vector<float>((size_t)w_ * h_ * 2));
          for (int i = 0; i < h_; ++i)
            std::memcpy(fl->data() + (size_t)i * w_ * 2, (const char*)flow + (size_t)i * flow_stride,
                        sizeof(float) * 2 * w_);
          flow_buffer_.push_back(fl);
        }
      }
      t0 = NowSec();
      seg_->graph()->AddNodesAndSpatialEdges(feature_buffer_.back()->data());
      if (feature_buffer_.size() > 1) {
        const float* cur = feature_buffer_.end()[-1]->data();
        const float* prev = feature_buffer_.end()[-2]->data();
        if (use_flow_) seg_->graph()->AddTemporalFlowEdges(cur, prev, flow_buffer_.back()->data());
        else seg_->graph()->AddTemporalEdges(cur, prev);
      }
      stage_sec[1] += NowSec() - t0;
      ++input_frames_;
    }
    if (flush || (int)feature_buffer_.size() - curr_chunk_start_ >= options_.chunk_size) {
      if (feature_buffer_.empty()) { seg_.reset(); return 0; }
      ChunkBoundaryOutput(flush, results);
      return (int)results->size();
    }
    return 0;
  }

  // dense_segmentation.cpp:164-198
  void Preprocess(const uint8_t* bgr, int row_stride, float* out) {
    if (options_.presmoothing == 2) {
      std::vector<float> tmp((size_t)w_ * h_ * 3);
      ConvertU8ToF32(bgr, w_, h_, row_stride, tmp.data());
      BilateralFilter(tmp.data(), w_, h_, 3.0, 0.25, out, options_.num_threads, nullptr, nullptr);
    } else if (options_.presmoothing == 0) {
      ConvertU8ToF32(bgr, w_, h_, row_stride, out);
    } else {
      std::fprintf(stderr, "vso: gaussian presmoothing (cv::GaussianBlur) is not restated\n");
      std::abort();
    }
  }

  // dense_segmentation.cpp:281-328
  void ChunkBoundaryOutput(bool flush, std::vector<std::unique_ptr<SegDesc>>* results) {
    SegmentAndOutputChunk(flush, results);
    if (flush) { seg_.reset(); return; }
    double t0 = NowSec();
    NewSegmentation(curr_chunk_start_ + options_.chunk_size);
    seg_->graph()->AddVirtualNodesConstrained(*overlap_segmentations_[0]);
    seg_->graph()->AddNodesAndSpatialEdgesConstrained(feature_buffer_[1]->data(), *overlap_segmentations_[1]);
    if (use_flow_) seg_->graph()->AddTemporalFlowVirtualEdges(flow_buffer_[1]->data());
    else seg_->graph()->AddTemporalVirtualEdges();
    for (int i = 2; i < overlap_frames_; ++i) {   // never runs with overlap_frames_ <= 2 (:317-326)
      if (i < constraint_frames_) seg_->graph()->AddNodesAndSpatialEdgesConstrained(feature_buffer_[i]->data(), *overlap_segmentations_[i]);
      else seg_->graph()->AddNodesAndSpatialEdges(feature_buffer_[i]->data());
      if (use_flow_) seg_->graph()->AddTemporalFlowEdges(feature_buffer_[i]->data(), feature_buffer_[i - 1]->data(), flow_buffer_[i]->data());
      else seg_->graph()->AddTemporalEdges(feature_buffer_[i]->data(), feature_buffer_[i - 1]->data());
    }
    // test tap: the hand-over state of this boundary (what vsb200_dense_export_halo gives for the product): the two
    // overlap frames' region-id maps, max_region_id_, the id of the chunk they constrain, frames output so far
    last_overlap_maps.assign((size_t)2 * w_ * h_, 0);
    SegDescToIdImage(*overlap_segmentations_[0], w_, last_overlap_maps.data());
    SegDescToIdImage(*overlap_segmentations_[1], w_, last_overlap_maps.data() + (size_t)w_ * h_);
    last_chain_state[0] = max_region_id_; last_chain_state[1] = chunk_id_; last_chain_state[2] = num_output_frames_;
    overlap_segmentations_.clear();
    stage_sec[1] += NowSec() - t0;
  }
  std::vector<int> last_overlap_maps;
  int last_chain_state[3] = {0, 0, 0};

  // dense_segmentation.cpp:330-432
  void SegmentAndOutputChunk(bool flush, std::vector<std::unique_ptr<SegDesc>>* results) {
    std::vector<const float*> flows;
    if (use_flow_) for (auto& f : flow_buffer_) flows.push_back(f ? f->data() : nullptr);
    seg_->RunOverSegmentation(use_flow_ ? &flows : nullptr, &stage_sec[2], &stage_sec[3],
                              &last_node_labels, &last_id_images, &last_slots, last_stats);
    double t0 = NowSec();
    const int buf = (int)feature_buffer_.size();
    const int overlap_start = buf - (flush ? 0 : overlap_frames_);
    const int last_output_frame = std::min<int>(buf - 1, overlap_start);
    const int max_result_frame = std::min<int>(buf - 1, last_output_frame + constraint_frames_);
    seg_->ConstrainSegmentationToFrameInterval(0, last_output_frame + 1);
    seg_->AdjustRegionAreaToFrameInterval(0, last_output_frame + 1);
    int new_max_region_id = 0;
    const bool use_constraints = chunk_id_ > 0;
    seg_->AssignUniqueRegionIds(use_constraints, max_region_id_, &new_max_region_id);
    max_region_id_ = new_max_region_id;
    const int chunk_size = last_output_frame - curr_chunk_start_ + 1;
    results->clear();
    overlap_segmentations_.clear();
    const int hierarchy_frame_idx = num_output_frames_;
    for (int frame_idx = curr_chunk_start_; frame_idx <= max_result_frame; ++frame_idx) {
      std::unique_ptr<SegDesc> desc(new SegDesc());
      const bool output_hierarchy = frame_idx == curr_chunk_start_;
      seg_->RetrieveSegmentation3D(frame_idx, output_hierarchy, desc.get());
      desc->chunk_size = chunk_size;
      desc->overlap_start = chunk_size;
      desc->hierarchy_frame_idx = hierarchy_frame_idx;
      if (frame_idx <= last_output_frame) {
        if (frame_idx < last_output_frame) {
          results->push_back(std::move(desc));
        } else {
          results->push_back(std::unique_ptr<SegDesc>(new SegDesc(*desc)));
        }
        ++num_output_frames_;
      }
      if (frame_idx >= last_output_frame) {
        if (desc) overlap_segmentations_.push_back(std::move(desc));
        else overlap_segmentations_.push_back(std::unique_ptr<SegDesc>(new SegDesc(*results->back())));
      }
    }
    feature_buffer_.erase(feature_buffer_.begin(), feature_buffer_.begin() + last_output_frame);
    if (use_flow_) flow_buffer_.erase(flow_buffer_.begin(), flow_buffer_.begin() + last_output_frame);
    curr_chunk_start_ = flush ? 0 : 1;
    if (!flush) {
      feature_buffer_[0].reset();
      if (use_flow_) flow_buffer_[0].reset();
    }
    ++chunk_id_;
    stage_sec[4] += NowSec() - t0;
  }

  double stage_sec[5] = {0, 0, 0, 0, 0};
  std::vector<int32_t> last_node_labels, last_id_images;
  int last_slots = 0;
  int64_t last_stats[3] = {0, 0, 0};
  int w() const { return w_; }
  int h() const { return h_; }

 private:
  vso_dense_opts options_;
  int w_, h_;
  bool use_flow_;
  int input_frames_ = 0, chunk_id_ = 0, overlap_frames_ = 2, constraint_frames_ = 1;
  int max_region_id_ = 0, num_output_frames_ = 0, curr_chunk_start_ = 0;
  std::vector<std::shared_ptr<std::vector<float>>> feature_buffer_, flow_buffer_;
  std::vector<std::unique_ptr<SegDesc>> overlap_segmentations_;
  std::unique_ptr<Segmentation> seg_;
};

}  // namespace vso

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
struct vso_dense {
  std::unique_ptr<vso::DenseSegmentation> seg;
  std::deque<std::pair<std::unique_ptr<vso::SegDesc>, int64_t>> ready;
  std::deque<int64_t> pts_queue;
  // flattened storage of the most recently popped result
  std::vector<int32_t> region_id, interval_offset, intervals, compound, neighbor_offset, neighbor_id;
  std::vector<float> moments;
};

extern "C" {

void vso_default_opts(vso_dense_opts* o) {
  o->presmoothing = 2;
  o->frac_min_region_size = 0.01;
  o->chunk_size = 20;
  o->chunk_overlap_ratio = 0.2;
  o->num_constraint_frames = 1;
  o->enforce_n4_connectivity = 1;
  o->enforce_spatial_connectedness = 1;
  o->color_distance = 1;
  o->num_threads = 1;
}

void vso_convert_u8_to_f32(const uint8_t* bgr, int w, int h, int row_stride, float* out) {
  vso::ConvertU8ToF32(bgr, w, h, row_stride, out);
}

void vso_bilateral(const float* in, int w, int h, float sigma_space, float sigma_color, float* out,
                   int num_threads, float* lut_out, float* scale_out) {
  vso::BilateralFilter(in, w, h, sigma_space, sigma_color, out, num_threads, lut_out, scale_out);
}

void vso_preprocess(const uint8_t* bgr, int w, int h, int row_stride, int presmoothing, float* out,
                    int num_threads) {
  if (presmoothing == 2) {
    std::vector<float> tmp((size_t)w * h * 3);
    vso::ConvertU8ToF32(bgr, w, h, row_stride, tmp.data());
    vso::BilateralFilter(tmp.data(), w, h, 3.0, 0.25, out, num_threads, nullptr, nullptr);
  } else {
    vso::ConvertU8ToF32(bgr, w, h, row_stride, out);
  }
}

// dense_segmentation_graph.h:956-1000 in planar layout.
void vso_spatial_weights(const float* img, int w, int h, int l1, float* out) {
  const size_t N = (size_t)w * h;
  auto dist = [&](const float* a, const float* b) { return l1 ? vso::ColorDiff3L1(a, b) : vso::ColorDiff3L2(a, b); };
  for (size_t k = 0; k < 4 * N; ++k) out[k] = -1.0f;
  for (int i = 0; i < h; ++i) {
    for (int j = 0; j < w; ++j) {
      const float* a = img + ((size_t)i * w + j) * 3;
      const size_t p = (size_t)i * w + j;
      if (j < w - 1) out[0 * N + p] = dist(a, a + 3);
      if (i < h - 1) {
        out[1 * N + p] = dist(a, a + (size_t)w * 3);
        if (j > 0) out[2 * N + p] = dist(a, a + (size_t)w * 3 - 3);
        if (j < w - 1) out[3 * N + p] = dist(a, a + (size_t)w * 3 + 3);
      }
    }
  }
}

// dense_segmentation_graph.h:1002-1142 in planar layout.
void vso_temporal_weights(const float* curr, const float* prev, const float* flow, int w, int h, int l1,
                          float* out) {
  const size_t N = (size_t)w * h;
  auto dist = [&](const float* a, const float* b) { return l1 ? vso::ColorDiff3L1(a, b) : vso::ColorDiff3L2(a, b); };
  for (size_t k = 0; k < 9 * N; ++k) out[k] = -1.0f;
  for (int i = 0; i < h; ++i) {
    for (int j = 0; j < w; ++j) {
      int x = j, y = i;
      if (flow) {
        const float* fp = flow + ((size_t)i * w + j) * 2;
        int px = j + fp[0];
        int py = i + fp[1];
        x = std::max(0, std::min(w - 1, px));
        y = std::max(0, std::min(h - 1, py));
      }
      const float* a = curr + ((size_t)i * w + j) * 3;
      const size_t p = (size_t)i * w + j;
      int d = 0;
      for (int dy = -1; dy <= 1; ++dy) {
        for (int dx = -1; dx <= 1; ++dx, ++d) {
          const int xx = x + dx, yy = y + dy;
          if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
          out[d * N + p] = dist(a, prev + ((size_t)yy * w + xx) * 3);
        }
      }
    }
  }
}

int vso_bucket_index(float weight) {
  const float max_weight = 1.0;
  const float scale = 2048 / (max_weight + 1e-6f);
  return (int)(std::min<float>(2048, weight * scale));
}

int vso_dense_create(const vso_dense_opts* o, int w, int h, int use_flow, vso_dense** out) {
  if (!o || !out || w <= 1 || h <= 1 || o->chunk_size < 3) return 1;
  const int overlap = std::min((int)(o->chunk_overlap_ratio * o->chunk_size + 0.5f), 2);
  if (overlap >= o->chunk_size || o->num_constraint_frames < 1 || overlap < 2) return 1;
  vso_dense* d = new vso_dense;
  d->seg.reset(new vso::DenseSegmentation(*o, w, h, use_flow != 0));
  *out = d;
  return 0;
}

static void Enqueue(vso_dense* d, std::vector<std::unique_ptr<vso::SegDesc>>* results, int* n_ready) {
  for (auto& r : *results) {
    const int64_t pts = d->pts_queue.front();
    d->pts_queue.pop_front();
    d->ready.emplace_back(std::move(r), pts);
  }
  if (n_ready) *n_ready = (int)results->size();
}

int vso_dense_push(vso_dense* d, const uint8_t* bgr, int row_stride, const float* flow_xy,
                   int flow_row_stride_bytes, int64_t pts, int* n_ready) {
  d->pts_queue.push_back(pts);
  std::vector<std::unique_ptr<vso::SegDesc>> results;
  d->seg->ProcessFrame(false, bgr, row_stride, flow_xy, flow_row_stride_bytes, &results);
  Enqueue(d, &results, n_ready);
  return 0;
}

int vso_dense_flush(vso_dense* d, int* n_ready) {
  std::vector<std::unique_ptr<vso::SegDesc>> results;
  d->seg->ProcessFrame(true, nullptr, 0, nullptr, 0, &results);
  Enqueue(d, &results, n_ready);
  return 0;
}

int vso_dense_pop(vso_dense* d, vso_frame_result* out) {
  if (d->ready.empty()) return 1;
  std::unique_ptr<vso::SegDesc> s = std::move(d->ready.front().first);
  const int64_t pts = d->ready.front().second;
  d->ready.pop_front();
  d->region_id.clear(); d->interval_offset.clear(); d->intervals.clear(); d->moments.clear();
  d->compound.clear(); d->neighbor_offset.clear(); d->neighbor_id.clear();
  d->interval_offset.push_back(0);
  for (const auto& r : s->region) {
    d->region_id.push_back(r.id);
    for (const auto& si : r.raster) {
      d->intervals.push_back(si.y); d->intervals.push_back(si.left_x); d->intervals.push_back(si.right_x);
    }
    d->interval_offset.push_back((int32_t)(d->intervals.size() / 3));
    const float m[6] = {r.moments.size, r.moments.mean_x, r.moments.mean_y, r.moments.moment_xx,
                        r.moments.moment_xy, r.moments.moment_yy};
    d->moments.insert(d->moments.end(), m, m + 6);
  }
  d->neighbor_offset.push_back(0);
  for (const auto& c : s->hierarchy0) {
    d->compound.push_back(c.id); d->compound.push_back(c.size);
    d->compound.push_back(c.start_frame); d->compound.push_back(c.end_frame);
    d->neighbor_id.insert(d->neighbor_id.end(), c.neighbor_id.begin(), c.neighbor_id.end());
    d->neighbor_offset.push_back((int32_t)d->neighbor_id.size());
  }
  out->width = s->frame_width; out->height = s->frame_height; out->chunk_id = s->chunk_id;
  out->chunk_size = s->chunk_size; out->overlap_start = s->overlap_start;
  out->hierarchy_frame_idx = s->hierarchy_frame_idx; out->connectedness = s->connectedness;
  out->n_regions = (int32_t)s->region.size();
  out->region_id = d->region_id.data(); out->interval_offset = d->interval_offset.data();
  out->intervals = d->intervals.data(); out->shape_moments = d->moments.data();
  out->n_compound = (int32_t)s->hierarchy0.size();
  out->compound = d->compound.data(); out->neighbor_offset = d->neighbor_offset.data();
  out->neighbor_id = d->neighbor_id.data();
  out->pts = pts;
  return 0;
}

void vso_dense_destroy(vso_dense* d) { delete d; }

// hand-over state of the last chunk boundary (test tap, see ChunkBoundaryOutput): *maps = int32 [2][h][w]
int vso_dense_last_overlap_state(vso_dense* d, const int32_t** maps, int32_t state[3]) {
  if (d->seg->last_overlap_maps.empty()) return 1;
  *maps = d->seg->last_overlap_maps.data();
  std::memcpy(state, d->seg->last_chain_state, sizeof(int32_t) * 3);
  return 0;
}
int vso_dense_last_chunk_slots(vso_dense* d) { return d->seg->last_slots; }
const int32_t* vso_dense_last_chunk_node_labels(vso_dense* d) { return d->seg->last_node_labels.data(); }
const int32_t* vso_dense_last_chunk_id_images(vso_dense* d) { return d->seg->last_id_images.data(); }
void vso_dense_last_chunk_merge_stats(vso_dense* d, int64_t stats[3]) {
  std::memcpy(stats, d->seg->last_stats, sizeof(int64_t) * 3);
}
void vso_dense_stage_seconds(vso_dense* d, double out[5]) {
  std::memcpy(out, d->seg->stage_sec, sizeof(double) * 5);
}

int vso_segment_chunk_labels(const float* frames, int w, int h, int t, int l1, int min_region_size,
                             int32_t* labels_out) {
  vso::DenseGraph g(w, h, t, l1 != 0, false);
  const size_t fs = (size_t)w * h * 3;
  for (int k = 0; k < t; ++k) {
    g.AddNodesAndSpatialEdges(frames + k * fs);
    if (k > 0) g.AddTemporalEdges(frames + k * fs, frames + (k - 1) * fs);
  }
  g.SegmentFullGraph(min_region_size, true);
  vso::RegionInfoList list;
  vso::RegionInfoPtrMap map;
  g.ObtainResults(&list, &map, nullptr, false, true);
  std::memcpy(labels_out, g.node_labels_after_flatten.data(), sizeof(int32_t) * (size_t)w * h * t);
  return 0;
}

}  // extern "C"
