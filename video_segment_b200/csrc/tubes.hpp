// tubes.hpp -- the host half of EnforceSpatialConnectedness (segmentation/dense_segmentation_graph.h:581-904,
// dense_segmentation_graph.cpp:35-209): which connected pieces of a region follow each other through time ("tubes"),
// which tubes are folded back together, which become regions of their own.  The per-pixel half runs on the device
// (shape.cu): N4 connected components of every region in every frame (K11) and their shape moments (K10).  What
// arrives here is one Piece per component -- frame, moments, a view of its scan intervals -- so this file is
// O(#components) control logic plus, when two pieces of one frame end up in the same tube, the moments of their union.
//
// The decisions are sequential and threshold on floats, so they are taken in the reference's order with its constants:
// match to the closest open tube (flow-displaced centre, size ratio > 0.75, distance < 4 % of the frame diagonal),
// fold tubes that are small (< 20 px per slice) or whose boxes overlap another tube in > 80 % of the common frames into
// the closest tube, join tubes that are temporal neighbours (ratio > 0.9, centres < 20 px apart).
#pragma once
#include <algorithm>
#include <limits>
#include <vector>

#include "shape_math.hpp"

namespace vsbt {

using vsbs::Interval;
using vsbs::Moments;
using vsbs::Shape;
using vsbs::Vec2;

struct Piece {                      // one N4 component of one region in one frame (a RunGroup of shape.cu, pass 0)
  int frame = 0;
  int group = 0;                    // index of the component in the device's group list
  Moments moments;
  const Interval* intervals = nullptr;
  int n_intervals = 0;
};

struct TubeSlice {
  int frame = -1;
  std::vector<int> pieces;          // indices into the region's piece list, ascending (= raster order of their first interval)
  Shape shape;
  std::vector<Interval> merged;     // scan intervals of a slice made of several pieces, in raster order (empty for one piece)
};
typedef std::vector<TubeSlice> Tube;

constexpr float kNowhere = std::numeric_limits<float>::max();

class TubeSplitter {
 public:
  // pieces: all components of ONE region, ascending by (frame, first interval).  flows[frame]: host backward flow of that
  // slot (interleaved x, y; entries may be null), or flows == nullptr.
  TubeSplitter(const std::vector<Piece>& pieces, int width, int height, const std::vector<const float*>* flows)
      : pieces_(pieces), width_(width), flows_(flows), inv_diagonal_(1.0f / std::hypot((float)width, (float)height)) {}

  // The tubes the region falls into after the fold / join passes; empty when it stays whole.
  std::vector<Tube> run() {
    std::vector<Tube> closed, open;
    size_t at = 0;
    while (at < pieces_.size()) {
      const int frame = pieces_[at].frame;
      std::vector<TubeSlice> fresh;
      for (; at < pieces_.size() && pieces_[at].frame == frame; ++at) {
        TubeSlice s;
        s.frame = frame;
        s.pieces.push_back((int)at);
        s.shape = vsbs::shape_from_moments(pieces_[at].moments);
        fresh.push_back(std::move(s));
      }
      if (open.empty()) {
        for (auto& s : fresh) open.push_back(Tube{std::move(s)});
        continue;
      }
      std::vector<Tube> carried;
      std::vector<char> continued(open.size(), 0);
      for (auto& s : fresh) {
        float dist = kNowhere;
        const int prev = closest_open_tube(open, s, &dist);
        bool follows = false;
        if (prev >= 0) {
          const float a = open[prev].back().shape.size, b = s.shape.size;
          const float ratio = std::min(a, b) / (std::max(a, b) + 1e-6);
          follows = ratio > 0.75 && dist * inv_diagonal_ < 0.04f;
        }
        if (follows) {
          continued[prev] = 1;
          open[prev].push_back(std::move(s));
          carried.emplace_back();
          carried.back().swap(open[prev]);          // a tube that was taken leaves `open` empty: no second taker
        } else {
          carried.push_back(Tube{std::move(s)});
        }
      }
      for (size_t k = 0; k < open.size(); ++k)
        if (!continued[k]) closed.push_back(std::move(open[k]));
      open.swap(carried);
    }
    for (auto& t : open) closed.push_back(std::move(t));
    if (closed.size() <= 1) return std::vector<Tube>();
    fold_small_and_overlapping(&closed);
    join_temporal_neighbours(&closed);
    return closed;
  }

 private:
  // FindPreviousTube (:601-628): the open tube whose last slice ended before this frame and lies closest to the
  // slice's centre displaced by the backward flow.  The running best index is kept in a float, as the reference does.
  int closest_open_tube(const std::vector<Tube>& open, const TubeSlice& s, float* dist_out) const {
    Vec2 c = s.shape.center;
    const float* flow = flows_ ? (*flows_)[s.frame] : nullptr;
    if (flow) {
      const float* at = flow + ((size_t)(int)c.y * width_) * 2 + 2 * (int)c.x;
      c.x += at[0];
      c.y += at[1];
    }
    float best = kNowhere, best_idx = -1;
    for (int k = 0; k < (int)open.size(); ++k) {
      if (open[k].empty() || open[k].back().frame >= s.frame) continue;
      const Vec2& o = open[k].back().shape.center;
      const float d = std::hypot(o.y - c.y, o.x - c.x);
      if (d < best) { best = d; best_idx = k; }
    }
    *dist_out = best;
    return (int)best_idx;
  }

  // Scan intervals of a slice in raster order: the piece's own view, or the merged list of a slice of several pieces.
  void intervals_of(const TubeSlice& s, const Interval** first, size_t* n) const {
    if (s.pieces.size() == 1) { *first = pieces_[s.pieces[0]].intervals; *n = (size_t)pieces_[s.pieces[0]].n_intervals; }
    else { *first = s.merged.data(); *n = s.merged.size(); }
  }

  // Union of two slices of one frame: one linear merge of their interval lists (raster order), the moments of the union
  // through one accumulator over the merged list.
  void unite(const TubeSlice& a, const TubeSlice& b, TubeSlice* out) const {
    out->frame = a.frame;
    out->pieces.resize(a.pieces.size() + b.pieces.size());
    std::merge(a.pieces.begin(), a.pieces.end(), b.pieces.begin(), b.pieces.end(), out->pieces.begin());
    const Interval *pa, *pb;
    size_t na, nb;
    intervals_of(a, &pa, &na);
    intervals_of(b, &pb, &nb);
    out->merged.resize(na + nb);
    std::merge(pa, pa + na, pb, pb + nb, out->merged.begin(),
               [](const Interval& p, const Interval& q) { return p.y != q.y ? p.y < q.y : p.lx < q.lx; });
    vsbs::MomentSum sum;
    for (const Interval& iv : out->merged) sum.add(iv.y, iv.lx, iv.rx);
    out->shape = vsbs::shape_from_moments(sum.mean());
  }

  // `from` folded into `into` (into's slices first where both have one for a frame; the union's shape is recomputed).
  // Both tubes are consumed: the caller replaces one with the result and drops the other.
  void fold(Tube& into, Tube& from, Tube* out) const {
    if (into.empty()) { out->swap(from); return; }
    if (from.empty()) { out->swap(into); return; }
    out->reserve(into.size() + from.size());
    size_t i = 0, j = 0;
    while (i < into.size() && j < from.size()) {
      if (into[i].frame < from[j].frame) out->push_back(std::move(into[i++]));
      else if (into[i].frame > from[j].frame) out->push_back(std::move(from[j++]));
      else {
        TubeSlice both;
        unite(into[i], from[j], &both);
        out->push_back(std::move(both));
        ++i; ++j;
      }
    }
    for (; i < into.size(); ++i) out->push_back(std::move(into[i]));
    for (; j < from.size(); ++j) out->push_back(std::move(from[j]));
  }

  // Calls visit(slice_a, slice_b) for every frame both tubes have; returns the number of such frames.
  template <class F>
  static int over_common_frames(const Tube& a, const Tube& b, F visit) {
    size_t i = 0, j = 0;
    int n = 0;
    while (i < a.size() && j < b.size()) {
      if (a[i].frame < b[j].frame) ++i;
      else if (a[i].frame > b[j].frame) ++j;
      else { visit(a[i], b[j]); ++n; ++i; ++j; }
    }
    return n;
  }

  static float mean_centre_distance(const Tube& a, const Tube& b) {
    float sum = 0;
    const int n = over_common_frames(a, b, [&](const TubeSlice& p, const TubeSlice& q) {
      sum += std::hypot(p.shape.center.y - q.shape.center.y, p.shape.center.x - q.shape.center.x);
    });
    return n > 0 ? sum / n : kNowhere;
  }

  static float box_overlap_fraction(const Tube& a, const Tube& b) {
    int hits = 0;
    const int n = over_common_frames(a, b, [&](const TubeSlice& p, const TubeSlice& q) {
      Vec2 bp[4], bq[4];
      vsbs::shape_box(p.shape, 10, bp);
      vsbs::shape_box(q.shape, 10, bq);
      if (vsbs::boxes_intersect(bp, bq)) ++hits;
    });
    return n > 0 ? hits * (1.0f / n) : kNowhere;
  }

  static float mean_slice_size(const Tube& t) {
    if (t.empty()) return 0;
    float total = 0;
    for (const auto& s : t) total += s.shape.size;
    return total / t.size();
  }

  static bool temporal_neighbours(const Tube& a, const Tube& b) {
    if (a.empty() || b.empty()) return false;
    const Shape *p, *q;
    if (a.front().frame - 1 == b.back().frame) { p = &a.front().shape; q = &b.back().shape; }
    else if (a.back().frame + 1 == b.front().frame) { p = &a.back().shape; q = &b.front().shape; }
    else return false;
    const float ratio = std::min(p->size, q->size) * (1.0f / std::max(p->size, q->size));
    return ratio > 0.9 && std::hypot(p->center.y - q->center.y, p->center.x - q->center.x) < 20;
  }

  void fold_small_and_overlapping(std::vector<Tube>* tubes) const {          // :779-800
    for (int k = 0; k < (int)tubes->size();) {
      bool fold_it = mean_slice_size((*tubes)[k]) < 20;
      for (int l = 0; !fold_it && l < (int)tubes->size(); ++l)
        fold_it = l != k && box_overlap_fraction((*tubes)[k], (*tubes)[l]) > 0.8;   // no common frame counts as overlap, as in the reference
      int target = -1;
      if (fold_it) {
        float best = kNowhere;
        for (int l = 0; l < (int)tubes->size(); ++l) {
          if (l == k) continue;
          const float d = mean_centre_distance((*tubes)[k], (*tubes)[l]);
          if (d < best) { best = d; target = l; }
        }
      }
      if (target < 0) { ++k; continue; }
      Tube merged;
      fold((*tubes)[target], (*tubes)[k], &merged);
      (*tubes)[target].swap(merged);
      tubes->erase(tubes->begin() + k);
    }
  }

  void join_temporal_neighbours(std::vector<Tube>* tubes) const {            // :802-823
    for (int k = 0; k < (int)tubes->size();) {
      int target = -1;
      for (int l = 0; l < (int)tubes->size() && target < 0; ++l)
        if (l != k && temporal_neighbours((*tubes)[k], (*tubes)[l])) target = l;
      if (target < 0) { ++k; continue; }
      Tube merged;
      fold((*tubes)[k], (*tubes)[target], &merged);
      (*tubes)[target].swap(merged);
      tubes->erase(tubes->begin() + k);
    }
  }

  const std::vector<Piece>& pieces_;
  int width_;
  const std::vector<const float*>* flows_;
  float inv_diagonal_;
};

}  // namespace vsbt
