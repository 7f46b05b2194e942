// host_shape.hpp -- host-side region bookkeeping of the streaming engine: scan-interval
// rasters, shape moments, N4 connected components and the spatio-temporal "tube" split of
// DenseSegmentationGraph::EnforceSpatialConnectedness.  This is O(#scan intervals) control
// logic that the reference also keeps on the host (protobuf objects); all per-pixel work
// (labels, N4 fix, RLE, relabel, neighbour pairs) runs in CUDA (results.cu).
//
// Reference: segmentation/dense_segmentation_graph.h:581-904, dense_segmentation_graph.cpp:35-209,
// segment_util/segmentation_util.cpp:243-410,484-570,644-693,1007-1101.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <limits>
#include <memory>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace vsbh {

struct Interval { int y, lx, rx; };
typedef std::vector<Interval> Raster;                       // lexicographic (y, lx)
struct Slice { int frame; std::shared_ptr<Raster> raster; };
typedef std::vector<Slice> Raster3D;                        // ascending frame

struct Moments { float size = 0, mx = 0, my = 0, xx = 0, xy = 0, yy = 0; };
struct Vec2 { float x = 0, y = 0; };
struct Shape {                                              // segmentation_util.h:137-150
  Vec2 center;
  float mag_major = 0, mag_minor = 0;
  Vec2 dir_major{1.f, 0.f}, dir_minor{0.f, 1.f};
  int size = 0;
};

// One over-segmentation region of the current chunk (RegionInformation, segmentation_common.h:39-116).
struct Region {
  int index = -1;          // position in the chunk's region list (first-seen order)
  int label = -1;          // device label (representative node id, or fresh id for split tubes)
  int size = 0;
  int constrained_id = -1;
  int region_id = -1;
  bool removed = false;    // FLAGGED_FOR_REMOVAL
  std::vector<int> neighbors;   // sorted region indices
  Raster3D raster;
};

inline int raster_area(const Raster& r) {
  int a = 0;
  for (const auto& s : r) a += s.rx - s.lx + 1;
  return a;
}

// ShapeMomentsFromRasterization (segmentation_util.cpp:652-693)
inline Moments moments_of(const Raster& raster) {
  float mean_x = 0, mean_y = 0, mxx = 0, myy = 0, mxy = 0, area = 0;
  for (const auto& s : raster) {
    const float m = s.lx, n = s.rx, cy = s.y;
    const float len = (n - m + 1);
    area += len;
    const float cx = (n + m) * 0.5;
    const float sum_x = cx * len, sum_y = cy * len;
    mean_x += sum_x;
    mean_y += sum_y;
    mxy += cy * sum_x;
    myy += cy * sum_y;
    mxx += len * (-m + 2 * m * m + n + 2 * m * n + 2 * n * n) / 6.0f;
  }
  const float inv = 1.0f / area;
  Moments o;
  o.size = area; o.mx = mean_x * inv; o.my = mean_y * inv; o.xx = mxx * inv; o.xy = mxy * inv; o.yy = myy * inv;
  return o;
}

// GetShapeDescriptorFromShapeMoment (segmentation_util.cpp:243-345)
inline Shape shape_of(const Moments& mo) {
  Shape sd;
  float area_sum = 0;
  const float area = mo.size;
  area_sum += area;
  float x = mo.mx * area, y = mo.my * area, xx = mo.xx * area, xy = mo.xy * area, yy = mo.yy * area;
  const float inv = 1.0f / area_sum;
  x *= inv; y *= inv; xx *= inv; xy *= inv; yy *= inv;
  sd.center = Vec2{x, y};
  sd.size = area_sum;
  if (area_sum < 10) return sd;
  const float var_xx = xx - x * x, var_xy = xy - x * y, var_yy = yy - y * y;
  const float trace = var_xx + var_yy;
  const float det = var_xx * var_yy - var_xy * var_xy;
  float disc = 0.25 * trace * trace - det;
  disc = std::max(0.0f, disc);
  const float sq = std::sqrt(disc);
  const float e1 = trace * 0.5 - sq, e2 = trace * 0.5 + sq;
  if (std::min(std::fabs(e1), std::fabs(e2)) < 1) return sd;
  Vec2 ev1{1.f, 0.f}, ev2{0.f, 1.f};
  const Vec2 v1{e1 - var_yy, var_xy}, v2{e2 - var_yy, var_xy};
  const float n1 = std::hypot(v1.y, v1.x), n2 = std::hypot(v2.y, v2.x);
  if (n1 > 1e-6f && n2 > 1e-6f && disc > 0.1) {
    const float s1 = 1.0f / n1, s2 = 1.0f / n2;
    ev1 = Vec2{v1.x * s1, v1.y * s1};
    ev2 = Vec2{v2.x * s2, v2.y * s2};
  }
  float sg1 = std::sqrt(std::fabs(e1)), sg2 = std::sqrt(std::fabs(e2));
  if (sg1 < sg2) { std::swap(sg1, sg2); std::swap(ev1, ev2); }
  const Vec2 nrm{-ev1.y, ev1.x};
  if (ev2.x * nrm.x + ev2.y * nrm.y < 0) ev2 = Vec2{-ev2.x, -ev2.y};
  sd.mag_major = sg1; sd.mag_minor = sg2; sd.dir_major = ev1; sd.dir_minor = ev2;
  return sd;
}

// MergeRasterization (segmentation_util.cpp:484-570)
inline void merge_rasters(const Raster& a, const Raster& b, Raster* out) {
  size_t i = 0, j = 0;
  Raster m;
  std::vector<int> offs;
  while (i < a.size() || j < b.size()) {
    const int ay = i < a.size() ? a[i].y : 1 << 30, by = j < b.size() ? b[j].y : 1 << 30;
    if (ay < by) m.push_back(a[i++]);
    else if (by < ay) m.push_back(b[j++]);
    else {
      offs.clear();
      while (true) {
        const bool lc = i < a.size() && a[i].y == ay, rc = j < b.size() && b[j].y == by;
        if (!lc && !rc) break;
        const int lx = lc ? a[i].lx : std::numeric_limits<int>::max();
        const int rx = rc ? b[j].lx : std::numeric_limits<int>::max();
        if (lx < rx) { offs.push_back(a[i].lx); offs.push_back(a[i].rx); ++i; }
        else { offs.push_back(b[j].lx); offs.push_back(b[j].rx); ++j; }
      }
      int k = 0, l = 0;
      const int n = (int)offs.size();
      while (k < n) {
        if (k + 2 == n) { m.push_back({ay, offs[l], offs[k + 1]}); break; }
        if (offs[k + 2] - 1 == offs[k + 1]) { k += 2; }
        else { m.push_back({ay, offs[l], offs[k + 1]}); k += 2; l = k; }
      }
    }
  }
  out->swap(m);
}

// ConnectedComponents(raster, N4_CONNECT) (segmentation_util.cpp:1007-1101): components in order of
// their first scan interval.
inline int components_n4(const Raster& raster, std::vector<Raster>* comps) {
  const int n = (int)raster.size();
  std::vector<int> par(n);
  std::iota(par.begin(), par.end(), 0);
  auto find = [&](int x) { while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; } return x; };
  int last_change = -1, last_y = -2, test = 0;
  for (int i = 0; i < n; ++i) {
    const Interval& c = raster[i];
    if (c.y != last_y) {
      test = (last_y + 1 == c.y) ? last_change : i;
      last_y = c.y;
      last_change = i;
    }
    for (int k = test; k < i; ++k) {
      const Interval& o = raster[k];
      if (std::abs(c.y - o.y) <= 1 && std::max(c.lx, o.lx) <= std::min(c.rx, o.rx)) {
        const int a = find(i), b = find(k);
        if (a != b) par[std::max(a, b)] = std::min(a, b);
      }
    }
  }
  int nc = 0;
  for (int i = 0; i < n; ++i) nc += (find(i) == i);
  if (nc == 1) { comps->push_back(raster); return 1; }
  std::unordered_map<int, int> rep2c;
  for (int i = 0; i < n; ++i) {
    const int r = find(i);
    auto it = rep2c.find(r);
    if (it == rep2c.end()) { rep2c[r] = (int)comps->size(); comps->push_back(Raster{raster[i]}); }
    else (*comps)[it->second].push_back(raster[i]);
  }
  return nc;
}

struct TubeSlice {
  int frame = -1;
  Raster raster;
  Shape shape;
  void update_shape() { shape = shape_of(moments_of(raster)); }
};
typedef std::vector<TubeSlice> Tube;

inline float tube_avg_slice_size(const Tube& t) {
  if (t.empty()) return 0;
  float a = 0;
  for (const auto& s : t) a += s.shape.size;
  return a / t.size();
}

inline void tube_merge(const Tube& a, const Tube& b, Tube* out) {
  if (a.empty()) { *out = b; return; }
  if (b.empty()) { *out = a; return; }
  size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    if (a[i].frame < b[j].frame) out->push_back(a[i++]);
    else if (a[i].frame > b[j].frame) out->push_back(b[j++]);
    else {
      TubeSlice m = a[i];
      merge_rasters(m.raster, b[j].raster, &m.raster);
      m.update_shape();
      out->push_back(m);
      ++i; ++j;
    }
  }
  while (i < a.size()) out->push_back(a[i++]);
  while (j < b.size()) out->push_back(b[j++]);
}

inline bool tubes_temporal_neighbors(const Tube& a, const Tube& b) {
  if (a.empty() || b.empty()) return false;
  Shape p, q;
  if (a[0].frame - 1 == b.back().frame) { p = a[0].shape; q = b.back().shape; }
  else if (a.back().frame + 1 == b[0].frame) { p = a.back().shape; q = b[0].shape; }
  else return false;
  const float ratio = std::min(p.size, q.size) * (1.0f / std::max(p.size, q.size));
  const float dx = p.center.x - q.center.x, dy = p.center.y - q.center.y;
  return ratio > 0.9 && std::hypot(dy, dx) < 20;
}

inline float tube_avg_distance(const Tube& a, const Tube& b) {
  if (a.empty() || b.empty()) return std::numeric_limits<float>::max();
  const int f0 = std::max(a[0].frame, b[0].frame), f1 = std::min(a.back().frame, b.back().frame);
  int i = 0, j = 0, wgt = 0;
  float sum = 0;
  for (int f = f0; f <= f1; ++f) {
    while (a[i].frame < f) ++i;
    while (b[j].frame < f) ++j;
    if (a[i].frame != f || b[j].frame != f) continue;
    const float dx = a[i].shape.center.x - b[j].shape.center.x, dy = a[i].shape.center.y - b[j].shape.center.y;
    sum += std::hypot(dy, dx);
    ++wgt;
  }
  return wgt > 0 ? sum / wgt : std::numeric_limits<float>::max();
}

inline void shape_box(const Shape& s, float border, Vec2 c[4]) {          // segmentation_util.cpp:364-379
  const float ma = s.mag_major * 1.65f + border, mi = s.mag_minor * 1.65f + border;
  const Vec2 mj{s.dir_major.x * ma, s.dir_major.y * ma}, mn{s.dir_minor.x * mi, s.dir_minor.y * mi};
  c[0] = Vec2{s.center.x - mj.x + mn.x, s.center.y - mj.y + mn.y};
  c[1] = Vec2{s.center.x - mj.x - mn.x, s.center.y - mj.y - mn.y};
  c[2] = Vec2{s.center.x + mj.x - mn.x, s.center.y + mj.y - mn.y};
  c[3] = Vec2{s.center.x + mj.x + mn.x, s.center.y + mj.y + mn.y};
}

inline bool boxes_intersect(const Vec2 a[4], const Vec2 b[4]) {           // segmentation_util.cpp:381-410
  for (int k = 0; k < 4; ++k) {
    const double adx = (float)(a[(k + 1) % 4].x - a[k].x), ady = (float)(a[(k + 1) % 4].y - a[k].y);
    for (int l = 0; l < 4; ++l) {
      const double bdx = (float)(b[(l + 1) % 4].x - b[l].x), bdy = (float)(b[(l + 1) % 4].y - b[l].y);
      const double ddx = (float)(b[l].x - a[k].x), ddy = (float)(b[l].y - a[k].y);
      const double kross = adx * bdy - ady * bdx;
      if (std::fabs(kross) < 1e-6) continue;
      const float inv = 1.0f / kross;
      const double t = (ddx * bdy - ddy * bdx) * inv, s = (ddx * ady - ddy * adx) * inv;
      if (t > -1e-6f && t < 1.0f + 1e-6f && s > -1e-6f && s < 1.0f + 1e-6f) return true;
    }
  }
  return false;
}

inline float tube_intersection(const Tube& a, const Tube& b) {
  if (a.empty() || b.empty()) return std::numeric_limits<float>::max();
  const int f0 = std::max(a[0].frame, b[0].frame), f1 = std::min(a.back().frame, b.back().frame);
  int i = 0, j = 0, hit = 0, wgt = 0;
  for (int f = f0; f <= f1; ++f) {
    while (a[i].frame < f) ++i;
    while (b[j].frame < f) ++j;
    if (a[i].frame != f || b[j].frame != f) continue;
    Vec2 ba[4], bb[4];
    shape_box(a[i].shape, 10, ba);
    shape_box(b[j].shape, 10, bb);
    if (boxes_intersect(ba, bb)) ++hit;
    ++wgt;
  }
  return wgt > 0 ? hit * (1.0f / wgt) : std::numeric_limits<float>::max();
}

inline int closest_tube(const Tube& t, const std::vector<Tube>& tubes, int ignore) {
  float best = std::numeric_limits<float>::max();
  int idx = -1;
  for (int k = 0; k < (int)tubes.size(); ++k) {
    if (k == ignore) continue;
    const float d = tube_avg_distance(t, tubes[k]);
    if (d < best) { best = d; idx = k; }
  }
  return idx;
}

// Splits one region into spatially connected tubes (EnforceSpatialConnectedness, :666-850).
// Returns the tubes after the reference's merge heuristics; empty if the region stays whole.
// flows[frame] (nullable entries) = host backward flow of that slot, interleaved x,y.
inline std::vector<Tube> split_region_into_tubes(const Raster3D& raster, int w, int h,
                                                 const std::vector<const float*>* flows) {
  std::vector<Tube> result, active;
  const float inv_diam = 1.0f / std::hypot((float)w, (float)h);
  for (const auto& sl : raster) {
    const int frame = sl.frame;
    std::vector<Raster> comps;
    components_n4(*sl.raster, &comps);
    std::vector<TubeSlice> slices;
    slices.reserve(comps.size());
    for (auto& c : comps) {
      TubeSlice s;
      s.frame = frame;
      s.raster.swap(c);
      s.update_shape();
      slices.push_back(std::move(s));
    }
    if (active.empty()) {
      for (auto& s : slices) active.push_back(Tube{std::move(s)});
      continue;
    }
    std::vector<Tube> next;
    std::vector<int> used(active.size(), 0);
    const float* flow = flows ? (*flows)[frame] : nullptr;
    for (auto& s : slices) {
      // FindPreviousTube (:601-628)
      Vec2 pc = s.shape.center;
      if (flow) {
        const float* fp = flow + ((size_t)(int)pc.y * w) * 2 + 2 * (int)pc.x;
        pc.x += fp[0];
        pc.y += fp[1];
      }
      float cd = std::numeric_limits<float>::max();
      float ci = -1;
      for (int k = 0; k < (int)active.size(); ++k) {
        if (active[k].empty() || active[k].back().frame >= frame) continue;
        const float dx = active[k].back().shape.center.x - pc.x, dy = active[k].back().shape.center.y - pc.y;
        const float d = std::hypot(dy, dx);
        if (d < cd) { cd = d; ci = k; }
      }
      const int prev = (int)ci;
      if (prev < 0) { next.push_back(Tube{std::move(s)}); continue; }
      const float ratio = std::min(active[prev].back().shape.size, s.shape.size) /
                          (std::max(active[prev].back().shape.size, s.shape.size) + 1e-6);
      if (ratio > 0.75 && cd * inv_diam < 0.04f) {
        ++used[prev];
        active[prev].push_back(std::move(s));
        next.push_back(Tube());
        next.back().swap(active[prev]);
      } else {
        next.push_back(Tube{std::move(s)});
      }
    }
    for (size_t k = 0; k < active.size(); ++k)
      if (used[k] == 0) result.push_back(std::move(active[k]));
    next.swap(active);
  }
  for (auto& t : active) result.push_back(std::move(t));
  if (result.size() <= 1) return std::vector<Tube>();

  for (int k = 0; k < (int)result.size();) {              // small or overlapping tubes (:779-800)
    bool merge = tube_avg_slice_size(result[k]) < 20;
    if (!merge) {
      for (int l = 0; l < (int)result.size(); ++l) {
        if (l == k) continue;
        if (tube_intersection(result[k], result[l]) > 0.8) { merge = true; break; }
      }
    }
    bool done = false;
    if (merge) {
      const int idx = closest_tube(result[k], result, k);
      if (idx >= 0) {
        Tube m;
        tube_merge(result[idx], result[k], &m);
        result[idx].swap(m);
        result.erase(result.begin() + k);
        done = true;
      }
    }
    if (!done) ++k;
  }
  for (int k = 0; k < (int)result.size();) {              // temporal neighbours (:802-823)
    bool merged = false;
    for (int l = 0; l < (int)result.size(); ++l) {
      if (l == k) continue;
      if (tubes_temporal_neighbors(result[k], result[l])) {
        Tube m;
        tube_merge(result[k], result[l], &m);
        result[l].swap(m);
        result.erase(result.begin() + k);
        merged = true;
        break;
      }
    }
    if (!merged) ++k;
  }
  return result;
}

}  // namespace vsbh
