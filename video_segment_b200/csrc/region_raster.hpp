// region_raster.hpp -- scan-interval rasters of the hierarchical stage's base level (host side of region_stage.cu):
// a region's raster per frame, the union of two disjoint rasters when two base regions are merged
// (MergeRasterization / MergeRasterization3D, segment_util/segmentation_util.cpp:484-642) and the ShapeMoments of a raster
// for the result record (shape_math.hpp's accumulator, the same one the dense engine's device kernel runs).
#pragma once
#include <algorithm>
#include <iterator>
#include <memory>
#include <vector>

#include "shape_math.hpp"

namespace vsbr {

using vsbs::Interval;
typedef std::vector<Interval> Raster;                       // ascending (y, lx)
struct Slice { int frame; std::shared_ptr<Raster> raster; };
typedef std::vector<Slice> Raster3D;                        // ascending frame

inline int raster_area(const Raster& r) {
  int px = 0;
  for (const Interval& s : r) px += s.rx - s.lx + 1;
  return px;
}

// Union of two disjoint rasters: all intervals in raster order; where both contribute to a row, intervals that touch
// (next.lx == previous.rx + 1) become one.
inline void merge_rasters(const Raster& a, const Raster& b, Raster* out) {
  Raster all;
  all.reserve(a.size() + b.size());
  std::merge(a.begin(), a.end(), b.begin(), b.end(), std::back_inserter(all),
             [](const Interval& p, const Interval& q) { return p.y != q.y ? p.y < q.y : p.lx < q.lx; });
  Raster joined;
  joined.reserve(all.size());
  size_t ia = 0, ib = 0;
  for (size_t k = 0; k < all.size();) {
    const int y = all[k].y;
    size_t e = k;
    while (e < all.size() && all[e].y == y) ++e;
    // rows that only one side contributes to are taken as they are
    while (ia < a.size() && a[ia].y < y) ++ia;
    while (ib < b.size() && b[ib].y < y) ++ib;
    const bool shared_row = ia < a.size() && a[ia].y == y && ib < b.size() && b[ib].y == y;
    for (size_t i = k; i < e; ++i) {
      if (shared_row && i > k && all[i].lx - 1 == joined.back().rx) joined.back().rx = all[i].rx;
      else joined.push_back(all[i]);
    }
    k = e;
  }
  out->swap(joined);
}

inline vsbs::Moments moments_of(const Raster& r) {
  vsbs::MomentSum sum;
  for (const Interval& s : r) sum.add(s.y, s.lx, s.rx);
  return sum.mean();
}

}  // namespace vsbr
