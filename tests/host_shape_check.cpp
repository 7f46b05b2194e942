// TEST INFRASTRUCTURE ONLY.  The product's host-side shape arithmetic (video_segment_b200/csrc/shape_math.hpp: the
// moment accumulator the device kernel of shape.cu and the host share, shape descriptors, oriented boxes;
// csrc/region_raster.hpp: raster union of the hierarchical stage) against the REFERENCE's own functions of
// segment_util/segmentation_util.cpp (compiled unmodified in oracle/_ref/libref_results.so) on random rasters:
// ShapeMomentsFromRasterization, GetShapeDescriptorFromShapeMoment, MergeRasterization, RasterizationArea,
// ShapeDescriptorBox / ShapeDescriptorBoxesIntersect.  Floats are compared by bits.  CPU test.  (Connected components
// and the per-region moments of the dense engine run on the device: tests/test_gpu_kernels.py.)
#include <stdint.h>
#include <string.h>

#include <random>
#include <string>
#include <vector>

#include <opencv2/core/core.hpp>

#include "segment_util/segmentation_util.h"

#include "../video_segment_b200/csrc/region_raster.hpp"
#include "../video_segment_b200/csrc/shape_math.hpp"

namespace {

using segmentation::Rasterization;

bool SameBits(float a, float b) { return memcmp(&a, &b, 4) == 0; }

// Random pixel blobs on a w x h grid -> scan intervals.  `part` (0 / 1) keeps the pixels of one half of a random
// two-colouring, so that part 0 and part 1 are disjoint rasters of the same mask.
void RandomRaster(std::mt19937& rng, int w, int h, std::vector<uint8_t>* mask, std::vector<uint8_t>* colour) {
  mask->assign((size_t)w * h, 0);
  colour->assign((size_t)w * h, 0);
  std::uniform_int_distribution<int> nb(1, 6), px(0, w - 1), py(0, h - 1), rad(1, 9), coin(0, 1);
  const int blobs = nb(rng);
  for (int b = 0; b < blobs; ++b) {
    const int cx = px(rng), cy = py(rng), rx = rad(rng), ry = rad(rng), c = coin(rng);
    for (int y = std::max(0, cy - ry); y <= std::min(h - 1, cy + ry); ++y)
      for (int x = std::max(0, cx - rx); x <= std::min(w - 1, cx + rx); ++x) {
        const float dx = (x - cx) / (float)rx, dy = (y - cy) / (float)ry;
        if (dx * dx + dy * dy <= 1.0f && (rng() % 11) != 0) {
          (*mask)[(size_t)y * w + x] = 1;
          (*colour)[(size_t)y * w + x] = (uint8_t)c;
        }
      }
  }
}

void ToRasters(const std::vector<uint8_t>& mask, const std::vector<uint8_t>& colour, int part, int w, int h,
               vsbr::Raster* mine, Rasterization* ref) {
  mine->clear();
  ref->Clear();
  for (int y = 0; y < h; ++y) {
    int x = 0;
    while (x < w) {
      auto on = [&](int xx) { return mask[(size_t)y * w + xx] && (part < 0 || colour[(size_t)y * w + xx] == part); };
      if (!on(x)) { ++x; continue; }
      int e = x;
      while (e + 1 < w && on(e + 1)) ++e;
      mine->push_back({y, x, e});
      auto* s = ref->add_scan_inter();
      s->set_y(y);
      s->set_left_x(x);
      s->set_right_x(e);
      x = e + 1;
    }
  }
}

bool SameRaster(const vsbr::Raster& a, const Rasterization& b) {
  if ((int)a.size() != b.scan_inter_size()) return false;
  for (int i = 0; i < (int)a.size(); ++i)
    if (a[i].y != b.scan_inter(i).y() || a[i].lx != b.scan_inter(i).left_x() || a[i].rx != b.scan_inter(i).right_x()) return false;
  return true;
}

}  // namespace

extern "C" int host_shape_check(unsigned seed, int n_cases, char* msg, int msg_cap) {
  std::mt19937 rng(seed);
  std::string first;
  int bad = 0;
  auto fail = [&](int c, const char* what) {
    if (first.empty()) first = "case " + std::to_string(c) + ": " + what;
    ++bad;
  };
  std::vector<vsbs::Shape> shapes_mine;
  std::vector<segmentation::ShapeDescriptor> shapes_ref;
  for (int c = 0; c < n_cases; ++c) {
    const int w = 8 + rng() % 56, h = 8 + rng() % 40;
    std::vector<uint8_t> mask, colour;
    RandomRaster(rng, w, h, &mask, &colour);
    vsbr::Raster all, p0, p1;
    Rasterization rall, r0, r1;
    ToRasters(mask, colour, -1, w, h, &all, &rall);
    ToRasters(mask, colour, 0, w, h, &p0, &r0);
    ToRasters(mask, colour, 1, w, h, &p1, &r1);
    if (all.empty()) continue;
    if (vsbr::raster_area(all) != segmentation::RasterizationArea(rall)) fail(c, "raster_area");
    // moments + shape descriptor
    segmentation::ShapeMoments rm;
    segmentation::ShapeMomentsFromRasterization(rall, &rm);
    const vsbs::Moments mm = vsbr::moments_of(all);
    if (!SameBits(mm.size, rm.size()) || !SameBits(mm.mean_x, rm.mean_x()) || !SameBits(mm.mean_y, rm.mean_y()) ||
        !SameBits(mm.xx, rm.moment_xx()) || !SameBits(mm.xy, rm.moment_xy()) || !SameBits(mm.yy, rm.moment_yy()))
      fail(c, "moments_of");
    segmentation::ShapeDescriptor rs;
    segmentation::GetShapeDescriptorFromShapeMoment(rm, &rs);
    const vsbs::Shape ms = vsbs::shape_from_moments(mm);
    if (!SameBits(ms.center.x, rs.center.x) || !SameBits(ms.center.y, rs.center.y) || !SameBits(ms.mag_major, rs.mag_major) ||
        !SameBits(ms.mag_minor, rs.mag_minor) || !SameBits(ms.dir_major.x, rs.dir_major.x) || !SameBits(ms.dir_major.y, rs.dir_major.y) ||
        !SameBits(ms.dir_minor.x, rs.dir_minor.x) || !SameBits(ms.dir_minor.y, rs.dir_minor.y))
      fail(c, "shape_of");
    shapes_mine.push_back(ms);
    shapes_ref.push_back(rs);
    // merging the two disjoint halves gives back the whole, interval for interval like the reference
    if (!p0.empty() && !p1.empty()) {
      vsbr::Raster merged;
      vsbr::merge_rasters(p0, p1, &merged);
      Rasterization rmerged;
      segmentation::MergeRasterization(r0, r1, &rmerged);
      if (!SameRaster(merged, rmerged)) fail(c, "merge_rasters");
    }
  }
  // oriented boxes of consecutive shapes: corners and the intersection verdict
  for (size_t k = 0; k + 1 < shapes_mine.size(); ++k) {
    for (float border : {0.0f, 2.0f}) {
      vsbs::Vec2 a[4], b[4];
      vsbs::shape_box(shapes_mine[k], border, a);
      vsbs::shape_box(shapes_mine[k + 1], border, b);
      std::vector<cv::Point2f> ra, rb;
      segmentation::ShapeDescriptorBox(shapes_ref[k], border, &ra);
      segmentation::ShapeDescriptorBox(shapes_ref[k + 1], border, &rb);
      for (int i = 0; i < 4; ++i)
        if (!SameBits(a[i].x, ra[i].x) || !SameBits(a[i].y, ra[i].y)) { fail((int)k, "shape_box"); break; }
      if (vsbs::boxes_intersect(a, b) != segmentation::ShapeDescriptorBoxesIntersect(ra, rb)) fail((int)k, "boxes_intersect");
    }
  }
  if (msg && msg_cap > 0) {
    strncpy(msg, first.c_str(), msg_cap - 1);
    msg[msg_cap - 1] = 0;
  }
  return bad;
}
