// shape_math.hpp -- the float arithmetic of the shape fields of the output message, shared by the device kernels
// (shape.cu) and the host (tubes.hpp, region_stage.cu): running shape moments over scan intervals, the shape
// descriptor of a set of moments, oriented boxes.  ShapeMoments are floats in segmentation.proto and the tube
// heuristics threshold on values derived from them, so every expression below is evaluated in the reference's
// operand order (segment_util/segmentation_util.cpp:243-410, 652-693); device code is built with -fmad=false and host
// code with -ffp-contract=off so that neither fuses a*b+c.
#pragma once
#include <cmath>
#include <utility>

#if defined(__CUDACC__)
#define VSB_HD __host__ __device__ __forceinline__
#else
#define VSB_HD inline
#endif

namespace vsbs {

struct Interval { int y, lx, rx; };                          // one scan interval, the layout of the device's int3 records
struct Moments { float size = 0, mean_x = 0, mean_y = 0, xx = 0, xy = 0, yy = 0; };

// Running sums over scan intervals taken in raster order (y, then left_x).  One accumulator per group of intervals;
// add() once per interval in order, mean() at the end.
struct MomentSum {
  float sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0, area = 0;
  VSB_HD void add(int y, int lx, int rx) {
    const float m = (float)lx, n = (float)rx, cy = (float)y;
    const float len = n - m + 1;
    area += len;
    const float cx = (n + m) * 0.5f;                         // exact: n + m < 2^24
    const float row_x = cx * len, row_y = cy * len;
    sx += row_x;
    sy += row_y;
    sxy += cy * row_x;
    syy += cy * row_y;
    sxx += len * (-m + 2 * m * m + n + 2 * m * n + 2 * n * n) / 6.0f;
  }
  VSB_HD Moments mean() const {
    const float inv = 1.0f / area;
    Moments o;
    o.size = area; o.mean_x = sx * inv; o.mean_y = sy * inv; o.xx = sxx * inv; o.xy = sxy * inv; o.yy = syy * inv;
    return o;
  }
};

struct Vec2 { float x = 0, y = 0; };
struct Shape {                                               // ShapeDescriptor, segmentation_util.h:137-150
  Vec2 center;
  float mag_major = 0, mag_minor = 0;
  Vec2 dir_major{1.f, 0.f}, dir_minor{0.f, 1.f};
  int size = 0;
};

// Centre, principal axes and their magnitudes (square roots of the covariance eigenvalues) of one set of moments.
inline Shape shape_from_moments(const Moments& mo) {
  Shape sd;
  const float total = 0 + mo.size;
  const float inv = 1.0f / total;
  const float x = mo.mean_x * mo.size * inv, y = mo.mean_y * mo.size * inv;
  const float xx = mo.xx * mo.size * inv, xy = mo.xy * mo.size * inv, yy = mo.yy * mo.size * inv;
  sd.center = Vec2{x, y};
  sd.size = (int)total;
  if (total < 10) return sd;                                 // too small for axes
  const float cxx = xx - x * x, cxy = xy - x * y, cyy = yy - y * y;
  const float trace = cxx + cyy;
  const float det = cxx * cyy - cxy * cxy;
  float disc = (float)(0.25 * trace * trace - det);
  disc = disc > 0.0f ? disc : 0.0f;
  const float root = std::sqrt(disc);
  const float lam_lo = (float)(trace * 0.5 - root), lam_hi = (float)(trace * 0.5 + root);
  if (std::fmin(std::fabs(lam_lo), std::fabs(lam_hi)) < 1) return sd;
  Vec2 axis_lo{1.f, 0.f}, axis_hi{0.f, 1.f};
  const Vec2 raw_lo{lam_lo - cyy, cxy}, raw_hi{lam_hi - cyy, cxy};
  const float len_lo = std::hypot(raw_lo.y, raw_lo.x), len_hi = std::hypot(raw_hi.y, raw_hi.x);
  if (len_lo > 1e-6f && len_hi > 1e-6f && disc > 0.1) {
    const float s_lo = 1.0f / len_lo, s_hi = 1.0f / len_hi;
    axis_lo = Vec2{raw_lo.x * s_lo, raw_lo.y * s_lo};
    axis_hi = Vec2{raw_hi.x * s_hi, raw_hi.y * s_hi};
  }
  float mag_a = std::sqrt(std::fabs(lam_lo)), mag_b = std::sqrt(std::fabs(lam_hi));
  if (mag_a < mag_b) { std::swap(mag_a, mag_b); std::swap(axis_lo, axis_hi); }
  const Vec2 normal{-axis_lo.y, axis_lo.x};                  // minor axis on the left of the major one
  if (axis_hi.x * normal.x + axis_hi.y * normal.y < 0) axis_hi = Vec2{-axis_hi.x, -axis_hi.y};
  sd.mag_major = mag_a; sd.mag_minor = mag_b; sd.dir_major = axis_lo; sd.dir_minor = axis_hi;
  return sd;
}

// Corners of the oriented box of a shape, 1.65 sigma plus a border (segmentation_util.cpp:364-379).
inline void shape_box(const Shape& s, float border, Vec2 c[4]) {
  const float ext_a = s.mag_major * 1.65f + border, ext_b = s.mag_minor * 1.65f + border;
  const Vec2 a{s.dir_major.x * ext_a, s.dir_major.y * ext_a}, b{s.dir_minor.x * ext_b, s.dir_minor.y * ext_b};
  c[0] = Vec2{s.center.x - a.x + b.x, s.center.y - a.y + b.y};
  c[1] = Vec2{s.center.x - a.x - b.x, s.center.y - a.y - b.y};
  c[2] = Vec2{s.center.x + a.x - b.x, s.center.y + a.y - b.y};
  c[3] = Vec2{s.center.x + a.x + b.x, s.center.y + a.y + b.y};
}

// Do any two sides of the two quadrilaterals cross?  (segmentation_util.cpp:381-410: float differences, double cross
// products, a float reciprocal.)
inline bool boxes_intersect(const Vec2 p[4], const Vec2 q[4]) {
  for (int i = 0; i < 4; ++i) {
    const Vec2& p0 = p[i];
    const Vec2& p1 = p[(i + 1) & 3];
    const double ux = (float)(p1.x - p0.x), uy = (float)(p1.y - p0.y);
    for (int j = 0; j < 4; ++j) {
      const Vec2& q0 = q[j];
      const Vec2& q1 = q[(j + 1) & 3];
      const double vx = (float)(q1.x - q0.x), vy = (float)(q1.y - q0.y);
      const double wx = (float)(q0.x - p0.x), wy = (float)(q0.y - p0.y);
      const double cross = ux * vy - uy * vx;
      if (std::fabs(cross) < 1e-6) continue;                 // parallel sides
      const float inv = 1.0f / cross;
      const double on_p = (wx * vy - wy * vx) * inv, on_q = (wx * uy - wy * ux) * inv;
      if (on_p > -1e-6f && on_p < 1.0f + 1e-6f && on_q > -1e-6f && on_q < 1.0f + 1e-6f) return true;
    }
  }
  return false;
}
}  // namespace vsbs
