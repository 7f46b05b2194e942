// vso_shape.cpp -- CPU ORACLE (test infrastructure): rasterization / shape helpers.
// Restates the subset of segment_util/segmentation_util.cpp used by the dense path.
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <unordered_map>

#include "vso_core.hpp"

namespace vso {

// segmentation/segmentation_common.h:144-152
template <class T> bool InsertSortedUniquely(const T& t, std::vector<T>* array) {
  auto pos = std::lower_bound(array->begin(), array->end(), t);
  if (pos == array->end() || *pos != t) {
    array->insert(pos, t);
    return true;
  }
  return false;
}
template bool InsertSortedUniquely<int>(const int&, std::vector<int>*);

// segment_util/segmentation_util.cpp:484-570
void MergeRasterization(const Rasterization& lhs, const Rasterization& rhs, Rasterization* out) {
  auto l = lhs.begin(), r = rhs.begin();
  const auto le = lhs.end(), re = rhs.end();
  std::vector<int> offs;
  Rasterization merged;
  while (l != le || r != re) {
    const int ly = (l == le ? 1 << 30 : l->y);
    const int ry = (r == re ? 1 << 30 : r->y);
    if (ly < ry) {
      merged.push_back(*l++);
    } else if (ry < ly) {
      merged.push_back(*r++);
    } else {
      offs.clear();
      bool lc, rc;
      while ((lc = (l != le && l->y == ly)) | (rc = (r != re && r->y == ry))) {
        const int lx = lc ? l->left_x : std::numeric_limits<int>::max();
        const int rx = rc ? r->left_x : std::numeric_limits<int>::max();
        if (lx < rx) {
          offs.push_back(l->left_x); offs.push_back(l->right_x); ++l;
        } else {
          offs.push_back(r->left_x); offs.push_back(r->right_x); ++r;
        }
      }
      int k = 0, ll = 0;
      const int sz_k = (int)offs.size();
      while (k < sz_k) {
        if (k + 2 == sz_k) {
          merged.push_back({ly, offs[ll], offs[k + 1]});
          break;
        } else if (offs[k + 2] - 1 == offs[k + 1]) {
          k += 2;
        } else {
          merged.push_back({ly, offs[ll], offs[k + 1]});
          k += 2;
          ll = k;
        }
      }
    }
  }
  out->swap(merged);
}

// segment_util/segmentation_util.cpp:644-650
int RasterizationArea(const Rasterization& r) {
  int area = 0;
  for (const auto& s : r) area += s.right_x - s.left_x + 1;
  return area;
}

// segment_util/segmentation_util.cpp:652-693
void ShapeMomentsFromRasterization(const Rasterization& raster, ShapeMoments* moments) {
  float mean_x = 0, mean_y = 0, moment_xx = 0, moment_yy = 0, moment_xy = 0, area_sum = 0;
  for (const auto& s : raster) {
    const float m = s.left_x;
    const float n = s.right_x;
    const float curr_y = s.y;
    const float len = (n - m + 1);
    area_sum += len;
    const float center_x = (n + m) * 0.5;
    const float sum_x = center_x * len;
    const float sum_y = curr_y * len;
    mean_x += sum_x;
    mean_y += sum_y;
    moment_xy += curr_y * sum_x;
    moment_yy += curr_y * sum_y;
    moment_xx += len * (-m + 2 * m * m + n + 2 * m * n + 2 * n * n) / 6.0f;
  }
  const float inv_area = 1.0f / area_sum;
  moments->size = area_sum;
  moments->mean_x = mean_x * inv_area;
  moments->mean_y = mean_y * inv_area;
  moments->moment_xx = moment_xx * inv_area;
  moments->moment_xy = moment_xy * inv_area;
  moments->moment_yy = moment_yy * inv_area;
}

// segment_util/segmentation_util.cpp:243-340 (single-moment case :342-345)
bool GetShapeDescriptorFromShapeMoment(const ShapeMoments& moment, ShapeDescriptor* sd) {
  float mixed_x = 0, mixed_y = 0, mixed_xx = 0, mixed_xy = 0, mixed_yy = 0, area_sum = 0;
  const float area = moment.size;
  area_sum += area;
  mixed_x += moment.mean_x * area;
  mixed_y += moment.mean_y * area;
  mixed_xx += moment.moment_xx * area;
  mixed_xy += moment.moment_xy * area;
  mixed_yy += moment.moment_yy * area;
  const float inv_area_sum = 1.0f / area_sum;
  mixed_x *= inv_area_sum; mixed_y *= inv_area_sum;
  mixed_xx *= inv_area_sum; mixed_xy *= inv_area_sum; mixed_yy *= inv_area_sum;
  sd->center = Point2f{mixed_x, mixed_y};
  sd->size = area_sum;
  if (area_sum < 10) return false;
  const float var_xx = mixed_xx - mixed_x * mixed_x;
  const float var_xy = mixed_xy - mixed_x * mixed_y;
  const float var_yy = mixed_yy - mixed_y * mixed_y;
  const float trace = var_xx + var_yy;
  const float det = var_xx * var_yy - var_xy * var_xy;
  float discriminant = 0.25 * trace * trace - det;
  discriminant = std::max(0.0f, discriminant);
  const float sqrt_disc = std::sqrt(discriminant);
  const float e_1 = trace * 0.5 - sqrt_disc;
  const float e_2 = trace * 0.5 + sqrt_disc;
  if (std::min(std::fabs(e_1), std::fabs(e_2)) < 1) return false;
  Point2f ev_1{1.0f, 0.0f}, ev_2{0.0f, 1.0f};
  const Point2f v_1{e_1 - var_yy, var_xy};
  const Point2f v_2{e_2 - var_yy, var_xy};
  const float v_1_norm = std::hypot(v_1.y, v_1.x);
  const float v_2_norm = std::hypot(v_2.y, v_2.x);
  if (v_1_norm > 1e-6f && v_2_norm > 1e-6f && discriminant > 0.1) {
    const float s1 = 1.0f / v_1_norm, s2 = 1.0f / v_2_norm;
    ev_1 = Point2f{v_1.x * s1, v_1.y * s1};
    ev_2 = Point2f{v_2.x * s2, v_2.y * s2};
  }
  float e_1_sigma = std::sqrt(std::fabs(e_1));
  float e_2_sigma = std::sqrt(std::fabs(e_2));
  if (e_1_sigma < e_2_sigma) {
    std::swap(e_1_sigma, e_2_sigma);
    std::swap(ev_1, ev_2);
  }
  const Point2f ev_1_normal{-ev_1.y, ev_1.x};
  if (ev_2.x * ev_1_normal.x + ev_2.y * ev_1_normal.y < 0) {
    ev_2 = Point2f{-ev_2.x, -ev_2.y};
  }
  sd->center = Point2f{mixed_x, mixed_y};
  sd->mag_major = e_1_sigma;
  sd->mag_minor = e_2_sigma;
  sd->dir_major = ev_1;
  sd->dir_minor = ev_2;
  return true;
}

namespace {
// segment_util/segmentation_util.cpp:1009-1023 (N4 case)
inline bool ScanIntervalsNeighboredN4(const ScanInterval& a, const ScanInterval& b) {
  return std::abs(a.y - b.y) <= 1 &&
         std::max(a.left_x, b.left_x) <= std::min(a.right_x, b.right_x);
}
struct DisjointSets {   // stands in for boost::disjoint_sets (any union-find gives the same partition)
  std::vector<int> parent;
  explicit DisjointSets(int n) : parent(n) { std::iota(parent.begin(), parent.end(), 0); }
  int find(int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; }
  void unite(int a, int b) { a = find(a); b = find(b); if (a != b) parent[std::max(a, b)] = std::min(a, b); }
};
}  // namespace

// segment_util/segmentation_util.cpp:1025-1101 (connect == N4_CONNECT).  Components are
// emitted in order of first scan interval, intervals within a component in raster order.
int ConnectedComponentsN4(const Rasterization& raster, std::vector<Rasterization>* components) {
  const int n = (int)raster.size();
  DisjointSets classes(n);
  int last_change_idx = -1, last_y = -2, test_idx = 0;
  for (int i = 0; i < n; ++i) {
    const ScanInterval& cur = raster[i];
    if (cur.y != last_y) {
      test_idx = (last_y + 1 == cur.y) ? last_change_idx : i;
      last_y = cur.y;
      last_change_idx = i;
    }
    for (int k = test_idx; k < i; ++k) {
      if (ScanIntervalsNeighboredN4(cur, raster[k])) classes.unite(i, k);
    }
  }
  int num_components = 0;
  for (int i = 0; i < n; ++i) num_components += (classes.find(i) == i);
  if (num_components == 1) {
    if (components) components->push_back(raster);
    return 1;
  }
  if (components) {
    components->reserve(num_components);
    std::unordered_map<int, int> rep_to_comp;
    for (int i = 0; i < n; ++i) {
      const int rep = classes.find(i);
      auto it = rep_to_comp.find(rep);
      if (it == rep_to_comp.end()) {
        rep_to_comp[rep] = (int)components->size();
        components->push_back(Rasterization{raster[i]});
      } else {
        (*components)[it->second].push_back(raster[i]);
      }
    }
  }
  return num_components;
}

// segment_util/segmentation_util.cpp:741-770 (level 0): later regions overwrite earlier.
void SegDescToIdImage(const SegDesc& seg, int width, int* id_image) {
  for (const auto& region : seg.region) {
    for (const auto& s : region.raster) {
      int* out = id_image + (size_t)s.y * width + s.left_x;
      for (int j = 0, len = s.right_x - s.left_x + 1; j < len; ++j) out[j] = region.id;
    }
  }
}

}  // namespace vso
