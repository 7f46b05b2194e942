// TEST INFRASTRUCTURE ONLY.  The product's host half of EnforceSpatialConnectedness (video_segment_b200/csrc/tubes.hpp:
// TubeSplitter, on the moment accumulator of csrc/shape_math.hpp) against the oracle's restatement of the reference
// (oracle/vso_graph.cpp: DenseGraph::EnforceSpatialConnectedness, dense_segmentation_graph.h:666-904) on random label
// volumes: blobs that move, break apart, shed specks and rejoin, with and without a flow field.  The device half (runs,
// N4 components, moments: csrc/shape.cu) is stood in for by a few lines of CPU code here; what is compared is the
// partition into regions after the split and the order of the regions (index of every voxel).  CPU test.
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <numeric>
#include <random>
#include <string>
#include <unordered_map>
#include <vector>

#include "../video_segment_b200/csrc/tubes.hpp"

extern "C" int vso_test_spatial_connectedness(const int32_t* labels, int w, int h, int frames, const float* flows,
                                              int32_t* region_index_out);

namespace {

struct Run { int frame, y, lx, rx, label; };

// what csrc/shape.cu hands to the host: runs in raster order, their N4 components (named by first run), per component
// the intervals in raster order and the moments
struct Components {
  std::vector<Run> runs;
  std::vector<int> comp_of_run;                       // component = index of its first run
  std::vector<std::vector<vsbs::Interval>> intervals; // per component (indexed by first run; empty for non-first runs)
};

Components components_of(const std::vector<int32_t>& labels, int w, int h, int frames) {
  Components c;
  std::vector<int> row_start;
  for (int t = 0; t < frames; ++t)
    for (int y = 0; y < h; ++y) {
      row_start.push_back((int)c.runs.size());
      const int32_t* row = &labels[((size_t)t * h + y) * w];
      int x = 0;
      while (x < w) {
        int e = x;
        while (e + 1 < w && row[e + 1] == row[x]) ++e;
        c.runs.push_back(Run{t, y, x, e, row[x]});
        x = e + 1;
      }
    }
  row_start.push_back((int)c.runs.size());
  const int n = (int)c.runs.size();
  std::vector<int> par(n);
  std::iota(par.begin(), par.end(), 0);
  auto find = [&](int x) { while (par[x] != x) { par[x] = par[par[x]]; x = par[x]; } return x; };
  for (int r = 0; r < frames * h; ++r) {
    if (r % h == 0) continue;
    for (int i = row_start[r]; i < row_start[r + 1]; ++i)
      for (int k = row_start[r - 1]; k < row_start[r]; ++k)
        if (c.runs[k].label == c.runs[i].label && c.runs[k].lx <= c.runs[i].rx && c.runs[k].rx >= c.runs[i].lx) {
          const int a = find(i), b = find(k);
          if (a != b) par[std::max(a, b)] = std::min(a, b);
        }
  }
  c.comp_of_run.resize(n);
  c.intervals.resize(n);
  for (int i = 0; i < n; ++i) {
    c.comp_of_run[i] = find(i);
    c.intervals[c.comp_of_run[i]].push_back(vsbs::Interval{c.runs[i].y, c.runs[i].lx, c.runs[i].rx});
  }
  return c;
}

}  // namespace

// Returns the number of volumes on which the product's split differs from the oracle's; msg = first difference.
extern "C" int host_tubes_check(unsigned seed, int n_cases, int with_flow, char* msg, int msg_cap) {
  std::mt19937 rng(seed);
  std::string first;
  int bad = 0, splits = 0;
  for (int cs = 0; cs < n_cases; ++cs) {
    const int w = 40 + rng() % 50, h = 30 + rng() % 40, frames = 3 + rng() % 6, n_labels = 2 + rng() % 4;
    std::vector<int32_t> labels((size_t)w * h * frames, 0);
    std::vector<float> flows;
    if (with_flow) flows.assign((size_t)w * h * frames * 2, 0.f);
    // moving blobs: several per label, so that a label is in several pieces in some frames and one piece in others
    struct Blob { float x, y, vx, vy; int rx, ry, label, from, to; };
    std::vector<Blob> blobs;
    const int n_blobs = 3 + rng() % 8;
    for (int b = 0; b < n_blobs; ++b) {
      Blob bl;
      bl.x = (float)(rng() % w); bl.y = (float)(rng() % h);
      bl.vx = ((int)(rng() % 13) - 6) * 0.8f; bl.vy = ((int)(rng() % 13) - 6) * 0.8f;
      bl.rx = 1 + rng() % 9; bl.ry = 1 + rng() % 9;
      bl.label = 1 + rng() % n_labels;
      bl.from = rng() % frames; bl.to = bl.from + rng() % frames;
      blobs.push_back(bl);
    }
    for (int t = 0; t < frames; ++t) {
      int32_t* img = &labels[(size_t)t * w * h];
      for (const Blob& bl : blobs) {
        if (t < bl.from || t > bl.to) continue;
        const float cx = bl.x + bl.vx * t, cy = bl.y + bl.vy * t;
        for (int y = std::max(0, (int)cy - bl.ry); y <= std::min(h - 1, (int)cy + bl.ry); ++y)
          for (int x = std::max(0, (int)cx - bl.rx); x <= std::min(w - 1, (int)cx + bl.rx); ++x) {
            const float dx = (x - cx) / bl.rx, dy = (y - cy) / bl.ry;
            if (dx * dx + dy * dy <= 1.0f) {
              img[(size_t)y * w + x] = bl.label;
              if (with_flow && t > 0) { flows[(((size_t)t * h + y) * w + x) * 2] = -bl.vx; flows[(((size_t)t * h + y) * w + x) * 2 + 1] = -bl.vy; }
            }
          }
      }
      for (int k = 0; k < (int)(w * h / 200); ++k) img[rng() % (w * h)] = 1 + rng() % n_labels;     // specks
    }
    // ---- oracle ----
    std::vector<int32_t> want((size_t)w * h * frames);
    const int n_want = vso_test_spatial_connectedness(labels.data(), w, h, frames, with_flow ? flows.data() : nullptr, want.data());
    // ---- product: components -> pieces per label (first-seen order) -> TubeSplitter -> largest tube keeps the region ----
    const Components cc = components_of(labels, w, h, frames);
    std::vector<int> label_order;
    std::unordered_map<int, int> region_of_label;
    std::vector<std::vector<vsbt::Piece>> pieces;
    std::vector<int> comp_ids;
    for (int i = 0; i < (int)cc.runs.size(); ++i) if (cc.comp_of_run[i] == i) comp_ids.push_back(i);   // ascending = order of first interval
    for (int c : comp_ids) {
      const Run& r0 = cc.runs[c];
      auto it = region_of_label.find(r0.label);
      if (it == region_of_label.end()) { it = region_of_label.emplace(r0.label, (int)pieces.size()).first; pieces.emplace_back(); }
      vsbs::MomentSum sum;
      for (const auto& iv : cc.intervals[c]) sum.add(iv.y, iv.lx, iv.rx);
      vsbt::Piece p;
      p.frame = r0.frame; p.group = c; p.moments = sum.mean();
      p.intervals = cc.intervals[c].data(); p.n_intervals = (int)cc.intervals[c].size();
      pieces[it->second].push_back(p);
    }
    std::vector<const float*> fl;
    if (with_flow) for (int t = 0; t < frames; ++t) fl.push_back(t == 0 ? nullptr : flows.data() + (size_t)t * w * h * 2);
    std::vector<int> region_of_comp(cc.runs.size(), -1);
    int n_regions = (int)pieces.size();
    for (int r = 0; r < (int)pieces.size(); ++r) {
      for (const auto& p : pieces[r]) region_of_comp[p.group] = r;
      if (pieces[r].size() < 2) continue;
      const std::vector<vsbt::Tube> tubes = vsbt::TubeSplitter(pieces[r], w, h, with_flow ? &fl : nullptr).run();
      if (tubes.empty()) continue;
      ++splits;
      int keep = -1, keep_score = 0;
      std::vector<float> areas(tubes.size());
      for (int k = 0; k < (int)tubes.size(); ++k) {
        float area = 0;
        for (const auto& s : tubes[k]) area += s.shape.size;
        areas[k] = area;
        if (area > keep_score) { keep_score = area; keep = k; }
      }
      for (int k = 0; k < (int)tubes.size(); ++k) {
        if (k == keep) continue;
        const int fresh = n_regions++;
        for (const auto& s : tubes[k]) for (int pi : s.pieces) region_of_comp[pieces[r][pi].group] = fresh;
      }
    }
    std::vector<int32_t> got((size_t)w * h * frames, -1);
    for (int i = 0; i < (int)cc.runs.size(); ++i) {
      const Run& r = cc.runs[i];
      for (int x = r.lx; x <= r.rx; ++x) got[((size_t)r.frame * h + r.y) * w + x] = region_of_comp[cc.comp_of_run[i]];
    }
    if (n_regions != n_want || got != want) {
      if (first.empty()) {
        size_t at = 0;
        while (at < got.size() && got[at] == want[at]) ++at;
        first = "case " + std::to_string(cs) + " (" + std::to_string(w) + "x" + std::to_string(h) + "x" + std::to_string(frames) + "): regions " +
                std::to_string(n_regions) + " vs " + std::to_string(n_want) + ", first differing voxel " + std::to_string(at);
      }
      ++bad;
    }
  }
  if (msg && msg_cap > 0) {
    const std::string out = first.empty() ? ("ok, " + std::to_string(splits) + " regions split") : first;
    strncpy(msg, out.c_str(), msg_cap - 1);
    msg[msg_cap - 1] = 0;
  }
  return bad;
}
