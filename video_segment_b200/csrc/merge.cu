// merge.cu -- bucket-ordered union-find merge of one chunk graph, sm_100a.
// Replaces FastSegmentationGraph::SegmentGraph (segmentation/segmentation_graph.h:339-463) with
// GetRegion / MergeRegions (:651-701) and ColorMeanDescriptorTraits (segmentation/pixel_distance.h:
// 469-521).  The reference scan is strictly sequential and order dependent.  This kernel keeps
// the reference's order semantics and extracts parallelism in three provably order-preserving ways:
//
//  (1) deterministic reservations: inside a bucket every pending edge reserves both of its
//      current roots with its position in the reference order (atomicMin); an edge that holds
//      both reservations is the next edge the serial scan would apply to those two regions, so
//      it is applied with the exact serial decision tree (same float formulas, no FMA).
//  (2) permanently inert edges (same root, or a finalised pair where both sides reached the
//      minimum region size) are dropped without ordering.
//  (3) "safe clusters": the connected components of a bucket's pending edges whose roots are
//      all un-finalised, constraint compatible and whose mean colours span less than the merge
//      threshold merge completely whatever the order (every partial mean stays inside the
//      hull), so they are merged by a lock-free union-find in one step; and small regions
//      hanging off a finalised region of >= min size are absorbed as soon as the edge is the
//      small region's next edge (the big side's class cannot change any more).
//  Bulk merges (3) compute the size-weighted mean from 64-bit fixed-point sums instead of the
//  reference's running float mean (differs by rounding only).
//
// One persistent cooperative launch walks all 2048 buckets; buckets with few pending edges
// are finished by block 0 alone behind __syncthreads instead of grid-wide barriers.
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace vsb {

constexpr uint32_t kDone = 0xFFFFFFFFu;
constexpr unsigned kTailEdges = 4096;     // <= this many edges: block 0 runs the whole bucket
constexpr unsigned long long kSerialSwitch = 256;   // grid round progress below this -> serial window mode
constexpr int kMergeThreads = 512;
constexpr int kMergeWarps = kMergeThreads / 32;
constexpr double kFix = 4294967296.0;     // 2^32 fixed point for descriptor sums

__device__ __forceinline__ void trace(const MergeParams& p, int slot, unsigned long long v) {
  if (p.trace) { ((volatile unsigned long long*)p.trace)[slot] = v; __threadfence_system(); }
}

struct GridBar {
  cg::grid_group g;
  __device__ void sync() { g.sync(); }
};
struct BlockBar {
  __device__ void sync() { __syncthreads(); }
};

__device__ __forceinline__ RegionRec load_rec(const RegionRec* r) {
  const int4 a = reinterpret_cast<const int4*>(r)[0];
  const int4 b = reinterpret_cast<const int4*>(r)[1];
  RegionRec o;
  o.sz = a.x; o.con = a.y; o.d0 = __int_as_float(a.z); o.d1 = __int_as_float(a.w);
  o.d2 = __int_as_float(b.x); o.fin = b.y; o.pad0 = b.z; o.pad1 = b.w;
  return o;
}
__device__ __forceinline__ void store_rec(RegionRec* r, const RegionRec& o) {
  reinterpret_cast<int4*>(r)[0] = make_int4(o.sz, o.con, __float_as_int(o.d0), __float_as_int(o.d1));
  reinterpret_cast<int4*>(r)[1] = make_int4(__float_as_int(o.d2), o.fin, 0, 0);
}

// GetRegion (segmentation_graph.h:651-669): find with path halving.
__device__ __forceinline__ int uf_find(int* parent, int x) {
  int p = parent[x];
  while (p != x) {
    const int gp = parent[p];
    if (gp != p) parent[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ int cl_find(const int* cl, int x) {
  int p = cl[x];
  while (p != x) { x = p; p = cl[x]; }
  return x;
}
// find with path halving; only used once all unions of the window are done (plain stores towards the root are benign)
__device__ __forceinline__ int cl_find_compress(int* cl, int x) {
  int p = cl[x];
  while (p != x) {
    const int gp = cl[p];
    if (gp != p) cl[x] = gp;
    x = p;
    p = gp;
  }
  return x;
}
__device__ __forceinline__ void cl_union(int* cl, int a, int b) {
  while (true) {
    a = cl_find(cl, a);
    b = cl_find(cl, b);
    if (a == b) return;
    const int hi = max(a, b), lo = min(a, b);
    if (atomicCAS(&cl[hi], hi, lo) == hi) return;
  }
}

// edge code -> node ids (region_1 = anchor in the current slot, region_2 = neighbour).
__device__ __forceinline__ void decode_edge(const MergeParams& p, uint32_t code, int& u, int& v) {
  const int n = p.w * p.h;
  int list;
  uint32_t upix, udir;
  edge_decode(code, (uint32_t)n, list, upix, udir);
  const int pix = (int)upix, dir = (int)udir;
  const int slot = (list + 1) >> 1;
  u = slot * n + pix;
  if ((list & 1) == 0) {   // spatial: R, B, BL, BR
    const int off = (dir == 0) ? 1 : (dir == 1) ? p.w : (dir == 2) ? p.w - 1 : p.w + 1;
    v = u + off;
  } else {                 // temporal: 3x3 about the (flow displaced, clamped) centre in slot-1
    int px = pix % p.w, py = pix / p.w;
    if (p.flows) {
      const float* f = p.flows + ((size_t)slot * n + pix) * 2;
      px = max(0, min(p.w - 1, (int)((float)px + f[0])));
      py = max(0, min(p.h - 1, (int)((float)py + f[1])));
    }
    const int dy = dir / 3 - 1, dx = dir % 3 - 1;
    v = (slot - 1) * n + (py + dy) * p.w + (px + dx);
  }
}

// ColorMeanDescriptorTraits::DescriptorDistance (pixel_distance.h:478-491)
__device__ __forceinline__ float raw_dist(const RegionRec& a, const RegionRec& b) {
  const float d1 = a.d0 - b.d0, d2 = a.d1 - b.d1, d3 = a.d2 - b.d2;
  return sqrtf((d1 * d1 + d2 * d2 + d3 * d3) * (1.0f / 3.0f));
}
__device__ __forceinline__ float desc_dist(const RegionRec& a, const RegionRec& b, float edge_w, float force_w) {
  const float dist = raw_dist(a, b);
  if (edge_w < force_w && (double)dist < 0.2) return 0.0f;
  return dist;
}

// MergeRegions (segmentation_graph.h:671-701) + MergeDescriptor (pixel_distance.h:494-504).
// A = rep_1 (id ia), B = rep_2 (id ib).  Returns the id of the surviving representative.
__device__ __forceinline__ int merge_regions(int* parent, int ia, RegionRec& A, int ib, RegionRec& B) {
  const bool a_wins = A.sz > B.sz;
  RegionRec& m = a_wins ? A : B;
  RegionRec& o = a_wins ? B : A;
  const float denom = 1.0f / (float)(o.sz + m.sz);
  const float fa = (float)o.sz * denom;
  const float fb = (float)m.sz * denom;
  m.d0 = fa * o.d0 + fb * m.d0;
  m.d1 = fa * o.d1 + fb * m.d1;
  m.d2 = fa * o.d2 + fb * m.d2;
  m.sz += o.sz;
  m.con = max(A.con, B.con);
  parent[a_wins ? ib : ia] = a_wins ? ia : ib;
  return a_wins ? ia : ib;
}

// The serial decision tree for one edge (segmentation_graph.h:375-440) on register copies of the
// two representatives' records (A = rep_1, B = rep_2).  Returns 1 if A survives a merge, 2 if B
// survives a merge, 0 if the regions stay separate (records may still have changed).
__device__ __forceinline__ int decide_pair(const MergeParams& p, RegionRec& A, RegionRec& B, float edge_w) {
  const int mins = p.min_region_size;
  auto merge = [&]() -> int {                 // MergeRegions (:671-701) + MergeDescriptor (pixel_distance.h:494-504)
    const bool a_wins = A.sz > B.sz;
    RegionRec& m = a_wins ? A : B;
    RegionRec& o = a_wins ? B : A;
    const float denom = __frcp_rn((float)(o.sz + m.sz));     // == 1.0f / x, correctly rounded
    const float fa = (float)o.sz * denom;
    const float fb = (float)m.sz * denom;
    m.d0 = fa * o.d0 + fb * m.d0;
    m.d1 = fa * o.d1 + fb * m.d1;
    m.d2 = fa * o.d2 + fb * m.d2;
    m.sz += o.sz;
    m.con = max(A.con, B.con);
    return a_wins ? 1 : 2;
  };
  // The gates are taken on the squared distance y (dist = sqrtf(y), monotone): DescriptorDistance returns 0 in the
  // force-merge buckets while dist < 0.2 (pixel_distance.h:478-491), so "d < 0.05" reads y < y_force there and
  // y < y_merge elsewhere; "d > 0.15" reads y >= y_split unless the force rule zeroes the distance.
  const bool force = edge_w < p.force_merge_weight;
  if (A.con < 0 || B.con < 0) {
    if (!A.fin && !B.fin) {
      const float d1 = A.d0 - B.d0, d2 = A.d1 - B.d1, d3 = A.d2 - B.d2;
      const float y = (d1 * d1 + d2 * d2 + d3 * d3) * (1.0f / 3.0f);
      if (y < (force ? p.y_force : p.y_merge)) return merge();          // MergeDistanceThreshold, pixel_distance.h:471
      A.fin = 1;
      B.fin = 1;
    }
    if (A.fin || B.fin) {
      if (A.sz < mins || B.sz < mins) return merge();
    }
    return 0;
  } else if (A.con == B.con) {
    const float d1 = A.d0 - B.d0, d2 = A.d1 - B.d1, d3 = A.d2 - B.d2;
    const float y = (d1 * d1 + d2 * d2 + d3 * d3) * (1.0f / 3.0f);
    if (!(force && y < p.y_force) && y >= p.y_split) {   // SplitDistanceThreshold, pixel_distance.h:472
      if ((double)A.sz < (double)B.sz * 0.3) A.con = -1;
      else if ((double)B.sz < (double)A.sz * 0.3) B.con = -1;
      else { A.con = -1; B.con = -1; }
      return 0;
    }
    return merge();
  }
  return 0;                                   // different constraint ids: never merge
}

// One edge whose two roots this thread owns.  Returns the surviving representative of a merge,
// or -1 if the regions stay separate.
__device__ __forceinline__ int exec_strict(const MergeParams& p, int ia, int ib, float edge_w,
                                           unsigned long long* stats) {
  RegionRec A = load_rec(&p.rec[ia]);
  RegionRec B = load_rec(&p.rec[ib]);
  const int r = decide_pair(p, A, B, edge_w);
  if (r == 1) { p.parent[ib] = ia; store_rec(&p.rec[ia], A); return ia; }
  if (r == 2) { p.parent[ia] = ib; store_rec(&p.rec[ib], B); return ib; }
  store_rec(&p.rec[ia], A);
  store_rec(&p.rec[ib], B);
  return -1;
}

__device__ __forceinline__ void acc_add(unsigned long long* acc, int root, const RegionRec& r) {
  unsigned long long* a = acc + (size_t)root * 4;
  const unsigned long long sz = (unsigned long long)r.sz;
  atomicAdd(&a[1], (unsigned long long)__double2ll_rn((double)r.d0 * kFix) * sz);
  atomicAdd(&a[2], (unsigned long long)__double2ll_rn((double)r.d1 * kFix) * sz);
  atomicAdd(&a[3], (unsigned long long)__double2ll_rn((double)r.d2 * kFix) * sz);
  atomicAdd(&a[0], sz);                // folds only run behind a barrier: no ordering needed between the four words
}

// fold pending bulk contributions into the representative's record (size-weighted mean)
__device__ __forceinline__ void acc_fold(const MergeParams& p, int root) {
  unsigned long long* a = p.acc + (size_t)root * 4;
  if (a[0] == 0ull) return;
  // all contributions were added before the preceding barrier; the size word is the gate, so
  // exactly one thread folds (zero-size contributions -- virtual nodes -- carry no colour)
  const unsigned long long sz = atomicExch(&a[0], 0ull);
  if (sz == 0ull) return;
  const unsigned long long s1 = atomicExch(&a[1], 0ull), s2 = atomicExch(&a[2], 0ull), s3 = atomicExch(&a[3], 0ull);
  RegionRec R = load_rec(&p.rec[root]);
  const double tot = (double)R.sz + (double)sz;
  if (tot > 0) {
    const double inv = 1.0 / (tot * kFix);
    R.d0 = (float)(((double)R.sz * (double)R.d0 * kFix + (double)s1) * inv);
    R.d1 = (float)(((double)R.sz * (double)R.d1 * kFix + (double)s2) * inv);
    R.d2 = (float)(((double)R.sz * (double)R.d2 * kFix + (double)s3) * inv);
  }
  R.sz += (int)sz;
  store_rec(&p.rec[root], R);
}

// Shared memory of a CTA: the staging area of the exact scans plus per-warp counters of the
// block-wide compactions.
struct ScanShared;
struct MergeShared;
__device__ __forceinline__ unsigned block_rank(MergeShared& S, bool flag, unsigned* total);

// ---------------------------------------------------------------------------------------------
// Window certification.  The pending edges of a bucket are processed in consecutive position
// windows (positions = reference order).  For the edges of one window, with the exact state at
// the window start, the kernel proves for most of them what the serial scan will do, whatever
// the order inside the window:
//   * "atoms" are the current roots; an atom of >= min_region_size voxels is a HUB, the rest is
//     small.  Small atoms joined by window edges form SUB-CLUSTERS (lock-free union-find).
//   * a sub-cluster of total mass < min_region_size that touches exactly one hub is absorbed by
//     it for sure: every edge between two regions of which one is small ends in a merge
//     (segmentation_graph.h:375-440: regular merge, or small-region merge after a failed test),
//     and the hub, being the larger side, keeps its own flag.  The only open question is whether
//     an un-finalised hub fails a test on the way (it would become finalised).  The hub's mean
//     stays within  Delta = R M / (S + M)  of its value at the window start (R = max colour
//     distance of the atoms it can absorb in this window, M = their total size, S = its own
//     size), and every region it can meet has its mean within R of it, so R + Delta <
//     threshold certifies that no test fails: the hub is FROZEN (its decision-relevant state --
//     big, flag, constraint -- cannot change in this window).  Finalised hubs are frozen by
//     definition (they only absorb small regions, no test is evaluated against them).
//   * a sub-cluster whose members are all un-finalised with a colour hull smaller than the
//     threshold merges completely (every partial mean stays inside the hull), alone or together
//     with its single un-finalised frozen hub.
//   * everything else (sub-clusters between two hubs, hubs near a failing test, constraint
//     conflicts, big-big tests) is left to the ordered rounds below, where a frozen hub may absorb
//     a small region as soon as the edge is that region's next edge.
// Windows that leave too many uncertified edges are halved and retried (a shorter window means a
// smaller drift bound); a live edge between two hubs ends the window in front of it and is
// executed alone.  Bulk merges compute the size-weighted mean from 64-bit fixed-point sums.
// ---------------------------------------------------------------------------------------------
constexpr int kScFin = 1;        // a member is finalised
constexpr int kScConMulti = 2;   // members / absorbable atoms carry different constraint ids
constexpr int kScHubs3 = 4;      // sub-cluster touches more than two hubs
constexpr int kScUnc = 8;        // hub: may meet another un-finalised hub through a shared sub-cluster
constexpr int kScCertYes = 32;     // cached verdict of subcluster_certified for this segment attempt
constexpr int kScCertNo = 64;
constexpr int kScSplitCap = 128;   // constrained chunks: an uncertified same-id edge touches this hub / sub-cluster (its id may be reset in the segment)
constexpr int kScUncAny = 16;    // hub: may meet a hub with the same constraint id through a shared sub-cluster (flags do not matter)
#ifndef VSB_WINDOW_TARGET
#define VSB_WINDOW_TARGET (1ull << 18)
#endif
#ifndef VSB_RESIDUAL_SPLIT
#define VSB_RESIDUAL_SPLIT 4096
#endif
#ifndef VSB_SEGMENT_MIN
#define VSB_SEGMENT_MIN 2048
#endif
#ifndef VSB_GROUP_SCAN_MIN
#define VSB_GROUP_SCAN_MIN 1024
#endif
constexpr unsigned long long kWindowTargetDefault = VSB_WINDOW_TARGET;   // live edges aimed at per window
constexpr unsigned long long kWindowMin = 4096;            // smallest raw window
constexpr unsigned long long kSegmentMinDefault = VSB_SEGMENT_MIN;           // segments are not halved below this many live edges
constexpr unsigned long long kResidualSplitDefault = VSB_RESIDUAL_SPLIT;        // uncertified edges that trigger a halving
constexpr int kP1U = 8;                                     // edges per thread in flight in the prune pass

// one atomic per converged group of lanes instead of one per live edge
__device__ __forceinline__ unsigned long long warp_slot(unsigned long long* cnt) {
  const unsigned act = __activemask();
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(act) - 1;
  unsigned long long base = 0;
  if ((int)lane == leader) base = atomicAdd(cnt, (unsigned long long)__popc(act));
  base = __shfl_sync(act, base, leader);
  return base + __popc(act & ((1u << lane) - 1u));
}

__device__ __forceinline__ NodeScratch load_sc(const NodeScratch* s) {
  NodeScratch o;
  const int4* q = reinterpret_cast<const int4*>(s);
  const int4 a = q[0], b = q[1], c = q[2], d = q[3];
  o.mn[0] = a.x; o.mn[1] = a.y; o.mn[2] = a.z; o.mx[0] = a.w;
  o.mx[1] = b.x; o.mx[2] = b.y; o.flags = b.z; o.con = b.w;
  o.mass = c.x; o.hub0 = c.y; o.hub1 = c.z; o.rbits = c.w;
  o.num = ((unsigned long long)(unsigned)d.y << 32) | (unsigned)d.x; o.claim = d.z; o.frozen = d.w;
  return o;
}
__device__ __forceinline__ void reset_sc(NodeScratch* s) {
  int4* q = reinterpret_cast<int4*>(s);
  q[0] = make_int4(0x7f7f7f7f, 0x7f7f7f7f, 0x7f7f7f7f, 0);
  q[1] = make_int4(0, 0, 0, kNoCon);
  q[2] = make_int4(0, -1, -1, 0);
  // num (epoch-tagged big-big key), claim / frozen tags stay
}

struct WindowThr { float thr_m, con_thr; int strict_con; };

// Is the decision-relevant state of hub H certified constant in this window?
__device__ __forceinline__ bool hub_frozen_eval(const RegionRec& H, const NodeScratch& S, const WindowThr& t) {
  // a constrained hub keeps its id whatever it meets (merging with another id never happens); an
  // unconstrained hub takes the id of the first constrained region it absorbs, so two ids make it order dependent
  if (S.flags & kScUncAny) return false;
  int conset = H.con;
  if (conset < 0) {
    if (S.flags & kScConMulti) return false;
    // A hub that may take a constraint id in this segment is not frozen: the id arrives at an unknown position, and
    // from there on its edges to regions of that id are same-id edges (measured: config B chunks 6 / 8, IoU 0.995 /
    // 0.982 -> 1.000 / 0.998 when such hubs go through the exact paths; diagnostic 512 switches this back off)
    if (S.con != kNoCon) { conset = S.con; if (!(t.strict_con & 2)) return false; }
  }
  (void)conset;      // the 0.15 split test against same-id pieces is checked per sub-cluster (subcluster_certified)
  const float R = __int_as_float(S.rbits);
  const double M = (double)S.mass, Sz = (double)H.sz;
  const float delta = (float)((double)R * M / (Sz + M)) * 1.0001f + 1e-7f;   // drift of the hub mean, see above
  if (!H.fin) {
    if (S.flags & kScUnc) return false;
    if (!(R + delta < t.thr_m)) return false;
  }
  return true;
}

// Is the sub-cluster with record S (root c) certified?  *target = node everything merges into.
__device__ __forceinline__ bool subcluster_certified(const MergeParams& p, int c, const NodeScratch& S, const WindowThr& t,
                                                     int mins, int* target) {
  *target = c;
  if (S.flags & (kScConMulti | kScHubs3)) return false;
  if ((t.strict_con & 1) && S.con != kNoCon) return false;       // diagnostic 256: sub-clusters with constrained atoms are never certified
  const float dx = __int_as_float(S.mx[0]) - __int_as_float(S.mn[0]);
  const float dy = __int_as_float(S.mx[1]) - __int_as_float(S.mn[1]);
  const float dz = __int_as_float(S.mx[2]) - __int_as_float(S.mn[2]);
  const float diam = sqrtf((dx * dx + dy * dy + dz * dz) * (1.0f / 3.0f));
  if (S.con != kNoCon && !(diam < t.con_thr)) return false;
  if (S.hub0 < 0) return !(S.flags & kScFin) && diam < t.thr_m;
  if (S.hub1 >= 0) return false;
  const RegionRec H = load_rec(&p.rec[S.hub0]);
  const NodeScratch HS = load_sc(&p.hull[S.hub0]);
  if (H.con >= 0 && S.con != kNoCon && S.con != H.con) return false;      // foreign id: those edges are kept, the rest races
  if (!hub_frozen_eval(H, HS, t)) return false;
  if (S.con != kNoCon) {
    // constrained pieces meeting a hub that carries (or will have taken) the same id are merged only
    // while the colour distance stays <= 0.15 (segmentation_graph.h:417-437): bound it by the farthest
    // corner of the sub-cluster's hull plus the hub's drift
    float s2 = 0.f;
    const float hm[3] = {H.d0, H.d1, H.d2};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const float dd = fmaxf(fabsf(__int_as_float(S.mx[k]) - hm[k]), fabsf(__int_as_float(S.mn[k]) - hm[k]));
      s2 += dd * dd;
    }
    const float far = sqrtf(s2 * (1.0f / 3.0f));
    const float Rh = __int_as_float(HS.rbits);
    const float delta = (float)((double)Rh * (double)HS.mass / ((double)H.sz + (double)HS.mass)) * 1.0001f + 1e-7f;
    if (!(far + delta < t.con_thr)) return false;
  }
  *target = S.hub0;
  if (S.mass < mins) return true;
  return !H.fin && !(S.flags & kScFin) && diam < t.thr_m;
}

// debugging tap: why is the sub-cluster not certified (mirrors subcluster_certified)
__device__ int subcluster_why(const MergeParams& p, const NodeScratch& S, const WindowThr& t, int mins) {
  if (S.flags & kScConMulti) return 1;
  if (S.flags & kScHubs3) return 2;
  const float dx = __int_as_float(S.mx[0]) - __int_as_float(S.mn[0]);
  const float dy = __int_as_float(S.mx[1]) - __int_as_float(S.mn[1]);
  const float dz = __int_as_float(S.mx[2]) - __int_as_float(S.mn[2]);
  const float diam = sqrtf((dx * dx + dy * dy + dz * dz) * (1.0f / 3.0f));
  if (S.con != kNoCon && !(diam < t.con_thr)) return 1;
  if (S.hub0 < 0) return (S.flags & kScFin) ? 3 : 4;
  if (S.hub1 >= 0) return 5;
  const RegionRec H = load_rec(&p.rec[S.hub0]);
  if (H.con >= 0 && S.con != kNoCon && S.con != H.con) return 1;
  const NodeScratch HS = load_sc(&p.hull[S.hub0]);
  if (!hub_frozen_eval(H, HS, t)) {
    // finer reason: +16 same-id / >2 hubs, +17 constraint ids conflict, +18 may meet an open hub (>2 hubs), +19 bound
    const int why = (HS.flags & kScUncAny) ? 16 : (H.con < 0 && (HS.flags & kScConMulti)) ? 17 : (!H.fin && (HS.flags & kScUnc)) ? 18 : 19;
    atomicAdd(&p.debug[kNumBuckets * 4 + why], 1ull);
    return 6;
  }
  return 7;
}

constexpr unsigned long long kBlockRoundsLimit = 2048;   // residual lists up to this size (= kScanMax) are finished by block 0 alone

// ---------------------------------------------------------------------------------------------
// Exact scan.  A small residual (<= kScanMax pending edges of one segment, in reference order) is
// finished by one CTA the way the reference does it: all threads stage the edges' current roots and
// their records in shared memory (compact local ids through a shared hash), ONE thread then walks the
// edges in order with a local union-find and the exact decision tree, and all threads write the
// result back.  Cost ~ 0.1 us per edge, independent of the dependency depth (an ordered round costs
// three grid barriers plus several dependent global loads, and a chain needs one round per link).
// ---------------------------------------------------------------------------------------------
constexpr int kScanMax = 2048;
constexpr unsigned long long kGroupScanMin = VSB_GROUP_SCAN_MIN;   // residuals above this are scanned group-parallel by all CTAs
constexpr int kScanHashBits = 13;                // 4 * kScanMax slots
static_assert((1 << kScanHashBits) == 4 * kScanMax, "hash size");
struct ScanShared;
__device__ __forceinline__ RegionRec scan_load(const int2 (*rec)[3], int j) {
  const int2 a = rec[j][0], b = rec[j][1], c = rec[j][2];
  RegionRec o;
  o.sz = a.x; o.con = a.y; o.d0 = __int_as_float(b.x); o.d1 = __int_as_float(b.y); o.d2 = __int_as_float(c.x); o.fin = c.y; o.pad0 = o.pad1 = 0;
  return o;
}
__device__ __forceinline__ void scan_store(int2 (*rec)[3], int j, const RegionRec& o) {
  rec[j][0] = make_int2(o.sz, o.con);
  rec[j][1] = make_int2(__float_as_int(o.d0), __float_as_int(o.d1));
  rec[j][2] = make_int2(__float_as_int(o.d2), o.fin);
}
struct ScanShared {
  unsigned short ea[kScanMax], eb[kScanMax];       // local ids of an edge's two roots
  uint32_t epos[kScanMax];
  int tru[kScanMax], trv[kScanMax];                 // global roots of an edge (staging)
  int gid[2 * kScanMax];                            // local id -> global root
  unsigned short par[2 * kScanMax];                 // local union-find
  int2 rec[2 * kScanMax][3];                        // 24-byte records (sz, con | d0, d1 | d2, fin): three 64-bit shared accesses each
  unsigned hkey[4 * kScanMax];                      // hash: global root + 1 (0 = empty)
  unsigned short hidx[4 * kScanMax];                // local id of the slot's root
  int n_snap, spec_fail;                            // split scan: speculated conditional meetings (snapshots used, a prediction failed)
  int n_roots;
  int hbits;                                        // hash size of the current scan: 4 x its edge count, rounded up to a power of two
};
struct MergeShared {
  ScanShared scan;
  unsigned warp_cnt[kMergeWarps];
  unsigned warp_live[kMergeWarps];
};

// block-wide exclusive scan of a 0/1 flag; returns the rank of the calling thread, total via *total
__device__ __forceinline__ unsigned block_rank(MergeShared& S, bool flag, unsigned* total) {
  const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const unsigned m = __ballot_sync(0xffffffffu, flag);
  if (lane == 0) S.warp_cnt[wid] = __popc(m);
  __syncthreads();
  unsigned base = 0, tot = 0;
#pragma unroll
  for (int k = 0; k < kMergeWarps; ++k) { const unsigned c = S.warp_cnt[k]; if (k < (int)wid) base += c; tot += c; }
  __syncthreads();
  *total = tot;
  return base + __popc(m & ((1u << lane) - 1u));
}

// phase 1 (insert) and phase 2 (lookup, after a block barrier) of the root -> local id hash
__device__ __forceinline__ void scan_insert(ScanShared& C, unsigned root) {
  unsigned slot = (root * 2654435761u) >> (32 - C.hbits);
  while (true) {
    const unsigned k = atomicCAS(&C.hkey[slot], 0u, root + 1u);
    if (k == 0u) {
      const int id = atomicAdd(&C.n_roots, 1);
      C.gid[id] = (int)root;
      C.hidx[slot] = (unsigned short)id;
      return;
    }
    if (k == root + 1u) return;
    slot = (slot + 1u) & ((1u << C.hbits) - 1u);
  }
}
__device__ __forceinline__ unsigned short scan_lookup(const ScanShared& C, unsigned root) {
  unsigned slot = (root * 2654435761u) >> (32 - C.hbits);
  while (C.hkey[slot] != root + 1u) slot = (slot + 1u) & ((1u << C.hbits) - 1u);
  return C.hidx[slot];
}

// A hub "commutes" in a segment when its decision-relevant state (big, flag, constraint id) cannot
// change there: finalised, or certified frozen for this segment (NodeScratch::frozen == wtag).
__device__ __forceinline__ bool commuting_hub(const MergeParams& p, int r, int wtag) {
  const int sz = p.rec[r].sz;
  if (sz < p.min_region_size) return false;
  return p.rec[r].fin != 0 || p.hull[r].frozen == wtag;
}

// all threads of ONE block; pend_list[0 .. n_pend) ascending positions relative to codes / done_flags.
// commute_tag >= 0: several CTAs scan disjoint groups of regions at the same time; commuting hubs are shared
// between them, so they only absorb (fixed-point accumulators, folded by the caller behind a barrier) and
// an edge that would need a hub's exact record defers its whole group (grp_of[i] = group of edge i) to the
// caller's serial tail: its edges stay pending.  commute_tag < 0: plain exact scan, every root is owned.
__device__ void exact_scan(const MergeParams& p, ScanShared& C, const int b, const uint32_t* codes, const uint32_t* pend_list,
                           const int n_pend, unsigned char* done_flags, const int commute_tag, const uint32_t* grp_of,
                           unsigned long long* deferred_flag) {
  const float edge_w = (float)b * (float)(1.0 / (double)bucket_scale());
  const int mins = p.min_region_size;
  const int tid = threadIdx.x, nthr = blockDim.x;
  long long t_scan0 = 0, t_walk0 = 0, t_walk1 = 0;
  if (p.debug && tid == 0) t_scan0 = clock64();
  unsigned char* const ishub = reinterpret_cast<unsigned char*>(C.tru);     // tru / trv are free once the ids are looked up
  unsigned char* const skip = reinterpret_cast<unsigned char*>(C.trv);      // per edge: left pending (deferred group)
  int hbits = 6;
  while ((1 << hbits) < 4 * n_pend) ++hbits;          // <= kScanHashBits since n_pend <= kScanMax
  for (int i = tid; i < (1 << hbits); i += nthr) { C.hkey[i] = 0u; C.hidx[i] = 0xFFFFu; }
  if (tid == 0) { C.n_roots = 0; C.hbits = hbits; }
  __syncthreads();
  for (int i = tid; i < n_pend; i += nthr) {
    const uint32_t pos = pend_list[i];
    int u, v;
    decode_edge(p, codes[pos], u, v);
    const int ru = uf_find(p.parent, u), rv = uf_find(p.parent, v);
    C.epos[i] = pos;
    C.tru[i] = ru; C.trv[i] = rv;
    scan_insert(C, (unsigned)ru);
    scan_insert(C, (unsigned)rv);
  }
  __syncthreads();
  for (int i = tid; i < n_pend; i += nthr) {
    C.ea[i] = scan_lookup(C, (unsigned)C.tru[i]);
    C.eb[i] = scan_lookup(C, (unsigned)C.trv[i]);
  }
  __syncthreads();
  const int n_roots = C.n_roots;
  for (int j = tid; j < n_roots; j += nthr) {
    C.par[j] = (unsigned short)j;
    scan_store(C.rec, j, load_rec(&p.rec[C.gid[j]]));
    ishub[j] = (commute_tag >= 0 && commuting_hub(p, C.gid[j], commute_tag)) ? 1 : 0;
  }
  for (int i = tid; i < n_pend; i += nthr) skip[i] = 0;
  __syncthreads();
  if (tid == 0) {
    if (p.debug) t_walk0 = clock64();
    auto findl = [&](int x) { int q = C.par[x]; while (q != x) { const int g = C.par[q]; C.par[x] = (unsigned short)g; x = q; q = g; } return x; };
    uint32_t deferred[16];
    int n_def = 0;
    for (int i = 0; i < n_pend; ++i) {
      if (n_def) {
        bool d = false;
        for (int k = 0; k < n_def; ++k) d = d || (deferred[k] == grp_of[i]);
        if (d) { skip[i] = 1; continue; }
      }
      const int a = findl(C.ea[i]), bq = findl(C.eb[i]);
      if (a == bq) continue;
      const bool ha = ishub[a] != 0, hb = ishub[bq] != 0;
      if (ha || hb) {
        // a shared hub: only outcomes that do not need (or change) its exact record are taken here
        bool conflict = false;
        if (ha && hb) {
          const RegionRec A = scan_load(C.rec, a), B = scan_load(C.rec, bq);
          const bool both_con = A.con >= 0 && B.con >= 0;
          const bool noop = (both_con && A.con != B.con) || (!both_con && (A.fin || B.fin));
          conflict = !noop;
        } else {
          const int h = ha ? a : bq, x = ha ? bq : a;
          const RegionRec H = scan_load(C.rec, h), X = scan_load(C.rec, x);
          if (X.con >= 0) {
            conflict = !(H.con >= 0 && H.con != X.con);          // different ids: the edge is kept, nothing happens
          } else if (!((H.fin || X.fin) && X.sz >= mins)) {       // not inert: the hub absorbs X (small side / certified test)
            C.par[x] = (unsigned short)h;
            acc_add(p.acc, C.gid[h], X);
          }
        }
        if (conflict) {
          skip[i] = 1;
          *deferred_flag = 1ull;
          // later scans of this CTA must skip the group as well: remember it in global memory (group id = a node id)
          p.hull[grp_of[i]].claim = -commute_tag - 1;
          if (n_def < 16) deferred[n_def++] = grp_of[i];
          else {
            // cannot track more groups locally: everything left in this batch stays pending and all of its groups are deferred
            for (int k = i + 1; k < n_pend; ++k) { skip[k] = 1; p.hull[grp_of[k]].claim = -commute_tag - 1; }
            break;
          }
        }
        continue;
      }
      RegionRec A = scan_load(C.rec, a), B = scan_load(C.rec, bq);
      const int r = decide_pair(p, A, B, edge_w);      // rep_1 = root of region_1 (the anchor), as in the reference
      if (r != 2) scan_store(C.rec, a, A);              // survivor, or flags / constraint changed without a merge
      if (r != 1) scan_store(C.rec, bq, B);
      if (r == 1) C.par[bq] = (unsigned short)a;
      else if (r == 2) C.par[a] = (unsigned short)bq;
    }
    if (p.debug) t_walk1 = clock64();
  }
  __syncthreads();
  for (int j = tid; j < n_roots; j += nthr) {
    int x = j;
    while (C.par[x] != x) x = C.par[x];
    const int g = C.gid[j];
    if (x == j) {
      if (!ishub[j]) store_rec(&p.rec[g], scan_load(C.rec, j));       // shared hubs are folded by the caller
    } else {
      p.parent[g] = C.gid[x];
    }
  }
  for (int i = tid; i < n_pend; i += nthr) if (!skip[i]) done_flags[C.epos[i]] = 1;
  if (tid == 0) atomicAdd(&p.stats[3], 1ull);
  __syncthreads();
  if (p.debug && tid == 0) {      // development tap: cycles of the one-thread walk / of the whole scan (staging + walk + write-back)
    atomicAdd(&p.debug[kNumBuckets * 4 + 50], (unsigned long long)(t_walk1 - t_walk0));
    atomicAdd(&p.debug[kNumBuckets * 4 + 51], (unsigned long long)(clock64() - t_scan0));
    atomicAdd(&p.debug[kNumBuckets * 4 + 52], (unsigned long long)n_roots);
  }
}

// ---------------------------------------------------------------------------------------------
// Split scan: the exact scan of one ordered list (<= kScanMax pending edges, every root owned by this
// CTA) without the one-thread walk over every edge (measured: 683 cycles per edge, 97 % of the scan).
// The walk is split along what actually depends on what:
//   * a region below min_region_size always merges with whatever a live edge joins it to
//     (segmentation_graph.h:375-440: regular merge, or small-region merge after a failed test), and
//     the bigger side survives.  So the evolution of the SMALL regions -- who merges with whom, which
//     piece reaches which big region ("hub") at which edge, the pieces' own means and flags -- does
//     not depend on the hubs' records at all.  Small regions joined by small-small edges form
//     sub-clusters; ONE THREAD PER SUB-CLUSTER replays its edges in order with the exact decision
//     tree and logs, per edge, what it does to a hub: ABSORB (hub, piece) or BB (two hubs meet: a
//     static big-big edge, or two pieces of the sub-cluster sitting in different hubs).  A piece
//     that grows to min_region_size is a hub from that edge on (its record is final for the thread).
//   * the hubs' records -- means, flags, constraint ids, big-big merges -- evolve along the logged
//     events only: one thread replays the events in reference order (absorb: ~100 cycles with the
//     hub's record in registers and the gates taken on the squared distance; big-big: the generic
//     decision tree).
//   * the one conditional case is a CONSTRAINED piece meeting a hub (same id: merge only within 0.15;
//     other id: nothing; unconstrained hub: it takes the id).  The replay thread predicts the outcome
//     from the hub's staged record -- decisive cases only: different ids, an unconstrained hub, or a
//     same-id distance more than 0.02 away from the gate -- keeps a snapshot of the piece and goes on;
//     the event thread re-decides the meeting with the hub's actual record and, if the piece's side of
//     the outcome differs from the prediction, the whole scan is redone by the one-thread walk
//     (nothing has been written back yet).  Non-decisive meetings, or more than 256 of them, suspend
//     the sub-cluster: the rest of its edges are decided by the event thread, in order.
// Same decisions, same float operations per merge as the one-thread walk; only independent work is
// reordered.
// ---------------------------------------------------------------------------------------------
constexpr unsigned kEvNone = 0, kEvAbsorb = 1, kEvBB = 2, kEvCond = 3;
constexpr int kSnapMax = 256;          // speculated conditional meetings per scan (two 4 KB banks of 32-byte snapshots)
struct __align__(16) CondSnap {        // the piece's record when it met the hub, and what the replay thread assumed
  int2 r0, r1, r2;                     // (sz, con) (d0, d1) (d2, fin)
  int piece;
  int flags;                           // 1: the hub is on the region_1 side, 2: assumed merged, 4: assumed the piece's id reset
};
__device__ void split_scan(const MergeParams& p, MergeShared& S, const int b, const uint32_t* codes, const uint32_t* pend_list,
                           const int n_pend, unsigned char* done_flags) {
  ScanShared& C = S.scan;
  const float edge_w = (float)b * (float)(1.0 / (double)bucket_scale());
  const bool force_bucket = edge_w < p.force_merge_weight;
  const float y_gate = force_bucket ? p.y_force : p.y_merge;
  const int mins = p.min_region_size;
  const int tid = threadIdx.x, nthr = blockDim.x;
  long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
  if (p.debug && tid == 0) t0 = clock64();
  // shared-memory views (everything below aliases arrays that are free once the local ids are looked up)
  unsigned char* const big = reinterpret_cast<unsigned char*>(C.tru);                  // [2 kScanMax] root is a hub (>= min size at staging)
  unsigned char* const gen = reinterpret_cast<unsigned char*>(C.trv) + kScanMax;       // [kScanMax] edge is decided by the event thread, generic path
  unsigned char* const evt = reinterpret_cast<unsigned char*>(C.trv) + 2 * kScanMax;   // [kScanMax] event type of the edge
  unsigned short* const ev_a = C.hidx;                    // [kScanMax] hub on the region_1 side / absorbing hub
  unsigned short* const ev_b = C.hidx + kScanMax;         // [kScanMax] hub on the region_2 side / absorbed piece
  unsigned short* const owner = C.hidx + 2 * kScanMax;    // [2 kScanMax] small root: hub it sits in (0xFFFF = free; itself = promoted)
  int* const sc = reinterpret_cast<int*>(C.hkey);         // [2 kScanMax] sub-cluster union-find over the small roots
  unsigned* const keys = C.hkey + 2 * kScanMax;           // [kScanMax] (sub-cluster << 11 | edge index), sorted
  unsigned short* const elist = reinterpret_cast<unsigned short*>(C.hkey + 3 * kScanMax);   // [kScanMax] edges the event thread looks at
  CondSnap* const snap_a = reinterpret_cast<CondSnap*>(C.hkey + 3 * kScanMax + kScanMax / 2);     // [128] (last 4 KB of hkey)
  CondSnap* const snap_b = reinterpret_cast<CondSnap*>(reinterpret_cast<unsigned char*>(C.tru) + 2 * kScanMax);   // [128] (upper half of tru)
  auto snap_at = [&](int k) -> CondSnap* { return k < kSnapMax / 2 ? snap_a + k : snap_b + (k - kSnapMax / 2); };
  // ---- staging: current roots of the edges -> compact local ids, records into shared memory ----
  int hbits = 6;
  while ((1 << hbits) < 4 * n_pend) ++hbits;
  for (int i = tid; i < (1 << hbits); i += nthr) { C.hkey[i] = 0u; C.hidx[i] = 0xFFFFu; }
  if (tid == 0) { C.n_roots = 0; C.hbits = hbits; C.n_snap = 0; C.spec_fail = 0; }
  __syncthreads();
  for (int i = tid; i < n_pend; i += nthr) {
    const uint32_t pos = pend_list[i];
    int u, v;
    decode_edge(p, codes[pos], u, v);
    const int ru = uf_find(p.parent, u), rv = uf_find(p.parent, v);
    C.epos[i] = pos;
    C.tru[i] = ru; C.trv[i] = rv;
    scan_insert(C, (unsigned)ru);
    scan_insert(C, (unsigned)rv);
  }
  __syncthreads();
  for (int i = tid; i < n_pend; i += nthr) {
    C.ea[i] = scan_lookup(C, (unsigned)C.tru[i]);
    C.eb[i] = scan_lookup(C, (unsigned)C.trv[i]);
  }
  __syncthreads();
  const int n_roots = C.n_roots;
  for (int j = tid; j < n_roots; j += nthr) {
    C.par[j] = (unsigned short)j;
    const RegionRec R = load_rec(&p.rec[C.gid[j]]);
    scan_store(C.rec, j, R);
    big[j] = R.sz >= mins ? 1 : 0;
    owner[j] = 0xFFFFu;
    sc[j] = j;
  }
  for (int i = tid; i < n_pend; i += nthr) { gen[i] = 0; evt[i] = kEvNone; }
  __syncthreads();
  // ---- sub-clusters: small roots joined by small-small edges ----
  for (int i = tid; i < n_pend; i += nthr) {
    const int a = C.ea[i], bq = C.eb[i];
    if (a != bq && !big[a] && !big[bq]) cl_union(sc, a, bq);
  }
  __syncthreads();
  int npad = 32;
  while (npad < n_pend) npad <<= 1;
  for (int i = tid; i < npad; i += nthr) {
    unsigned key = 0xFFFFFFFFu;
    if (i < n_pend) {
      const int a = C.ea[i], bq = C.eb[i];
      if (a != bq) {
        if (big[a] && big[bq]) { evt[i] = kEvBB; ev_a[i] = (unsigned short)a; ev_b[i] = (unsigned short)bq; }
        else key = ((unsigned)cl_find(sc, big[a] ? bq : a) << 11) | (unsigned)i;
      }
    }
    keys[i] = key;
  }
  __syncthreads();
  // bitonic sort of the keys: a sub-cluster's edges end up next to each other, in reference order
  for (int k = 2; k <= npad; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = tid; t < npad / 2; t += nthr) {
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;
        const unsigned x = keys[lo], y = keys[hi];
        const bool up = ((lo & k) == 0);
        if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
      }
      __syncthreads();
    }
  if (p.debug && tid == 0) t1 = clock64();
  // ---- one thread per sub-cluster: replay its edges in order ----
  for (int q0 = tid; q0 < n_pend; q0 += nthr) {
    const unsigned k0 = keys[q0];
    if (k0 == 0xFFFFFFFFu) continue;
    if (q0 > 0 && (keys[q0 - 1] >> 11) == (k0 >> 11)) continue;       // not the first edge of its sub-cluster
    auto findl = [&](int x) { int q = C.par[x]; while (q != x) { const int g = C.par[q]; C.par[x] = (unsigned short)g; x = q; q = g; } return x; };
    for (int q = q0; q < n_pend; ++q) {
      const unsigned key = keys[q];
      if ((key >> 11) != (k0 >> 11)) break;
      const int i = (int)(key & 2047u);
      const int ra = findl(C.ea[i]), rb = findl(C.eb[i]);
      if (ra == rb) continue;
      const int ha = big[ra] ? ra : (owner[ra] != 0xFFFFu ? (int)owner[ra] : -1);
      const int hb = big[rb] ? rb : (owner[rb] != 0xFFFFu ? (int)owner[rb] : -1);
      if (ha >= 0 && hb >= 0) {
        if (ha != hb) { ev_a[i] = (unsigned short)ha; ev_b[i] = (unsigned short)hb; evt[i] = kEvBB; }
        continue;
      }
      if (ha >= 0 || hb >= 0) {
        const int h = ha >= 0 ? ha : hb, x = ha >= 0 ? rb : ra;
        if (C.rec[x][0].y >= 0) {
          // a constrained piece meets a hub: conditional on the hub's record.  Decisive cases are predicted from the hub's
          // staged record and verified by the event thread; the others suspend the sub-cluster.
          const bool hub_is_a = ha >= 0;
          RegionRec H0 = scan_load(C.rec, h), Xc = scan_load(C.rec, x);
          const RegionRec Xb = Xc;
          bool decisive = true;
          if (H0.con >= 0 && H0.con == Xb.con) {
            const float dist = raw_dist(H0, Xb);
            const float gate = force_bucket ? 0.2f : 0.15f;
            decisive = fabsf(dist - gate) > 0.02f;
          }
          int k = kSnapMax;
          if (decisive) k = atomicAdd(&C.n_snap, 1);
          if (k < kSnapMax) {
            const int r = hub_is_a ? decide_pair(p, H0, Xc, edge_w) : decide_pair(p, Xc, H0, edge_w);
            const bool merged = r != 0;
            const bool reset = !merged && Xc.con != Xb.con;
            CondSnap* sn = snap_at(k);
            sn->r0 = make_int2(Xb.sz, Xb.con);
            sn->r1 = make_int2(__float_as_int(Xb.d0), __float_as_int(Xb.d1));
            sn->r2 = make_int2(__float_as_int(Xb.d2), Xb.fin);
            sn->piece = x;
            sn->flags = (hub_is_a ? 1 : 0) | (merged ? 2 : 0) | (reset ? 4 : 0);
            ev_a[i] = (unsigned short)h; ev_b[i] = (unsigned short)k; evt[i] = kEvCond;
            if (merged) owner[x] = (unsigned short)h;
            else if (reset) C.rec[x][0].y = Xc.con;
            continue;
          }
          for (int q2 = q; q2 < n_pend; ++q2) {
            const unsigned k2 = keys[q2];
            if ((k2 >> 11) != (k0 >> 11)) break;
            gen[k2 & 2047u] = 1;
          }
          break;
        }
        ev_a[i] = (unsigned short)h; ev_b[i] = (unsigned short)x; evt[i] = kEvAbsorb;
        owner[x] = (unsigned short)h;
        continue;
      }
      RegionRec A = scan_load(C.rec, ra), B = scan_load(C.rec, rb);
      const int r = decide_pair(p, A, B, edge_w);        // rep_1 = root of region_1 (the anchor), as in the reference
      if (r != 2) scan_store(C.rec, ra, A);
      if (r != 1) scan_store(C.rec, rb, B);
      if (r == 1) { C.par[rb] = (unsigned short)ra; if (A.sz >= mins) owner[ra] = (unsigned short)ra; }
      else if (r == 2) { C.par[ra] = (unsigned short)rb; if (B.sz >= mins) owner[rb] = (unsigned short)rb; }
    }
  }
  __syncthreads();
  if (p.debug && tid == 0) t2 = clock64();
  // ---- edges with something for the event thread, in order ----
  int n_ev = 0;
  for (int base = 0; base < n_pend; base += nthr) {
    const int i = base + tid;
    const bool flag = i < n_pend && (evt[i] != kEvNone || gen[i]);
    unsigned total;
    const unsigned rank = block_rank(S, flag, &total);
    if (flag) elist[n_ev + rank] = (unsigned short)i;
    n_ev += (int)total;
  }
  __syncthreads();
  // ---- event thread(s): warp 0 walks the events in order.  The 32 lanes fetch 32 events (ids, the absorbed piece's
  // record) at once; the events are then replayed one by one with the current hub's record held in registers by every
  // lane (uniform, redundant arithmetic: the serial chain per absorb is a root lookup plus ~20 float operations);
  // big-big and generic events are decided by lane 0 with the full decision tree. ----
  if (tid < 32) {
    const unsigned lane = tid;
    auto findl = [&](int x) { int q = C.par[x]; while (q != x) { const int g = C.par[q]; C.par[x] = (unsigned short)g; x = q; q = g; } return x; };
    int cur = -1;                // hub whose record is held in registers
    RegionRec H;
    H.sz = 0; H.con = -1; H.d0 = H.d1 = H.d2 = 0.f; H.fin = 0; H.pad0 = H.pad1 = 0;
    unsigned long long n_abs = 0, n_bb = 0, n_gen = 0, n_swap = 0;
    for (int e0 = 0; e0 < n_ev; e0 += 32) {
      const int cnt = min(32, n_ev - e0);
      int my_i = 0, my_kind = 0, my_a = 0, my_b = 0, my_sz = 0, my_fin = 0;      // kind: 0 absorb, 1 big-big, 2 generic
      float my_d0 = 0.f, my_d1 = 0.f, my_d2 = 0.f;
      if ((int)lane < cnt) {
        my_i = elist[e0 + lane];
        if (gen[my_i]) { my_kind = 2; my_a = C.ea[my_i]; my_b = C.eb[my_i]; }
        else {
          my_a = ev_a[my_i]; my_b = ev_b[my_i];
          if (evt[my_i] == kEvCond) my_kind = 3;
          else if (evt[my_i] == kEvAbsorb) {
            const int2 x0 = C.rec[my_b][0], x1 = C.rec[my_b][1], x2 = C.rec[my_b][2];      // (sz, con) (d0, d1) (d2, fin); con < 0
            my_sz = x0.x; my_fin = x2.y;
            my_d0 = __int_as_float(x1.x); my_d1 = __int_as_float(x1.y); my_d2 = __int_as_float(x2.x);
          } else my_kind = 1;
        }
      }
      for (int t = 0; t < cnt; ++t) {
        const int kind = __shfl_sync(0xffffffffu, my_kind, t);
        const int ea_ = __shfl_sync(0xffffffffu, my_a, t), eb_ = __shfl_sync(0xffffffffu, my_b, t);
        if (kind == 0) {
          const int xsz = __shfl_sync(0xffffffffu, my_sz, t), xfin = __shfl_sync(0xffffffffu, my_fin, t);
          const float xd0 = __shfl_sync(0xffffffffu, my_d0, t), xd1 = __shfl_sync(0xffffffffu, my_d1, t), xd2 = __shfl_sync(0xffffffffu, my_d2, t);
          const int h = findl(ea_);
          if (h != cur) {
            if (cur >= 0 && lane == 0) scan_store(C.rec, cur, H);
            __syncwarp();
            H = scan_load(C.rec, h);
            cur = h;
            ++n_swap;
          }
          if (!H.fin && !xfin) {
            // the test of segmentation_graph.h:377-390 on the squared distance (sqrtf is monotone; gates precomputed)
            const float d1 = H.d0 - xd0, d2 = H.d1 - xd1, d3 = H.d2 - xd2;
            const float y = (d1 * d1 + d2 * d2 + d3 * d3) * (1.0f / 3.0f);
            if (!(y < y_gate)) H.fin = 1;
          }
          // MergeRegions + MergeDescriptor: the hub is the bigger side (the piece is below min size)
          const float denom = __frcp_rn((float)(xsz + H.sz));
          const float fa = (float)xsz * denom, fb = (float)H.sz * denom;
          H.d0 = fa * xd0 + fb * H.d0;
          H.d1 = fa * xd1 + fb * H.d1;
          H.d2 = fa * xd2 + fb * H.d2;
          H.sz += xsz;
          if (lane == 0) C.par[eb_] = (unsigned short)h;
          ++n_abs;
          continue;
        }
        // big-big / generic / speculated meetings: lane 0 with the full decision tree, on the records in shared memory
        if (cur >= 0) { if (lane == 0) scan_store(C.rec, cur, H); cur = -1; }
        if (kind == 3) {
          int fail = 0;
          if (lane == 0) {
            const CondSnap sn = *snap_at(eb_);
            const int h = findl(ea_);
            RegionRec Hh = scan_load(C.rec, h), X;
            X.sz = sn.r0.x; X.con = sn.r0.y; X.d0 = __int_as_float(sn.r1.x); X.d1 = __int_as_float(sn.r1.y); X.d2 = __int_as_float(sn.r2.x);
            X.fin = sn.r2.y; X.pad0 = X.pad1 = 0;
            const int xcon0 = X.con;
            const int r = (sn.flags & 1) ? decide_pair(p, Hh, X, edge_w) : decide_pair(p, X, Hh, edge_w);
            const bool merged = r != 0, reset = !merged && X.con != xcon0;
            if (merged != ((sn.flags & 2) != 0) || reset != ((sn.flags & 4) != 0)) { fail = 1; C.spec_fail = 1; }
            else {
              scan_store(C.rec, h, Hh);
              if (merged) C.par[sn.piece] = (unsigned short)h;
            }
          }
          fail = __shfl_sync(0xffffffffu, fail, 0);
          ++n_gen;
          if (fail) { e0 = n_ev; break; }
          continue;
        }
        if (lane == 0) {
          const int a = findl(ea_), bq = findl(eb_);
          if (a != bq) {
            RegionRec A = scan_load(C.rec, a), B = scan_load(C.rec, bq);
            const bool both_con = A.con >= 0 && B.con >= 0;
            const bool nothing = (both_con && A.con != B.con) || (!both_con && (A.fin || B.fin) && A.sz >= mins && B.sz >= mins);
            if (!nothing) {
              const int r = decide_pair(p, A, B, edge_w);        // rep_1 = root of region_1 (the anchor), as in the reference
              if (r != 2) scan_store(C.rec, a, A);
              if (r != 1) scan_store(C.rec, bq, B);
              if (r == 1) C.par[bq] = (unsigned short)a;
              else if (r == 2) C.par[a] = (unsigned short)bq;
            }
          }
        }
        __syncwarp();
        if (kind == 1) ++n_bb; else ++n_gen;
      }
    }
    if (cur >= 0 && lane == 0) scan_store(C.rec, cur, H);
    if (p.debug && lane == 0) {
      t3 = clock64();
      atomicAdd(&p.debug[kNumBuckets * 4 + 57], n_abs); atomicAdd(&p.debug[kNumBuckets * 4 + 58], n_bb);
      atomicAdd(&p.debug[kNumBuckets * 4 + 59], n_gen); atomicAdd(&p.debug[kNumBuckets * 4 + 60], n_swap);
    }
  }
  __syncthreads();
  if (C.spec_fail) {
    // a predicted meeting came out differently with the hub's actual record: nothing has been written back, the
    // one-thread walk redoes the list from global memory
    if (p.debug && tid == 0) atomicAdd(&p.debug[kNumBuckets * 4 + 62], 1ull);
    __syncthreads();
    exact_scan(p, C, b, codes, pend_list, n_pend, done_flags, -1, nullptr, nullptr);
    return;
  }
  for (int j = tid; j < n_roots; j += nthr) {
    int x = j;
    while (C.par[x] != x) x = C.par[x];
    const int g = C.gid[j];
    if (x == j) store_rec(&p.rec[g], scan_load(C.rec, j));
    else p.parent[g] = C.gid[x];
  }
  for (int i = tid; i < n_pend; i += nthr) done_flags[C.epos[i]] = 1;
  if (tid == 0) atomicAdd(&p.stats[3], 1ull);
  __syncthreads();
  if (p.debug && tid == 0) atomicAdd(&p.debug[kNumBuckets * 4 + 61], (unsigned long long)min(C.n_snap, kSnapMax));
  if (p.debug && tid == 0) {      // development tap: cycles per stage
    atomicAdd(&p.debug[kNumBuckets * 4 + 50], (unsigned long long)(t3 - t2));       // event thread
    atomicAdd(&p.debug[kNumBuckets * 4 + 51], (unsigned long long)(clock64() - t0));  // whole scan
    atomicAdd(&p.debug[kNumBuckets * 4 + 52], (unsigned long long)n_roots);
    atomicAdd(&p.debug[kNumBuckets * 4 + 53], (unsigned long long)(t1 - t0));       // staging + sub-clusters + sort
    atomicAdd(&p.debug[kNumBuckets * 4 + 54], (unsigned long long)(t2 - t1));       // sub-cluster replay
    atomicAdd(&p.debug[kNumBuckets * 4 + 55], (unsigned long long)n_ev);
    atomicAdd(&p.debug[kNumBuckets * 4 + 56], (unsigned long long)n_pend);
  }
}

struct RoundState { unsigned epoch, buf; bool from_master; unsigned long long n_src, prev_live; };

// Ordered rounds on the pending edges of segment [seg_lo, seg_hi): deterministic reservations (the
// earliest pending edge of both of its roots runs the exact serial decision tree) + absorption by
// frozen hubs.  Lists ping-pong between live_b (buf 0) and live_c (buf 1); the first round reads the
// window's master list.  Returns 0 = segment finished, 1 = dependency chain (progress per round below
// the threshold; the list of the last P1 is in buffer st.buf), 2 = list not larger than small_limit
// (handed over at a round boundary), 3 = watchdog.
template <class Bar>
__device__ int ordered_rounds(const MergeParams& p, Bar& bar, const unsigned tid, const unsigned nthr, const int b,
                              const float edge_w, const int wtag, const uint32_t* master, const unsigned long long seg_lo,
                              const unsigned long long seg_hi, RoundState& st, const unsigned long long small_limit,
                              const unsigned long long stall_progress, unsigned long long& guard) {
  const int mins = p.min_region_size;
  while (true) {
    if (++guard > (1ull << 22)) { if (tid == 0) { printf("vsb200 merge: round watchdog bucket %d\n", b); p.stats[7] = 1ull; } return 3; }
    const unsigned long long key_hi = ((unsigned long long)(0xFFFFFFFFu - st.epoch)) << 32;
    uint32_t* dst = st.buf ? p.live_c : p.live_b;
    const uint32_t* src = st.from_master ? master : (st.buf ? p.live_b : p.live_c);
    unsigned long long* dst_cnt = &p.counters[st.buf];
    // ---- P1: find roots, drop inert edges, reserve ----
    const unsigned long long i_begin = st.from_master ? seg_lo : 0ull, i_end = st.from_master ? seg_hi : st.n_src;
    for (unsigned long long i = i_begin + tid; i < i_end; i += nthr) {
      const uint4 e0 = reinterpret_cast<const uint4*>(src)[i];
      const uint32_t code = e0.x, pos = e0.w;
      if (st.from_master) {
        if (p.done[pos]) continue;
      } else if (code == kDone) continue;          // executed in the previous round's commit
      // the entry carries the roots of an earlier pass: climbing from them is shorter than from the voxels
      const int ru = uf_find(p.parent, (int)e0.y), rv = uf_find(p.parent, (int)e0.z);
      bool drop = (ru == rv);
      bool bigbig = false;
      if (!drop) {
        const RegionRec A = load_rec(&p.rec[ru]), B = load_rec(&p.rec[rv]);
        const bool both_con = (A.con >= 0 && B.con >= 0);
        // (edges between different constraint ids are only here as guards: they wait for their turn like any other edge)
        drop = (!both_con && (A.fin || B.fin) && A.sz >= mins && B.sz >= mins);
        bigbig = !drop && A.sz >= mins && B.sz >= mins;
      }
      if (drop) { p.done[pos] = p.has_constraints ? 3 : 1; continue; }      // constrained chunks: dormant, the refresh looks at it again
      if (bigbig) {
        // a pending edge between two big regions: absorptions into either of them that come later in
        // reference order wait for it (it may merge the hub away in the round it executes)
        atomicMin(&p.hull[ru].num, key_hi | code);
        atomicMin(&p.hull[rv].num, key_hi | code);
      }
      const unsigned long long slot = warp_slot(dst_cnt);
      if (slot < p.live_cap) {
        reinterpret_cast<uint4*>(dst)[slot] = make_uint4(code, (uint32_t)ru, (uint32_t)rv, pos);
        atomicMin(&p.res[ru], key_hi | code);
        atomicMin(&p.res[rv], key_hi | code);
      }
    }
    bar.sync();
    unsigned long long n_live = *((volatile unsigned long long*)dst_cnt);
    if (n_live > p.live_cap) n_live = p.live_cap;
    if (n_live == 0) return 0;
    if (tid == 0) { trace(p, 0, (unsigned long long)b); trace(p, 1, guard); trace(p, 2, n_live); trace(p, 3, 1); }
    const unsigned long long need = stall_progress ? stall_progress : max(8ull, n_live >> 6);
    if (st.prev_live - n_live < need) return 1;                      // chain regime
    if (small_limit && n_live <= small_limit) {
      // hand the list over at a round boundary: the next P1 (block 0) re-reads it
      if (tid == 0) p.counters[st.buf ^ 1] = 0ull;
      bar.sync();
      st.n_src = n_live; st.from_master = false; st.prev_live = ~0ull >> 1; st.buf ^= 1; ++st.epoch;
      return 2;
    }
    st.prev_live = n_live;
    // ---- P3: commit (strict owners + absorption by frozen hubs) ----
    for (unsigned long long i = tid; i < n_live; i += nthr) {
      const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
      const int ru = (int)e.y, rv = (int)e.z;
      const unsigned long long key = key_hi | e.x;
      const bool own_u = (p.res[ru] == key), own_v = (p.res[rv] == key);
      bool done = false;
      if (own_u && own_v) {
        exec_strict(p, ru, rv, edge_w, p.stats);
        done = true;
      } else if (own_u || own_v) {
        const RegionRec A = load_rec(&p.rec[ru]), B = load_rec(&p.rec[rv]);
        // the edge is the next edge of the SMALL side it owns; the other side is a hub whose
        // decision-relevant state cannot change in this segment (and that no earlier big-big edge can
        // merge away in this round).  Only small regions are absorbed this way: a big one may itself be
        // collecting absorptions in this round, which would be lost with it.
        if (own_v && B.con < 0 && A.sz >= mins && B.sz < mins && (A.fin || p.hull[ru].frozen == wtag)) {
          const unsigned long long bb = p.hull[ru].num;
          if (!((bb >> 32) == (key_hi >> 32) && (uint32_t)bb < e.x)) { p.parent[rv] = ru; acc_add(p.acc, ru, B); done = true; }
        } else if (own_u && A.con < 0 && B.sz >= mins && A.sz < mins && (B.fin || p.hull[rv].frozen == wtag)) {
          const unsigned long long bb = p.hull[rv].num;
          if (!((bb >> 32) == (key_hi >> 32) && (uint32_t)bb < e.x)) { p.parent[ru] = rv; acc_add(p.acc, rv, A); done = true; }
        }
      }
      if (done) { p.done[e.w] = 1; reinterpret_cast<uint4*>(dst)[i].x = kDone; }
    }
    bar.sync();
    // ---- P4: fold bulk contributions ----
    for (unsigned long long i = tid; i < n_live; i += nthr) {
      const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
      acc_fold(p, (int)e.y);
      acc_fold(p, (int)e.z);
    }
    if (tid == 0) {
      p.counters[st.buf ^ 1] = 0ull;          // the other buffer becomes the next destination
      atomicAdd(&p.stats[0], 1ull);
    }
    bar.sync();
    st.n_src = n_live;
    st.from_master = false;
    st.buf ^= 1;
    ++st.epoch;
  }
}

// counters: [0] live count of buffer A, [1] of buffer B, [2] round epoch, [3] serial result flag,
//           [4] first hub-hub position of the window, [5] uncertified edge count, [6] window tag
// live entry = 4 words: code, ru, rv, position in the window
// development tap: wall time per phase (global thread 0), added to debug[kNumBuckets * 4 + 20 + slot]
#define VSB_CPASS(slot) do { if (p.debug && tid == 0 && blockIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.debug[kNumBuckets * 4 + 40 + (slot)] += t_ - t_cpass; t_cpass = t_; } } while (0)
#define VSB_PHASE(slot) do { if (p.debug && tid == 0 && blockIdx.x == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); p.debug[kNumBuckets * 4 + 20 + (slot)] += t_ - t_phase; t_phase = t_; } } while (0)

template <class Bar, bool kIsGrid>
__device__ void run_bucket(const MergeParams& p, Bar& bar, MergeShared& S, const unsigned tid, const unsigned nthr,
                           const int b, const uint32_t* bucket_codes, const unsigned long long bucket_edges) {
  const float inv_scale = (float)(1.0 / (double)bucket_scale());   // segmentation_graph.h:348
  const float edge_w = (float)b * inv_scale;
  const bool force_bucket = edge_w < p.force_merge_weight;
  WindowThr wt;
  wt.thr_m = (force_bucket ? 0.2f : 0.05f) * 0.999f - 2e-5f;
  wt.con_thr = force_bucket ? wt.thr_m : (0.15f * 0.999f - 2e-5f);
  wt.strict_con = (p.dev_flags >> 8) & 3;
  const int mins = p.min_region_size;
  unsigned epoch = (unsigned)(*((volatile unsigned long long*)&p.counters[2]));
  int wtag = (int)(*((volatile unsigned long long*)&p.counters[6]));
  bar.sync();                                   // everybody has read the persistent counters
  unsigned long long w0 = 0;
  const unsigned long long kWindowTarget = p.window_target, kResidualSplit = p.residual_split, kSegmentMin = p.segment_min;
  unsigned long long target = kWindowTarget;   // live edges aimed at per window: grows while windows certify cleanly
  unsigned long long raw = min(bucket_edges, kWindowTarget);
  const bool tiny_windows = p.has_constraints && (p.dev_flags & 64);      // diagnostic: prune state (nearly) exact
  const bool no_cert = p.has_constraints && (p.dev_flags & 32);           // diagnostic: every edge through rounds + exact scans
  if (tiny_windows) { target = 4096; raw = min(bucket_edges, 4096ull); }
  unsigned long long guard = 0;
  uint32_t* const master = p.live_a;            // live list of the window (code, ru, rv, position)
  unsigned long long t_phase = 0, t_cpass = 0;
  if (p.debug && tid == 0 && blockIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_phase));
  const unsigned lane = threadIdx.x & 31u;
  while (w0 < bucket_edges) {
    const unsigned long long n_edges = min(raw, bucket_edges - w0);
    const uint32_t* codes = bucket_codes + w0;
    if (tid == 0) p.counters[4] = ~0ull;
    // ---- P1: roots, inert edges, first hub-hub edge -> the window's live list in REFERENCE ORDER.
    // Every warp takes a contiguous chunk of positions (kP1U x 32 edges in flight: the pass is bound by
    // the latency of the dependent parent / record loads), stages the survivors at their position and
    // counts them; after a barrier the chunks are compacted behind each other (stable). ----
    const unsigned n_warps = nthr >> 5, gw = tid >> 5;
    const unsigned long long chunk = (((n_edges + n_warps - 1) / n_warps) + 31ull) & ~31ull;
    const unsigned long long c0 = min(n_edges, (unsigned long long)gw * chunk), c1 = min(n_edges, c0 + chunk);
    uint4* const staging = reinterpret_cast<uint4*>(p.live_c);
    uint32_t* const keep_mask = p.live_b;                         // one word per 32 positions
    uint32_t* const hub_mask = p.live_b + (n_edges >> 5) + 1;
    unsigned long long* const warp_cnt = p.counters + 16;
    {
      unsigned my_count = 0, my_live = 0;
      for (unsigned long long base = c0; base < c1; base += 32ull * kP1U) {
        uint32_t code[kP1U];
        int us[kP1U], vs[kP1U], pu[kP1U], pv[kP1U], rus[kP1U], rvs[kP1U];
        bool in[kP1U], keep[kP1U], hubf[kP1U], livef[kP1U];
#pragma unroll
        for (int k = 0; k < kP1U; ++k) {
          const unsigned long long i = base + 32ull * k + lane;
          in[k] = i < c1;
          code[k] = in[k] ? __ldg(&codes[i]) : 0u;
        }
#pragma unroll
        for (int k = 0; k < kP1U; ++k) {
          us[k] = 0; vs[k] = 0;
          if (in[k]) decode_edge(p, code[k], us[k], vs[k]);
        }
#pragma unroll
        for (int k = 0; k < kP1U; ++k) { pu[k] = p.parent[us[k]]; pv[k] = p.parent[vs[k]]; }
#pragma unroll
        for (int k = 0; k < kP1U; ++k) {
          rus[k] = (pu[k] == us[k]) ? us[k] : uf_find(p.parent, pu[k]);
          rvs[k] = (pv[k] == vs[k]) ? vs[k] : uf_find(p.parent, pv[k]);
          if (in[k]) {
            if (pu[k] != us[k] && rus[k] != pu[k]) p.parent[us[k]] = rus[k];     // path compression
            if (pv[k] != vs[k] && rvs[k] != pv[k]) p.parent[vs[k]] = rvs[k];
          }
        }
        int4 a0[kP1U], b0[kP1U];      // first halves of the two records (sz, con, d0, d1) and fin words
        int af[kP1U], bf[kP1U];
#pragma unroll
        for (int k = 0; k < kP1U; ++k) {
          const bool need = in[k] && rus[k] != rvs[k];
          const int ia = need ? rus[k] : 0, ib = need ? rvs[k] : 0;
          a0[k] = reinterpret_cast<const int4*>(&p.rec[ia])[0]; af[k] = p.rec[ia].fin;
          b0[k] = reinterpret_cast<const int4*>(&p.rec[ib])[0]; bf[k] = p.rec[ib].fin;
        }
#pragma unroll
        for (int k = 0; k < kP1U; ++k) {
          keep[k] = false; hubf[k] = false; livef[k] = false;
          if (!in[k]) continue;
          const unsigned long long i = base + 32ull * k + lane;
          bool drop = (rus[k] == rvs[k]);
          bool dormant = false;
          if (!drop) {
            const int asz = a0[k].x, acon = a0[k].y, bsz = b0[k].x, bcon = b0[k].y;
            const bool both_con = (acon >= 0 && bcon >= 0);
            // different constraint ids: nothing happens NOW, but a split may reset one of the ids before the scan
            // reaches the edge (segmentation_graph.h:417-437), so the entry stays in the list as dormant (done = 3)
            // (constrained chunks: an inert pair, one side finalised and both big, may still be FORCED together once both
            // sides carry one constraint id -- it stays dormant as well)
            const bool inert = !both_con && (af[k] || bf[k]) && asz >= mins && bsz >= mins;
            dormant = (both_con && acon != bcon) || (inert && p.has_constraints);
            drop = dormant || inert;
            hubf[k] = !drop && asz >= mins && bsz >= mins;
          }
          p.done[i] = dormant ? 3 : (drop ? 1 : 0);   // done flags of the window are (re)written here
          keep[k] = !drop || dormant;
          livef[k] = !drop;
          if (keep[k]) staging[i] = make_uint4(code[k], (uint32_t)rus[k], (uint32_t)rvs[k], (uint32_t)i);
        }
#pragma unroll
        for (int k = 0; k < kP1U; ++k) {           // warp-uniform: chunk bounds are multiples of 32
          const unsigned long long g0 = base + 32ull * k;
          const unsigned m = __ballot_sync(0xffffffffu, keep[k]);
          const unsigned m2 = __ballot_sync(0xffffffffu, hubf[k]);
          if (lane == 0 && g0 < c1) { keep_mask[g0 >> 5] = m; hub_mask[g0 >> 5] = m2; }
          my_count += __popc(m);
          my_live += __popc(__ballot_sync(0xffffffffu, livef[k]));
        }
      }
      // two-level counts: warps of a block in shared memory, blocks in global memory
      const unsigned wib = threadIdx.x >> 5;
      if (lane == 0) { S.warp_cnt[wib] = my_count; S.warp_live[wib] = my_live; }
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned long long t = 0, tl = 0;
        for (int k = 0; k < kMergeWarps; ++k) { t += S.warp_cnt[k]; tl += S.warp_live[k]; }
        warp_cnt[kIsGrid ? blockIdx.x : 0u] = t;
        warp_cnt[1024 + (kIsGrid ? blockIdx.x : 0u)] = tl;
      }
    }
    bar.sync();
    unsigned long long n_master = 0, n_live0 = 0;
    {
      const unsigned nblk = kIsGrid ? gridDim.x : 1u, blk = kIsGrid ? blockIdx.x : 0u, wib = threadIdx.x >> 5;
      unsigned long long before = 0, total = 0;
      unsigned long long total_live = 0;
      for (unsigned j = lane; j < nblk; j += 32) { const unsigned long long c = *((volatile unsigned long long*)&warp_cnt[j]); total += c; if (j < blk) before += c; total_live += *((volatile unsigned long long*)&warp_cnt[1024 + j]); }
      for (int o = 16; o > 0; o >>= 1) { before += __shfl_xor_sync(0xffffffffu, before, o); total += __shfl_xor_sync(0xffffffffu, total, o); total_live += __shfl_xor_sync(0xffffffffu, total_live, o); }
      n_live0 = total_live;
      for (unsigned k = 0; k < wib; ++k) before += S.warp_cnt[k];
      n_master = total;
      unsigned long long off = before;
      for (unsigned long long g0 = c0; g0 < c1; g0 += 32) {
        const unsigned m = keep_mask[g0 >> 5], m2 = hub_mask[g0 >> 5];
        if ((m >> lane) & 1u) {
          const unsigned long long idx = off + __popc(m & ((1u << lane) - 1u));
          reinterpret_cast<uint4*>(master)[idx] = staging[g0 + lane];
          if ((m2 >> lane) & 1u) atomicMin(&p.counters[4], idx);
        }
        off += __popc(m);
      }
    }
    bar.sync();
    VSB_PHASE(0);                                  // raw prune pass
    unsigned long long hh = *((volatile unsigned long long*)&p.counters[4]);   // index of the first live edge between two hubs
    // ---------------- segments of the window: [seg_lo, seg_hi) are INDICES into the ordered live list ----------------
    unsigned long long seg_lo = 0, seg_len = n_master;
    bool any_split = false;
    unsigned long long unc_sum = 0;
    while (seg_lo < n_master) {
      if (++guard > (1ull << 22)) { if (tid == 0) { printf("vsb200 merge: window watchdog bucket %d\n", b); p.stats[7] = 1ull; } return; }
      const unsigned long long seg_end = min(min(hh, n_master), seg_lo + seg_len);   // a hub-hub edge ends the segment in front of it
      unsigned long long seg_hi = seg_end;
      bool have_live = false;
      // ---- certification of [seg_lo, seg_hi), halving the segment while too much stays uncertified ----
      while (seg_hi > seg_lo) {
        ++wtag;
        if (tid == 0) { p.counters[5] = 0ull; p.counters[7] = 0ull; }
        bar.sync();
        if (p.debug && tid == 0 && blockIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_cpass));
#define VSB_IN_SEG(e) (!p.done[(e).w])
        // ---- C1: sub-clusters of small atoms ----
        {
          unsigned long long mine = 0;
          for (unsigned long long i = seg_lo + tid; i < seg_hi; i += nthr) {
            const uint4 e = reinterpret_cast<const uint4*>(master)[i];
            if (!VSB_IN_SEG(e)) continue;
            ++mine;
            const int sa = p.rec[e.y].sz, sb = p.rec[e.z].sz;
            if (sa < mins && sb < mins) cl_union(p.cl, (int)e.y, (int)e.z);
          }
          for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
          if (lane == 0 && mine) atomicAdd(&p.counters[7], mine);
        }
        bar.sync();
        VSB_CPASS(0);
        if (*((volatile unsigned long long*)&p.counters[7]) == 0ull) break;      // nothing live in this segment
        have_live = true;
        // ---- C2: sub-cluster records (once per atom) and hub adjacency.  Atoms of one sub-cluster
        // sit next to each other in reference order: lanes that update the same record are combined
        // with __match_any_sync / __reduce_*_sync, one atomic per group. ----
        const int tag_a = 2 * wtag, tag_b = 2 * wtag + 1;
        for (unsigned long long i0 = seg_lo + (tid - lane); i0 < seg_hi; i0 += nthr) {
          const unsigned long long i = i0 + lane;
          bool in = i < seg_hi;
          uint4 e = make_uint4(kDone, 0u, 0u, 0u);
          if (in) { e = reinterpret_cast<const uint4*>(master)[i]; in = VSB_IN_SEG(e); }
          int hub = -1, sub = -1;
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            bool valid = false;
            int c = -1;
            RegionRec R;
            R.sz = 0; R.con = -1; R.d0 = R.d1 = R.d2 = 0.f; R.fin = 0; R.pad0 = R.pad1 = 0;
            if (in) {
              const int r = s2 ? (int)e.z : (int)e.y;
              R = load_rec(&p.rec[r]);
              if (R.sz >= mins) hub = r;
              else {
                c = cl_find_compress(p.cl, r);
                sub = c;
                valid = atomicExch(&p.hull[r].claim, tag_a) != tag_a;
              }
            }
            const unsigned act = __ballot_sync(0xffffffffu, valid);
            if (valid) {
              const unsigned peers = __match_any_sync(act, c);
              const int m0 = __reduce_min_sync(peers, __float_as_int(R.d0)), x0 = __reduce_max_sync(peers, __float_as_int(R.d0));
              const int m1 = __reduce_min_sync(peers, __float_as_int(R.d1)), x1 = __reduce_max_sync(peers, __float_as_int(R.d1));
              const int m2 = __reduce_min_sync(peers, __float_as_int(R.d2)), x2 = __reduce_max_sync(peers, __float_as_int(R.d2));
              const int mass = __reduce_add_sync(peers, R.sz);
              const unsigned fin = __reduce_or_sync(peers, R.fin ? 1u : 0u);
              NodeScratch* sc = &p.hull[c];
              if (lane == (unsigned)(__ffs(peers) - 1)) {
                atomicMin(&sc->mn[0], m0); atomicMax(&sc->mx[0], x0);
                atomicMin(&sc->mn[1], m1); atomicMax(&sc->mx[1], x1);
                atomicMin(&sc->mn[2], m2); atomicMax(&sc->mx[2], x2);
                atomicAdd(&sc->mass, mass);
                if (fin) atomicOr(&sc->flags, kScFin);
              }
              if (R.con >= 0) {
                const int old = atomicCAS(&sc->con, kNoCon, R.con);
                if (old != kNoCon && old != R.con) atomicOr(&sc->flags, kScConMulti);
              }
            }
          }
          if (hub >= 0 && sub >= 0) {
            NodeScratch* sc = &p.hull[sub];
            int old = *((volatile int*)&sc->hub0);
            if (old == -1) old = atomicCAS(&sc->hub0, -1, hub);
            if (old != -1 && old != hub) {
              old = *((volatile int*)&sc->hub1);
              if (old == -1) old = atomicCAS(&sc->hub1, -1, hub);
              if (old != -1 && old != hub) atomicOr(&sc->flags, kScHubs3);
            }
          }
        }
        bar.sync();
        VSB_CPASS(1);
        // ---- C3: what every hub may absorb in this segment (once per atom and hub) ----
        for (unsigned long long i0 = seg_lo + (tid - lane); i0 < seg_hi; i0 += nthr) {
          const unsigned long long i = i0 + lane;
          bool in = i < seg_hi;
          uint4 e = make_uint4(kDone, 0u, 0u, 0u);
          if (in) { e = reinterpret_cast<const uint4*>(master)[i]; in = VSB_IN_SEG(e); }
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            bool valid = false;
            RegionRec R;
            R.sz = 0; R.con = -1; R.d0 = R.d1 = R.d2 = 0.f; R.fin = 0; R.pad0 = R.pad1 = 0;
            NodeScratch SC;
            SC.hub0 = SC.hub1 = -1; SC.flags = 0; SC.con = kNoCon;
            if (in) {
              const int r = s2 ? (int)e.z : (int)e.y;
              R = load_rec(&p.rec[r]);
              if (R.sz >= mins) {
                // hub side of a hub-small edge: a sub-cluster with more than two hubs makes all of them uncertain
                const int o = s2 ? (int)e.y : (int)e.z;
                if (p.rec[o].sz < mins) {
                  const int c = cl_find_compress(p.cl, o);
                  if ((p.hull[c].flags & kScHubs3) && (p.hull[r].flags & (kScUnc | kScUncAny)) != (kScUnc | kScUncAny))
                    atomicOr(&p.hull[r].flags, R.con >= 0 ? (kScUnc | kScUncAny) : kScUnc);
                }
              } else if (atomicExch(&p.hull[r].claim, tag_b) != tag_b) {
                SC = load_sc(&p.hull[cl_find_compress(p.cl, r)]);
                valid = true;
              }
            }
            // Two hubs sharing a sub-cluster may meet inside the segment (a dynamic big-big decision).  If
            // the rules forbid or neutralise the meeting (different constraint ids; one side finalised and not
            // both carrying one id) nothing happens.  Otherwise each hub counts the other as something it may
            // absorb: the frozen bound then also certifies that their meeting ends in a merge.  Done once per
            // sub-cluster, by the lane that claimed the sub-cluster's root atom.
            bool pair_unc_any = false;
            if (valid && SC.hub0 >= 0 && SC.hub1 >= 0 && !(SC.flags & kScHubs3) && cl_find_compress(p.cl, s2 ? (int)e.z : (int)e.y) == (s2 ? (int)e.z : (int)e.y)) {
              const RegionRec H0 = load_rec(&p.rec[SC.hub0]), H1 = load_rec(&p.rec[SC.hub1]);
              const bool both_con = H0.con >= 0 && H1.con >= 0;
              bool contribute = false;
              if (both_con) {
                if (H0.con == H1.con) { if (H0.fin != H1.fin) pair_unc_any = true; else contribute = true; }
              } else if (!H0.fin && !H1.fin) contribute = true;
              if (contribute && (p.dev_flags & 1)) { contribute = false; pair_unc_any = true; atomicOr(&p.hull[SC.hub0].flags, kScUnc); atomicOr(&p.hull[SC.hub1].flags, kScUnc); }
              if (contribute) {
                const int dbits = __float_as_int(raw_dist(H0, H1));
                NodeScratch* h0 = &p.hull[SC.hub0];
                NodeScratch* h1 = &p.hull[SC.hub1];
                atomicMax(&h0->rbits, dbits); atomicAdd(&h0->mass, H1.sz);
                atomicMax(&h1->rbits, dbits); atomicAdd(&h1->mass, H0.sz);
                if (H0.con < 0 && H1.con >= 0) { const int old = atomicCAS(&h0->con, kNoCon, H1.con); if (old != kNoCon && old != H1.con) atomicOr(&h0->flags, kScConMulti); }
                if (H1.con < 0 && H0.con >= 0) { const int old = atomicCAS(&h1->con, kNoCon, H0.con); if (old != kNoCon && old != H0.con) atomicOr(&h1->flags, kScConMulti); }
              }
              if (pair_unc_any) { atomicOr(&p.hull[SC.hub0].flags, kScUncAny); atomicOr(&p.hull[SC.hub1].flags, kScUncAny); }
            }
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              const int h = q ? SC.hub1 : SC.hub0;
              const bool v = valid && h >= 0;
              const unsigned act = __ballot_sync(0xffffffffu, v);
              if (!v) continue;
              const unsigned peers = __match_any_sync(act, h);
              const RegionRec H = load_rec(&p.rec[h]);
              const float d = raw_dist(H, R);
              const int rmax = __reduce_max_sync(peers, __float_as_int(d));
              const int mass = __reduce_add_sync(peers, R.sz);
              NodeScratch* hs = &p.hull[h];
              if (lane == (unsigned)(__ffs(peers) - 1)) {
                atomicMax(&hs->rbits, rmax);
                atomicAdd(&hs->mass, mass);
              }
              const int hflags = *((volatile int*)&hs->flags);
              if ((SC.flags & kScHubs3) && !(hflags & kScUnc)) atomicOr(&hs->flags, kScUnc);   // more than two hubs: not analysed
              if (((SC.flags & kScHubs3) && H.con >= 0) && !(hflags & kScUncAny)) atomicOr(&hs->flags, kScUncAny);
              if (SC.flags & kScConMulti) { if (!(hflags & kScConMulti)) atomicOr(&hs->flags, kScConMulti); }
              else if (SC.con != kNoCon) {
                int old = *((volatile int*)&hs->con);
                if (old == kNoCon) old = atomicCAS(&hs->con, kNoCon, SC.con);
                if (old != kNoCon && old != SC.con && !(hflags & kScConMulti)) atomicOr(&hs->flags, kScConMulti);
              }
            }
          }
        }
        bar.sync();
        VSB_CPASS(2);
        // ---- C4: count the edges the certificates do not cover ----
        {
          unsigned long long mine = 0;
          for (unsigned long long i = seg_lo + tid; i < seg_hi; i += nthr) {
            const uint4 e = reinterpret_cast<const uint4*>(master)[i];
            if (!VSB_IN_SEG(e)) continue;
            const int sa = p.rec[e.y].sz, sb = p.rec[e.z].sz;
            bool cert = false;
            if (no_cert) {
            } else if (sa < mins || sb < mins) {
              const int c = cl_find_compress(p.cl, sa < mins ? (int)e.y : (int)e.z);
              // the verdict is a function of the sub-cluster: the first edges to ask compute it, the rest read it
              const int fl = *((volatile int*)&p.hull[c].flags);
              if (fl & (kScCertYes | kScCertNo)) cert = (fl & kScCertYes) != 0;
              else {
                int target;
                const NodeScratch SCc = load_sc(&p.hull[c]);
                cert = subcluster_certified(p, c, SCc, wt, mins, &target);
                atomicOr(&p.hull[c].flags, cert ? kScCertYes : kScCertNo);
              }
              if (!cert && p.debug) atomicAdd(&p.debug[kNumBuckets * 4 + 8 + subcluster_why(p, load_sc(&p.hull[c]), wt, mins)], 1ull);
            } else if (p.debug) atomicAdd(&p.debug[kNumBuckets * 4 + 8], 1ull);
            if (!cert) ++mine;
            if (!cert && p.has_constraints) {
              // an uncertified same-id edge may reset a constraint id (split); dormant edges of whatever it touches
              // become guards of this segment (C5)
              const int ca = p.rec[e.y].con, cb = p.rec[e.z].con;
              if (ca >= 0 && ca == cb) {
                // the split resets the id of the side smaller than 0.3 x the other one (both if neither is): a hub
                // keeps its id whenever the whole sub-cluster on the other side weighs less than 0.3 x the hub
                NodeScratch* ta = &p.hull[sa >= mins ? (int)e.y : cl_find_compress(p.cl, (int)e.y)];
                NodeScratch* tb = &p.hull[sb >= mins ? (int)e.z : cl_find_compress(p.cl, (int)e.z)];
                bool ma = true, mb = true;
                if (sa >= mins && sb < mins) ma = !((double)tb->mass < (double)sa * 0.29);
                if (sb >= mins && sa < mins) mb = !((double)ta->mass < (double)sb * 0.29);
                if (ma && !(*((volatile int*)&ta->flags) & kScSplitCap)) atomicOr(&ta->flags, kScSplitCap);
                if (mb && !(*((volatile int*)&tb->flags) & kScSplitCap)) atomicOr(&tb->flags, kScSplitCap);
              }
            }
          }
          for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
          if (lane == 0 && mine) atomicAdd(&p.counters[5], mine);
        }
        bar.sync();
        VSB_CPASS(3);
        const unsigned long long n_unc = *((volatile unsigned long long*)&p.counters[5]);
        const bool split = (n_unc > kResidualSplit && seg_hi - seg_lo > kSegmentMin);
        // ---- C5: apply the certified merges (skipped when the segment is halved) ----
        if (!split) {
          for (unsigned long long i = seg_lo + tid; i < seg_hi; i += nthr) {
            const uint4 e = reinterpret_cast<const uint4*>(master)[i];
            const unsigned char dn0 = p.done[e.w];
            if (dn0 == 3) {
              // dormant edge (different constraint ids): a guard of this segment if a split may reach one of its
              // sides -- then the exact paths decide it in order; otherwise it is a certified no-op
              bool guard = false;
              const int con_y = p.rec[e.y].con, con_z = p.rec[e.z].con;
              if (con_y >= 0 && con_z >= 0) {
#pragma unroll
                for (int s2 = 0; s2 < 2; ++s2) {
                  const int r = s2 ? (int)e.z : (int)e.y;
                  if (p.rec[r].sz >= mins) guard = guard || (p.hull[r].flags & kScSplitCap);
                  else {
                    const NodeScratch SCg = load_sc(&p.hull[cl_find_compress(p.cl, r)]);
                    guard = guard || (SCg.flags & kScSplitCap);
                    if (SCg.hub0 >= 0) guard = guard || (p.hull[SCg.hub0].flags & kScSplitCap);
                    if (SCg.hub1 >= 0) guard = guard || (p.hull[SCg.hub1].flags & kScSplitCap);
                    if (SCg.flags & kScHubs3) guard = true;
                  }
                }
              } else {
                // inert pair of hubs: a guard if the unconstrained side(s) may take, in this segment, the id that
                // makes it a same-id pair (what a hub may absorb carries at most one id, or several: unknown)
                const NodeScratch Hy = load_sc(&p.hull[e.y]), Hz = load_sc(&p.hull[e.z]);
                const int cand_y = con_y >= 0 ? con_y : ((Hy.flags & kScConMulti) ? -2 : (Hy.con != kNoCon ? Hy.con : -1));
                const int cand_z = con_z >= 0 ? con_z : ((Hz.flags & kScConMulti) ? -2 : (Hz.con != kNoCon ? Hz.con : -1));
                guard = (cand_y != -1 && cand_z != -1) && (cand_y == -2 || cand_z == -2 || cand_y == cand_z);
              }
              p.done[e.w] = guard ? 0 : 1;
              continue;
            }
            if (dn0) continue;
            const int sa = p.rec[e.y].sz, sb = p.rec[e.z].sz;    // sizes are not folded before the next barrier
            if (sa >= mins && sb >= mins) continue;
            const int c = cl_find_compress(p.cl, sa < mins ? (int)e.y : (int)e.z);
            const NodeScratch SC = load_sc(&p.hull[c]);
            const int target = (SC.hub0 >= 0) ? SC.hub0 : c;        // certified sub-clusters touch at most one hub
            if (no_cert) continue;
            if (!(SC.flags & kScCertYes)) {
              // the ordered rounds may let frozen hubs absorb: publish the certificate
#pragma unroll
              for (int q = 0; q < 2; ++q) {
                const int h = q ? SC.hub1 : SC.hub0;
                if (h < 0) continue;
                if (p.hull[h].frozen != wtag && hub_frozen_eval(load_rec(&p.rec[h]), load_sc(&p.hull[h]), wt)) p.hull[h].frozen = wtag;
              }
              continue;
            }
#pragma unroll
            for (int s2 = 0; s2 < 2; ++s2) {
              const int r = s2 ? (int)e.z : (int)e.y;
              if (r == target) continue;
              const int old = atomicExch(&p.parent[r], target);
              if (old == r) {
                const RegionRec R = load_rec(&p.rec[r]);
                if (R.con >= 0) atomicMax(&p.rec[target].con, R.con);
                acc_add(p.acc, target, R);
              }
            }
            p.done[e.w] = 2;                      // certified (2 = merged in this pass: scratch reset below still sees it)
          }
        }
        bar.sync();
        VSB_CPASS(4);
        // ---- C6: fold the bulk contributions, scratch back to idle ----
        for (unsigned long long i = seg_lo + tid; i < seg_hi; i += nthr) {
          const uint4 e = reinterpret_cast<const uint4*>(master)[i];
          const unsigned char dn = p.done[e.w];
          if (dn == 1 || dn == 3) continue;       // was not part of this attempt / still dormant (the segment is being halved)
#pragma unroll
          for (int s2 = 0; s2 < 2; ++s2) {
            const int r = s2 ? (int)e.z : (int)e.y;
            acc_fold(p, r);
            p.cl[r] = r;
            reset_sc(&p.hull[r]);
          }
          if (dn == 2) p.done[e.w] = 1;
        }
        if (tid == 0) { atomicAdd(&p.stats[5], 1ull); if (!split) atomicAdd(&p.stats[6], n_unc); }
        bar.sync();
        VSB_CPASS(5);
        if (!split) { unc_sum += n_unc; break; }
        any_split = true;
        {
          // shrink in proportion to the excess (at least by half): uncertified edges cluster around a few hubs
          const unsigned long long len = seg_hi - seg_lo;
          unsigned long long nl = len * kResidualSplit / n_unc;
          nl = min(nl, (len + 1) / 2);
          nl = max(nl, kSegmentMin);
          seg_hi = seg_lo + nl;
          seg_len = nl;
        }
      }
#undef VSB_IN_SEG
      VSB_PHASE(1);                                // certification attempts
      // ---------------- ordered rounds on what is left of the segment ----------------
      if (have_live) {
        RoundState st;
        st.epoch = epoch + 1; st.buf = 0; st.from_master = true; st.n_src = n_master; st.prev_live = ~0ull >> 1;
        if (tid == 0) { p.counters[0] = 0ull; p.counters[1] = 0ull; }
        bar.sync();
        // big lists: grid-wide rounds; small lists: block 0 alone behind __syncthreads (a round then costs
        // a few microseconds instead of three grid barriers); dependency chains: serial window mode
        int status = ordered_rounds<Bar>(p, bar, tid, nthr, b, edge_w, wtag, master, seg_lo, seg_hi, st,
                                         (kIsGrid && !(p.dev_flags & 4)) ? kBlockRoundsLimit : 0ull, kSerialSwitch, guard);
        epoch = st.epoch;
        VSB_PHASE(2);                              // ordered rounds
        if (status == 3) return;                 // watchdog
        if (status == 1 || status == 2) {
          // ordered list of the segment's pending positions (stable compaction of the done flags by the
          // whole grid: every block takes a contiguous slice), written to the list buffer not in use
          uint32_t* pend_list = st.buf ? p.live_b : p.live_c;
          const unsigned nblk = kIsGrid ? gridDim.x : 1u, blk = kIsGrid ? blockIdx.x : 0u;
          const unsigned long long range = seg_hi - seg_lo;
          const unsigned long long slice = (range + nblk - 1) / nblk;
          const unsigned long long s_lo = seg_lo + min(range, (unsigned long long)blk * slice), s_hi = min(seg_hi, s_lo + slice);
          unsigned long long* blockcnt = p.counters + 16;
          {
            unsigned cnt = 0;
            for (unsigned long long i = s_lo + threadIdx.x; i < s_hi; i += blockDim.x) cnt += (p.done[master[4 * i + 3]] == 0) ? 1u : 0u;
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
            if ((threadIdx.x & 31) == 0) S.warp_cnt[threadIdx.x >> 5] = cnt;
            __syncthreads();
            if (threadIdx.x == 0) {
              unsigned long long t = 0;
              for (int k = 0; k < kMergeWarps; ++k) t += S.warp_cnt[k];
              blockcnt[blk] = t;
            }
            __syncthreads();
          }
          bar.sync();
          unsigned long long offset = 0, n_pend = 0;
          for (unsigned k = lane; k < nblk; k += 32) { const unsigned long long c = *((volatile unsigned long long*)&blockcnt[k]); if (k < blk) offset += c; n_pend += c; }
          for (int o = 16; o > 0; o >>= 1) { offset += __shfl_xor_sync(0xffffffffu, offset, o); n_pend += __shfl_xor_sync(0xffffffffu, n_pend, o); }
          for (unsigned long long base = s_lo; base < s_hi; base += blockDim.x) {
            const unsigned long long i = base + threadIdx.x;
            uint32_t pos = 0;
            bool flag = false;
            if (i < s_hi) { pos = master[4 * i + 3]; flag = (p.done[pos] == 0); }
            unsigned total;
            const unsigned rank = block_rank(S, flag, &total);
            if (flag) pend_list[offset + rank] = pos;
            offset += total;
          }
          bar.sync();
          VSB_PHASE(3);                            // ordered compaction of the pending positions
          bool serial_tail = true;
          if (kIsGrid && n_pend > kGroupScanMin && !(p.dev_flags & 16)) {
            // ---- group-parallel exact scans: regions joined by pending edges (not through commuting hubs)
            // form independent groups; every CTA scans the groups hashed to it, in reference order ----
            uint32_t* const grp = st.buf ? p.live_c : p.live_b;           // the list buffer the compaction did not use
            uint2* const roots = reinterpret_cast<uint2*>(grp + n_pend + (n_pend & 1ull));
            if (tid == 0) p.counters[7] = 0ull;
            for (unsigned long long i = tid; i < n_pend; i += nthr) {
              int u, v;
              decode_edge(p, codes[pend_list[i]], u, v);
              const int ru = uf_find(p.parent, u), rv = uf_find(p.parent, v);
              roots[i] = make_uint2((unsigned)ru, (unsigned)rv);
              if (ru != rv && !commuting_hub(p, ru, wtag) && !commuting_hub(p, rv, wtag)) cl_union(p.cl, ru, rv);
            }
            bar.sync();
            for (unsigned long long i = tid; i < n_pend; i += nthr) {
              const uint2 r = roots[i];
              const bool cu = commuting_hub(p, (int)r.x, wtag), cv = commuting_hub(p, (int)r.y, wtag);
              // an edge between two commuting hubs has no group: it is left to the serial tail
              grp[i] = (cu && cv) ? 0xFFFFFFFFu : (uint32_t)cl_find_compress(p.cl, cu ? (int)r.y : (int)r.x);
              if (cu && cv && r.x != r.y) p.counters[7] = 1ull;
            }
            bar.sync();
            {
              ScanShared& C = S.scan;
              uint32_t* const q_pos = p.scan_queue + (unsigned long long)blockIdx.x * 2ull * kScanMax;   // per-CTA queue of (position, group)
              uint32_t* const q_grp = q_pos + kScanMax;
              const int deftag = -wtag - 1;
              int q_n = 0;
              for (unsigned long long base = 0; base < n_pend; base += blockDim.x) {
                const unsigned long long i = base + threadIdx.x;
                uint32_t g = 0xFFFFFFFFu;
                bool mine = false;
                if (i < n_pend) {
                  g = grp[i];
                  mine = g != 0xFFFFFFFFu && (g * 2654435761u >> 8) % gridDim.x == blockIdx.x && p.hull[g].claim != deftag;
                }
                unsigned total;
                const unsigned rank = block_rank(S, mine, &total);
                if (mine) { q_pos[q_n + rank] = pend_list[i]; q_grp[q_n + rank] = g; }
                q_n += (int)total;
                __syncthreads();
                if (q_n + (int)blockDim.x > kScanMax) {
                  exact_scan(p, C, b, codes, q_pos, q_n, p.done, wtag, q_grp, &p.counters[7]);
                  q_n = 0;
                }
              }
              if (q_n) exact_scan(p, C, b, codes, q_pos, q_n, p.done, wtag, q_grp, &p.counters[7]);
            }
            bar.sync();
            // fold what the commuting hubs absorbed, group scratch back to idle
            for (unsigned long long i = tid; i < n_pend; i += nthr) {
              const uint2 r = roots[i];
              acc_fold(p, (int)r.x); acc_fold(p, (int)r.y);
              p.cl[r.x] = (int)r.x; p.cl[r.y] = (int)r.y;
            }
            bar.sync();
            serial_tail = *((volatile unsigned long long*)&p.counters[7]) != 0ull;
            if (serial_tail) {
              // deferred groups / hub-hub edges: what is still pending runs through ONE exact scan, in order
              if (!kIsGrid || blockIdx.x == 0) {
                unsigned long long m = 0;
                for (unsigned long long base = 0; base < n_pend; base += blockDim.x) {
                  const unsigned long long i = base + threadIdx.x;
                  const bool flag = (i < n_pend) && (p.done[pend_list[i]] == 0);
                  unsigned total;
                  const unsigned rank = block_rank(S, flag, &total);
                  if (flag) grp[m + rank] = pend_list[i];
                  m += total;
                  __syncthreads();
                }
                for (unsigned long long q0 = 0; q0 < m; q0 += kScanMax)
                  { const int nq = (int)min((unsigned long long)kScanMax, m - q0); if (p.dev_flags & 128) exact_scan(p, S.scan, b, codes, grp + q0, nq, p.done, -1, nullptr, nullptr); else split_scan(p, S, b, codes, grp + q0, nq, p.done); }
              }
              bar.sync();
            }
            serial_tail = false;
          }
          if (serial_tail && (!kIsGrid || blockIdx.x == 0)) {
            // the pending edges in reference order, kScanMax at a time (each batch reloads the current roots)
            for (unsigned long long q0 = 0; q0 < n_pend; q0 += kScanMax)
              { const int nq = (int)min((unsigned long long)kScanMax, n_pend - q0); if (p.dev_flags & 128) exact_scan(p, S.scan, b, codes, pend_list + q0, nq, p.done, -1, nullptr, nullptr); else split_scan(p, S, b, codes, pend_list + q0, nq, p.done); }
          }
          bar.sync();
          if (p.debug && tid == 0 && blockIdx.x == 0) {
            if (n_pend <= (unsigned long long)kScanMax) { p.debug[kNumBuckets * 4 + 27] += n_pend; } else { p.debug[kNumBuckets * 4 + 28] += 1ull; p.debug[kNumBuckets * 4 + 29] += n_pend; }
          }
          if (n_pend <= (unsigned long long)kScanMax) { VSB_PHASE(5); } else { VSB_PHASE(6); }   // exact scan / serial window mode
        }
      }
      // ---- the hub-hub edge that ended the segment runs alone, exactly (a real big-big decision) ----
      const bool at_hubhub = (seg_hi == hh && hh < n_master);
      if (at_hubhub && tid == 0) {
        const uint4 e = reinterpret_cast<const uint4*>(master)[hh];
        if (!p.done[e.w]) {
          const int ru = uf_find(p.parent, (int)e.y), rv = uf_find(p.parent, (int)e.z);
          if (ru != rv) exec_strict(p, ru, rv, edge_w, p.stats);
          p.done[e.w] = 1;
        }
      }
      seg_lo = at_hubhub ? hh + 1 : seg_hi;
      if (seg_lo >= n_master) break;
      if (seg_hi == seg_end) seg_len = min(seg_len * 2, n_master);     // the whole planned segment went through: grow again
      // ---- refresh: roots of the entries still ahead, drop what became inert, next hub-hub edge ----
      if (tid == 0) p.counters[4] = ~0ull;
      ++epoch;
      bar.sync();
      for (unsigned long long i = seg_lo + tid; i < n_master; i += nthr) {
        uint4 e = reinterpret_cast<const uint4*>(master)[i];
        const unsigned char dn = p.done[e.w];
        if (dn == 1) continue;
        const int ru = uf_find(p.parent, (int)e.y), rv = uf_find(p.parent, (int)e.z);
        bool drop = (ru == rv);
        bool hubhub = false, dormant = false;
        if (!drop) {
          const RegionRec A = load_rec(&p.rec[ru]), B = load_rec(&p.rec[rv]);
          const bool both_con = (A.con >= 0 && B.con >= 0);
          const bool inert = !both_con && (A.fin || B.fin) && A.sz >= mins && B.sz >= mins;
          dormant = (both_con && A.con != B.con) || (inert && p.has_constraints);
          drop = inert && !dormant;
          hubhub = !drop && !dormant && A.sz >= mins && B.sz >= mins;
        }
        if (drop) { p.done[e.w] = 1; continue; }
        if (dormant) {
          if (dn != 3) p.done[e.w] = 3;
          if (ru != (int)e.y || rv != (int)e.z) { e.y = (uint32_t)ru; e.z = (uint32_t)rv; reinterpret_cast<uint4*>(master)[i] = e; }
          continue;
        }
        if (dn == 3) p.done[e.w] = 0;          // a dormant edge woke up (one side lost its constraint id)
        if (hubhub) atomicMin(&p.counters[4], i);
        if (ru != (int)e.y || rv != (int)e.z) { e.y = (uint32_t)ru; e.z = (uint32_t)rv; reinterpret_cast<uint4*>(master)[i] = e; }
      }
      bar.sync();
      hh = *((volatile unsigned long long*)&p.counters[4]);
      VSB_PHASE(4);                                // hub-hub edge + refresh
    }
    // next window: aim at kWindowTarget live edges
    w0 += n_edges;
    {
      if (any_split) target = max(target / 2, kWindowTarget / 4);
      else if (unc_sum * 64 <= n_live0 && !(p.dev_flags & 2)) target = min(target * 2, kWindowTarget * 32);
      const unsigned long long live0 = n_live0 ? n_live0 : 1;
      unsigned long long next = raw;
      if (live0 * 2 < target) next = raw * 2;
      if (live0 * 8 < target) next = raw * 4;
      if (live0 > target * 2) next = raw / 2;
      raw = max(next, kWindowMin);
      if (tiny_windows) { raw = 4096; target = 4096; }
    }
    ++epoch;
    bar.sync();
  }
  if (tid == 0) {
    p.counters[0] = 0ull;
    p.counters[1] = 0ull;
    p.counters[2] = (unsigned long long)(epoch + 1);
    p.counters[6] = (unsigned long long)(wtag + 1);
  }
}

__global__ void __launch_bounds__(kMergeThreads) merge_kernel(MergeParams p) {
  extern __shared__ __align__(16) unsigned char merge_smem[];
  MergeShared& S = *reinterpret_cast<MergeShared*>(merge_smem);
  GridBar gbar{cg::this_grid()};
  BlockBar bbar;
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned gn = gridDim.x * blockDim.x;
  for (int b = 0; b < kNumBuckets; ++b) {
    const unsigned long long s0 = p.bucket_start[b], s1 = p.bucket_start[b + 1];
    if (s1 == s0) continue;
    unsigned long long t_bucket = 0;
    if (p.debug && gtid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_bucket));
    if (s1 - s0 <= kTailEdges) {
      if (blockIdx.x == 0) run_bucket<BlockBar, false>(p, bbar, S, threadIdx.x, blockDim.x, b, p.codes + s0, s1 - s0);
    } else {
      run_bucket<GridBar, true>(p, gbar, S, gtid, gn, b, p.codes + s0, s1 - s0);
    }
    gbar.sync();
    if (p.debug && gtid == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      p.debug[b * 4 + 0] = t1 - t_bucket;
      p.debug[b * 4 + 1] = p.stats[0];
      p.debug[b * 4 + 2] = s1 - s0;
      p.debug[b * 4 + 3] = p.stats[5];
    }
  }
}

__global__ void init_nodes_kernel(const float* __restrict__ frame, const int* __restrict__ con_ids, int base,
                                  int n, int* __restrict__ parent, RegionRec* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RegionRec r;
  r.sz = 1;                                             // AddNodesWithDescriptors (dense_segmentation_graph.h:1180-1199)
  r.con = con_ids ? con_ids[i] : -1;                    // AddNodesConstrainedWithDescriptors (:1201-1228)
  r.d0 = frame[(size_t)i * 3];
  r.d1 = frame[(size_t)i * 3 + 1];
  r.d2 = frame[(size_t)i * 3 + 2];
  r.fin = 0; r.pad0 = 0; r.pad1 = 0;
  parent[base + i] = base + i;
  store_rec(&rec[base + i], r);
}

int launch_init_nodes(const float* frame, const int* constraint_ids, int slot, int w, int h, int* parent,
                      RegionRec* rec, cudaStream_t s) {
  const int n = w * h;
  init_nodes_kernel<<<(n + 255) / 256, 256, 0, s>>>(frame, constraint_ids, slot * n, n, parent, rec);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

// AddVirtualNodesConstrained (dense_segmentation_graph.h:327-367): size-0 nodes, pre-merged per
// constraint id; representative = first pixel (raster order) carrying the id.
__global__ void virtual_first_kernel(const int* __restrict__ ids, int base, int n, int* __restrict__ first_of_id) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) atomicMin(&first_of_id[ids[i]], base + i);
}
__global__ void virtual_nodes_kernel(const int* __restrict__ ids, int base, int n, const int* __restrict__ first_of_id,
                                     int* __restrict__ parent, RegionRec* __restrict__ rec) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  RegionRec r;
  r.sz = 0; r.con = ids[i]; r.d0 = r.d1 = r.d2 = 0.f; r.fin = 0; r.pad0 = r.pad1 = 0;
  parent[base + i] = first_of_id[ids[i]];
  store_rec(&rec[base + i], r);
}

int launch_init_virtual_nodes(const int* constraint_ids, int slot, int w, int h, int* parent, RegionRec* rec,
                              int* first_of_id, int max_id, cudaStream_t s) {
  const int n = w * h;
  VSB_CUDA_OK(cudaMemsetAsync(first_of_id, 0x7f, sizeof(int) * (size_t)(max_id + 1), s));
  virtual_first_kernel<<<(n + 255) / 256, 256, 0, s>>>(constraint_ids, slot * n, n, first_of_id);
  virtual_nodes_kernel<<<(n + 255) / 256, 256, 0, s>>>(constraint_ids, slot * n, n, first_of_id, parent, rec);
  VSB_CUDA_OK(cudaGetLastError());
  return 0;
}

int launch_merge(const MergeParams& p_in, cudaStream_t s) {
  MergeParams p = p_in;
  // Default 17 = hub-pair certificates off (1) and group-parallel scans off (16): with both on, a 1080p chunk
  // came out at IoU 0.93 against the oracle (tools/gpu_debug_1080p.py; either switch alone restores the exact
  // partition), so the two stay development features (VSB200_MERGE_FLAGS=0) until the certificate is proven.
  p.dev_flags = getenv("VSB200_MERGE_FLAGS") ? atoi(getenv("VSB200_MERGE_FLAGS")) : 17;
  p.window_target = getenv("VSB200_WINDOW_TARGET") ? strtoull(getenv("VSB200_WINDOW_TARGET"), nullptr, 10) : kWindowTargetDefault;
  p.residual_split = getenv("VSB200_RESIDUAL_SPLIT") ? strtoull(getenv("VSB200_RESIDUAL_SPLIT"), nullptr, 10) : kResidualSplitDefault;
  p.segment_min = getenv("VSB200_SEGMENT_MIN") ? strtoull(getenv("VSB200_SEGMENT_MIN"), nullptr, 10) : kSegmentMinDefault;
  {
    // images of the two distance gates on the squared distance y (dist = sqrtf(y), correctly rounded on host and device):
    // the smallest float y whose sqrtf fails the gate, found by bisection over the bit patterns of [0, 1]
    auto gate = [](int which) {
      auto pass = [which](float y) {
        const float d = sqrtf(y);
        return which == 1 ? ((double)d < 0.2) : which == 2 ? !(d > 0.15f) : (d < 0.05f);
      };
      uint32_t lo = 0u, hi = 0x3f800000u;      // pass(0) holds, pass(1) fails
      while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        float y; memcpy(&y, &mid, 4);
        if (pass(y)) lo = mid; else hi = mid;
      }
      float y; memcpy(&y, &hi, 4);
      return y;
    };
    p.y_merge = gate(0);
    p.y_force = gate(1);
    p.y_split = gate(2);
  }
  int dev = 0, sms = 0, per_sm = 0;
  VSB_CUDA_OK(cudaGetDevice(&dev));
  VSB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = sizeof(MergeShared);
  VSB_CUDA_OK(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  VSB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_kernel, kMergeThreads, smem));
  if (per_sm < 1) { set_error("merge kernel does not fit on an SM"); return 3; }
  per_sm = per_sm > 4 ? 4 : per_sm;
  MergeParams pp = p;
  void* args[] = {&pp};
  VSB_CUDA_OK(cudaLaunchCooperativeKernel((void*)merge_kernel, dim3(sms * per_sm), dim3(kMergeThreads), args, smem, s));
  return 0;
}

}  // namespace vsb
