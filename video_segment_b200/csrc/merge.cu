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

#include "common.cuh"

namespace cg = cooperative_groups;

namespace vsb {

constexpr uint32_t kDone = 0xFFFFFFFFu;
constexpr unsigned kTailEdges = 4096;     // <= this many pending edges: block 0 finishes the bucket
constexpr int kMergeThreads = 256;
constexpr double kFix = 4294967296.0;     // 2^32 fixed point for descriptor sums

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
  const uint32_t e = code >> 4;
  const int dir = (int)(code & 15u);
  const int list = (int)(e / (uint32_t)n);
  const int pix = (int)(e - (uint32_t)list * (uint32_t)n);
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

// The serial decision tree for one edge (segmentation_graph.h:375-440); caller owns both roots.
__device__ __forceinline__ void exec_strict(const MergeParams& p, int ia, int ib, float edge_w,
                                            unsigned long long* stats) {
  RegionRec A = load_rec(&p.rec[ia]);
  RegionRec B = load_rec(&p.rec[ib]);
  const int mins = p.min_region_size;
  if (A.con < 0 || B.con < 0) {
    if (!A.fin && !B.fin) {
      const float d = desc_dist(A, B, edge_w, p.force_merge_weight);
      if (d < 0.05f) {                       // MergeDistanceThreshold, pixel_distance.h:471
        const int m = merge_regions(p.parent, ia, A, ib, B);
        store_rec(&p.rec[m], m == ia ? A : B);
        return;
      }
      A.fin = 1;
      B.fin = 1;
    }
    if (A.fin || B.fin) {
      if (A.sz < mins || B.sz < mins) {
        const int m = merge_regions(p.parent, ia, A, ib, B);
        store_rec(&p.rec[m], m == ia ? A : B);
        return;
      }
    }
    store_rec(&p.rec[ia], A);
    store_rec(&p.rec[ib], B);
  } else if (A.con == B.con) {
    const float d = desc_dist(A, B, edge_w, p.force_merge_weight);
    if (d > 0.15f) {                          // SplitDistanceThreshold, pixel_distance.h:472
      if ((double)A.sz < (double)B.sz * 0.3) A.con = -1;
      else if ((double)B.sz < (double)A.sz * 0.3) B.con = -1;
      else { A.con = -1; B.con = -1; }
      store_rec(&p.rec[ia], A);
      store_rec(&p.rec[ib], B);
    } else {
      const int m = merge_regions(p.parent, ia, A, ib, B);
      store_rec(&p.rec[m], m == ia ? A : B);
    }
  }
  // different constraint ids: never merge, nothing changes
}

__device__ __forceinline__ void acc_add(unsigned long long* acc, int root, const RegionRec& r) {
  unsigned long long* a = acc + (size_t)root * 4;
  const unsigned long long sz = (unsigned long long)r.sz;
  atomicAdd(&a[1], (unsigned long long)__double2ll_rn((double)r.d0 * kFix) * sz);
  atomicAdd(&a[2], (unsigned long long)__double2ll_rn((double)r.d1 * kFix) * sz);
  atomicAdd(&a[3], (unsigned long long)__double2ll_rn((double)r.d2 * kFix) * sz);
  __threadfence();
  atomicAdd(&a[0], sz);
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

// counters: [0] live count of buffer A, [1] of buffer B, [2] round epoch
// live entry = 4 words: code, ru, rv, cluster root
template <class Bar>
__device__ void run_bucket(const MergeParams& p, Bar& bar, const unsigned tid, const unsigned nthr, const int b,
                           const uint32_t* src_codes, unsigned long long src_n, bool p1_done_in,
                           unsigned buf_in, unsigned epoch_in) {
  const float inv_scale = (float)(1.0 / (double)bucket_scale());   // segmentation_graph.h:348
  const float edge_w = (float)b * inv_scale;
  const bool force_bucket = edge_w < p.force_merge_weight;
  const float safe_thr = (force_bucket ? 0.2f : 0.05f) * 0.999f;
  const int mins = p.min_region_size;
  unsigned buf = buf_in;            // index of the buffer P1 writes to
  unsigned epoch = epoch_in;
  bool from_codes = (src_codes != nullptr);
  bool first_round = from_codes;
  bool p1_done = p1_done_in;
  unsigned long long n_src = src_n;
  while (true) {
    uint32_t* dst = buf ? p.live_b : p.live_a;
    const uint32_t* src = buf ? p.live_a : p.live_b;
    unsigned long long* dst_cnt = &p.counters[buf];
    const unsigned long long key_hi = ((unsigned long long)(0xFFFFFFFFu - epoch)) << 32;
    if (!p1_done) {
      // ---- P1: find roots, drop inert edges, reserve ----
      for (unsigned long long i = tid; i < n_src; i += nthr) {
        const uint32_t code = from_codes ? src_codes[i] : src[i * 4];
        if (code == kDone) continue;
        int u, v;
        decode_edge(p, code, u, v);
        const int ru = uf_find(p.parent, u), rv = uf_find(p.parent, v);
        if (ru == rv) continue;
        const RegionRec A = load_rec(&p.rec[ru]), B = load_rec(&p.rec[rv]);
        const bool both_con = (A.con >= 0 && B.con >= 0);
        if (both_con && A.con != B.con) continue;                       // kept for ever
        if (!both_con && (A.fin || B.fin) && A.sz >= mins && B.sz >= mins) continue;   // inert
        const unsigned long long slot = atomicAdd(dst_cnt, 1ull);
        if (slot < p.live_cap) {
          reinterpret_cast<uint4*>(dst)[slot] = make_uint4(code, (uint32_t)ru, (uint32_t)rv, 0u);
          atomicMin(&p.res[ru], key_hi | code);
          atomicMin(&p.res[rv], key_hi | code);
        }
      }
      bar.sync();
    }
    p1_done = false;
    unsigned long long n_live = *((volatile unsigned long long*)dst_cnt);
    if (n_live > p.live_cap) n_live = p.live_cap;   // cannot happen: cap = largest bucket
    if (n_live == 0) break;
    if (first_round && tid == 0) p.stats[4] = n_live;
    if (first_round) {
      // ---- P2a: clusters of this bucket's pending edges ----
      for (unsigned long long i = tid; i < n_live; i += nthr) {
        const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
        cl_union(p.cl, (int)e.y, (int)e.z);
      }
      bar.sync();
      // ---- P2b: per-cluster colour hull / flags ----
      for (unsigned long long i = tid; i < n_live; i += nthr) {
        const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
        const int c = cl_find(p.cl, (int)e.y);
        reinterpret_cast<uint4*>(dst)[i].w = (uint32_t)c;
        int* hl = p.hull + (size_t)c * 8;
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int r = s ? (int)e.z : (int)e.y;
          const RegionRec R = load_rec(&p.rec[r]);
          atomicMin(&hl[0], __float_as_int(R.d0)); atomicMax(&hl[3], __float_as_int(R.d0));
          atomicMin(&hl[1], __float_as_int(R.d1)); atomicMax(&hl[4], __float_as_int(R.d1));
          atomicMin(&hl[2], __float_as_int(R.d2)); atomicMax(&hl[5], __float_as_int(R.d2));
          if (R.fin) atomicOr(&hl[6], 1);
          if (R.con >= 0) { atomicMin(&hl[7], R.con); atomicOr(&hl[6], 2); }
          else atomicOr(&hl[6], 4);
        }
      }
      bar.sync();
      // ---- P2c: merge safe clusters in bulk ----
      for (unsigned long long i = tid; i < n_live; i += nthr) {
        const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
        const int c = (int)e.w;
        int* hl = p.hull + (size_t)c * 8;
        const int flags = hl[6];
        bool safe = (flags & 1) == 0;
        if (safe && (flags & 2)) {
          // constrained members: all must carry one id -> compare min with max via the records
          // (max is tracked through the cluster root's own constraint below); use min only and
          // verify both endpoints agree.
          const RegionRec A = load_rec(&p.rec[(int)e.y]), B = load_rec(&p.rec[(int)e.z]);
          const int cmin = hl[7];
          if ((A.con >= 0 && A.con != cmin) || (B.con >= 0 && B.con != cmin)) safe = false;
        }
        if (safe) {
          const float dx = __int_as_float(hl[3]) - __int_as_float(hl[0]);
          const float dy = __int_as_float(hl[4]) - __int_as_float(hl[1]);
          const float dz = __int_as_float(hl[5]) - __int_as_float(hl[2]);
          const float diam = sqrtf((dx * dx + dy * dy + dz * dz) * (1.0f / 3.0f));
          safe = diam < safe_thr;
        }
        if (!safe) atomicOr(&hl[6], 8);   // one objecting edge makes the whole cluster ordered
      }
      bar.sync();
      for (unsigned long long i = tid; i < n_live; i += nthr) {
        const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
        const int c = (int)e.w;
        const int* hl = p.hull + (size_t)c * 8;
        if (hl[6] & 8) continue;             // unsafe cluster -> ordered rounds
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          const int r = s ? (int)e.z : (int)e.y;
          if (r == c) continue;
          const int old = atomicExch(&p.parent[r], c);
          if (old == r) {
            const RegionRec R = load_rec(&p.rec[r]);
            if (R.con >= 0) atomicMax(&p.rec[c].con, R.con);
            acc_add(p.acc, c, R);
          }
        }
        reinterpret_cast<uint4*>(dst)[i].x = kDone;
      }
      bar.sync();
    }
    // ---- P3: commit (strict owners + absorption by finalised big regions); reset cluster scratch ----
    for (unsigned long long i = tid; i < n_live; i += nthr) {
      const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
      const int ru = (int)e.y, rv = (int)e.z;
      if (first_round) {
        const int c = (int)e.w;
        int* hl = p.hull + (size_t)c * 8;
        reinterpret_cast<int4*>(hl)[0] = make_int4(0x7f7f7f7f, 0x7f7f7f7f, 0x7f7f7f7f, 0);
        reinterpret_cast<int4*>(hl)[1] = make_int4(0, 0, 0, 0x7f7f7f7f);
        p.cl[ru] = ru;
        p.cl[rv] = rv;
      }
      if (e.x == kDone) continue;
      const unsigned long long key = key_hi | e.x;
      const bool own_u = (p.res[ru] == key), own_v = (p.res[rv] == key);
      if (own_u && own_v) {
        exec_strict(p, ru, rv, edge_w, p.stats);
        reinterpret_cast<uint4*>(dst)[i].x = kDone;
        continue;
      }
      if (!own_u && !own_v) continue;
      const RegionRec A = load_rec(&p.rec[ru]), B = load_rec(&p.rec[rv]);
      // x = the side this edge is the next edge of; hub = finalised region of >= min size
      if (own_v && A.fin && A.sz >= mins && B.sz < mins && B.con < 0) {
        p.parent[rv] = ru;
        acc_add(p.acc, ru, B);
        reinterpret_cast<uint4*>(dst)[i].x = kDone;
      } else if (own_u && B.fin && B.sz >= mins && A.sz < mins && A.con < 0) {
        p.parent[ru] = rv;
        acc_add(p.acc, rv, A);
        reinterpret_cast<uint4*>(dst)[i].x = kDone;
      }
    }
    bar.sync();
    // ---- P4: fold bulk contributions ----
    for (unsigned long long i = tid; i < n_live; i += nthr) {
      const uint4 e = reinterpret_cast<const uint4*>(dst)[i];
      acc_fold(p, (int)e.y);
      acc_fold(p, (int)e.z);
      if (first_round) acc_fold(p, (int)e.w);
    }
    if (tid == 0) {
      p.counters[buf ^ 1] = 0ull;          // the other buffer becomes the next destination
      atomicAdd(&p.stats[0], 1ull);
    }
    bar.sync();
    // next round reads what this round wrote
    n_src = n_live;
    from_codes = false;
    first_round = false;
    buf ^= 1;
    ++epoch;
  }
  if (tid == 0) {
    p.counters[0] = 0ull;
    p.counters[1] = 0ull;
    p.counters[2] = (unsigned long long)(epoch + 1);
  }
}

__global__ void __launch_bounds__(kMergeThreads) merge_kernel(MergeParams p) {
  GridBar gbar{cg::this_grid()};
  BlockBar bbar;
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned gn = gridDim.x * blockDim.x;
  for (int b = 0; b < kNumBuckets; ++b) {
    const unsigned long long s0 = p.bucket_start[b], s1 = p.bucket_start[b + 1];
    if (s1 == s0) continue;
    const unsigned epoch = (unsigned)(*((volatile unsigned long long*)&p.counters[2]));
    unsigned long long t_bucket = 0;
    if (p.debug && gtid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_bucket));
    if (s1 - s0 <= kTailEdges) {
      if (blockIdx.x == 0) run_bucket(p, bbar, threadIdx.x, blockDim.x, b, p.codes + s0, s1 - s0, false, 0u, epoch);
    } else {
      run_bucket(p, gbar, gtid, gn, b, p.codes + s0, s1 - s0, false, 0u, epoch);
    }
    gbar.sync();
    if (p.debug && gtid == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      p.debug[b * 4 + 0] = t1 - t_bucket;
      p.debug[b * 4 + 1] = p.stats[0];
      p.debug[b * 4 + 2] = s1 - s0;
      p.debug[b * 4 + 3] = p.stats[4];
    }
  }
}

size_t merge_scratch_bytes(int w, int h, int slots, unsigned long long max_bucket_edges) {
  const size_t n = (size_t)w * h * slots;
  size_t b = 0;
  b += n * sizeof(unsigned long long);        // res
  b += n * 4 * sizeof(unsigned long long);    // acc
  b += n * sizeof(int);                       // cl
  b += n * 8 * sizeof(int);                   // hull
  b += 2 * max_bucket_edges * 16;             // live buffers
  b += 16 * sizeof(unsigned long long);       // counters + stats
  return b + 4096;
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

int launch_merge(const MergeParams& p, cudaStream_t s) {
  int dev = 0, sms = 0, per_sm = 0;
  VSB_CUDA_OK(cudaGetDevice(&dev));
  VSB_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  VSB_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, merge_kernel, kMergeThreads, 0));
  if (per_sm < 1) { set_error("merge kernel does not fit on an SM"); return 3; }
  per_sm = per_sm > 4 ? 4 : per_sm;
  MergeParams pp = p;
  void* args[] = {&pp};
  VSB_CUDA_OK(cudaLaunchCooperativeKernel((void*)merge_kernel, dim3(sms * per_sm), dim3(kMergeThreads), args, 0, s));
  return 0;
}

}  // namespace vsb
