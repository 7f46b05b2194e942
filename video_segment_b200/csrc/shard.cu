// shard.cu -- frame-group sharding of one video over the GPUs of a box, C ABI vsb200_shard_*.
// The reference chain is sequential across chunks: chunk k+1 is constrained by the region ids chunk k gave its two
// overlap frames (overlap_segmentations_, dense_segmentation.cpp:300-328) and numbers new regions after
// max_region_id_ (:360-365).  Sharded, rank g runs the chain of frame group g on its own GPU; at a group boundary
//   C1  the two overlap frames' region-id maps of group g go to rank g+1          ncclSend / ncclRecv (NVLink)
//   C2  the groups' region-id counts are all-gathered -> exclusive prefix           ncclAllGather
// on a side stream of the handle, and the successor makes its ids consistent with the predecessor's on the device:
// every id it uses in the shared frame is mapped to the predecessor id covering most of its pixels (vote kernels:
// pair counting in a hash table, arg-max per id), ids born later get `prefix + id`; the table is applied to id
// images on the device (relabel kernel) and to the per-region id arrays of the results on the host.
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy a host process already loaded, e.g. PyTorch's, else the
// system one), so libvsb200.so itself keeps depending on libcudart only.
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <memory>
#include <vector>

#include "../../include/vsb200.h"
#include "common.cuh"

using namespace vsb;

namespace {

// the slice of nccl.h this file needs (NCCL 2.x ABI)
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
typedef int ncclDataType_t;
constexpr ncclDataType_t kNcclInt32 = 2, kNcclInt64 = 4;

struct Nccl {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

Nccl* nccl() {
  static Nccl n;
  static bool tried = false;
  if (tried) return n.so ? &n : nullptr;
  tried = true;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char* nm : names) { n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (n.so) break; }
  if (!n.so) { set_error("NCCL not found: %s", dlerror()); return nullptr; }
#define VSB_SYM(field, name) *(void**)(&n.field) = dlsym(n.so, name); if (!n.field) { set_error("NCCL symbol %s missing", name); n.so = nullptr; return nullptr; }
  VSB_SYM(GetUniqueId, "ncclGetUniqueId") VSB_SYM(CommInitRank, "ncclCommInitRank") VSB_SYM(CommDestroy, "ncclCommDestroy")
  VSB_SYM(Send, "ncclSend") VSB_SYM(Recv, "ncclRecv") VSB_SYM(AllGather, "ncclAllGather") VSB_SYM(GroupStart, "ncclGroupStart")
  VSB_SYM(GroupEnd, "ncclGroupEnd") VSB_SYM(GetErrorString, "ncclGetErrorString") VSB_SYM(GetVersion, "ncclGetVersion")
#undef VSB_SYM
  return &n;
}

#define SH_CUDA(expr)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VSB200_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)
#define SH_NCCL(expr)                                                                       \
  do {                                                                                      \
    ncclResult_t _r = (expr);                                                               \
    if (_r != 0) {                                                                          \
      set_error("%s failed: %s (%s:%d)", #expr, N->GetErrorString(_r), __FILE__, __LINE__);  \
      return VSB200_ERR_CUDA;                                                               \
    }                                                                                       \
  } while (0)

// ---- seam vote: (successor id, predecessor id) -> pixels, then the best predecessor per successor id ----
constexpr unsigned long long kEmpty = ~0ull;

__global__ void vote_count_kernel(const int* __restrict__ succ, const int* __restrict__ pred, size_t n, unsigned long long* __restrict__ keys,
                                  unsigned* __restrict__ counts, unsigned cap_mask) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int s = succ[i], p = pred[i];
    if (s < 0 || p < 0) continue;
    const unsigned long long key = ((unsigned long long)(unsigned)s << 32) | (unsigned)p;
    // lanes of a warp mostly look at the same pair: one table update per distinct pair and warp
    const unsigned peers = __match_any_sync(__activemask(), key);
    if ((threadIdx.x & 31u) != (unsigned)(__ffs(peers) - 1)) continue;
    unsigned h = (unsigned)((key * 0x9E3779B97F4A7C15ull) >> 32);
    for (unsigned probe = 0; probe <= cap_mask; ++probe) {
      const unsigned slot = (h + probe) & cap_mask;
      unsigned long long cur = keys[slot];
      if (cur == kEmpty) { const unsigned long long old = atomicCAS(&keys[slot], kEmpty, key); cur = (old == kEmpty) ? key : old; }
      if (cur == key) { atomicAdd(&counts[slot], (unsigned)__popc(peers)); break; }
    }
  }
}

// best[s] = max over predecessors of (pixels << 32 | ~pred): most pixels, ties to the smaller predecessor id
__global__ void vote_best_kernel(const unsigned long long* __restrict__ keys, const unsigned* __restrict__ counts, unsigned cap,
                                 unsigned long long* __restrict__ best, int n_ids) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cap) return;
  const unsigned long long key = keys[i];
  if (key == kEmpty) return;
  const int s = (int)(key >> 32);
  if (s >= n_ids) return;
  const unsigned p = (unsigned)(key & 0xffffffffu);
  atomicMax(&best[s], ((unsigned long long)counts[i] << 32) | (unsigned long long)(0xffffffffu - p));
}

__global__ void vote_table_kernel(const unsigned long long* __restrict__ best, int n_ids, long long id_offset, int* __restrict__ table) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_ids) return;
  const unsigned long long b = best[s];
  table[s] = b ? (int)(0xffffffffu - (unsigned)(b & 0xffffffffu)) : (int)(id_offset + s);
}

__global__ void relabel_ids_kernel(int* __restrict__ ids, size_t n, const int* __restrict__ table, int n_ids) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int v = ids[i];
    if (v >= 0 && v < n_ids) ids[i] = table[v];
  }
}

}  // namespace

struct vsb200_shard {
  int rank = 0, world = 1, device = 0, w = 0, h = 0;
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  int* d_out = nullptr;            // [2][h][w] this group's overlap id maps (C1 payload)
  int* d_in = nullptr;             // [2][h][w] the predecessor's
  long long* d_counts = nullptr;   // [world + 1]
  unsigned long long* d_keys = nullptr; unsigned* d_cnt = nullptr; unsigned vote_cap = 1u << 20;
  unsigned long long* d_best = nullptr; int best_cap = 0;
  bool have_pred = false;
  double exchange_ms = 0;
  long long exchanges = 0, launches = 0;
  ~vsb200_shard() {
    Nccl* N = nccl();
    if (comm && N) N->CommDestroy(comm);
    auto F = [](void* p) { if (p) cudaFree(p); };
    F(d_out); F(d_in); F(d_counts); F(d_keys); F(d_cnt); F(d_best);
    if (stream) cudaStreamDestroy(stream);
  }
};

extern "C" {

int vsb200_shard_unique_id(uint8_t id_out[128]) {
  Nccl* N = nccl();
  if (!N || !id_out) return N ? VSB200_ERR_INVALID : VSB200_ERR_UNSUPPORTED;
  ncclUniqueId id;
  SH_NCCL(N->GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return VSB200_OK;
}

int vsb200_shard_create(const uint8_t id[128], int rank, int world, int device, int width, int height, vsb200_shard** out) {
  if (!out || world < 1 || rank < 0 || rank >= world || width < 2 || height < 2 || (world > 1 && !id)) { set_error("shard_create: bad arguments"); return VSB200_ERR_INVALID; }
  if (vsb200_device_count() <= 0) { set_error("no sm_100 CUDA device available: this path has no CPU fallback"); return VSB200_ERR_NO_DEVICE; }
  std::unique_ptr<vsb200_shard> s(new vsb200_shard);
  s->rank = rank; s->world = world; s->device = device; s->w = width; s->h = height;
  SH_CUDA(cudaSetDevice(device));
  SH_CUDA(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  const size_t npx = (size_t)width * height;
  SH_CUDA(cudaMalloc(&s->d_out, npx * 2 * sizeof(int)));
  SH_CUDA(cudaMalloc(&s->d_in, npx * 2 * sizeof(int)));
  SH_CUDA(cudaMalloc(&s->d_counts, (size_t)(world + 1) * sizeof(long long)));
  SH_CUDA(cudaMalloc(&s->d_keys, (size_t)s->vote_cap * 8));
  SH_CUDA(cudaMalloc(&s->d_cnt, (size_t)s->vote_cap * 4));
  if (world > 1) {
    Nccl* N = nccl();
    if (!N) return VSB200_ERR_UNSUPPORTED;
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    SH_NCCL(N->CommInitRank(&s->comm, world, uid, rank));
  }
  *out = s.release();
  return VSB200_OK;
}

// C1 + C2 at a group boundary.  `d` must stand right after a chunk boundary (vsb200_dense_export_halo's condition).
// id_offsets_out: [world + 1] exclusive prefix of the groups' region-id counts (last entry = total); *have_pred = the
// predecessor's maps arrived.
int vsb200_shard_exchange(vsb200_shard* s, vsb200_dense* d, int64_t* id_offsets_out, int* have_pred) {
  if (!s || !d || !id_offsets_out) { set_error("shard_exchange: bad arguments"); return VSB200_ERR_INVALID; }
  SH_CUDA(cudaSetDevice(s->device));
  const size_t npx = (size_t)s->w * s->h;
  int32_t chain[3];
  if (int rc = vsb200_dense_export_halo(d, s->d_out, s->d_out + npx, chain)) return rc;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, s->stream);
  const long long mine = chain[0];
  std::vector<long long> counts(s->world, 0);
  if (s->world > 1) {
    Nccl* N = nccl();
    SH_CUDA(cudaMemcpyAsync(s->d_counts + s->world, &mine, sizeof(long long), cudaMemcpyHostToDevice, s->stream));
    SH_NCCL(N->GroupStart());
    if (s->rank + 1 < s->world) SH_NCCL(N->Send(s->d_out, npx * 2, kNcclInt32, s->rank + 1, s->comm, s->stream));
    if (s->rank > 0) SH_NCCL(N->Recv(s->d_in, npx * 2, kNcclInt32, s->rank - 1, s->comm, s->stream));
    SH_NCCL(N->GroupEnd());
    SH_NCCL(N->AllGather(s->d_counts + s->world, s->d_counts, 1, kNcclInt64, s->comm, s->stream));
    SH_CUDA(cudaMemcpyAsync(counts.data(), s->d_counts, sizeof(long long) * s->world, cudaMemcpyDeviceToHost, s->stream));
  } else {
    counts[0] = mine;
  }
  cudaEventRecord(e1, s->stream);
  SH_CUDA(cudaStreamSynchronize(s->stream));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  s->exchange_ms += ms;
  ++s->exchanges;
  long long acc = 0;
  for (int g = 0; g < s->world; ++g) { id_offsets_out[g] = acc; acc += counts[g]; }
  id_offsets_out[s->world] = acc;
  s->have_pred = s->rank > 0;
  if (have_pred) *have_pred = s->have_pred ? 1 : 0;
  return VSB200_OK;
}

const int32_t* vsb200_shard_pred_maps(vsb200_shard* s) { return (s && s->have_pred) ? s->d_in : nullptr; }

// Seam vote on the device: this group's own id map of the shared frame (its first output frame) against the
// predecessor's map of the same frame (the second of the received maps) -> dev_table_out[n_ids]: local id -> global id.
int vsb200_shard_vote(vsb200_shard* s, const int32_t* dev_own_first_map, int n_ids, int64_t id_offset, int32_t* dev_table_out) {
  if (!s || !dev_table_out || n_ids < 1) { set_error("shard_vote: bad arguments"); return VSB200_ERR_INVALID; }
  SH_CUDA(cudaSetDevice(s->device));
  if (n_ids > s->best_cap) {
    if (s->d_best) cudaFree(s->d_best);
    s->best_cap = n_ids * 2;
    SH_CUDA(cudaMalloc(&s->d_best, (size_t)s->best_cap * 8));
  }
  SH_CUDA(cudaMemsetAsync(s->d_best, 0, (size_t)n_ids * 8, s->stream));
  if (s->have_pred && dev_own_first_map) {
    const size_t npx = (size_t)s->w * s->h;
    SH_CUDA(cudaMemsetAsync(s->d_keys, 0xff, (size_t)s->vote_cap * 8, s->stream));
    SH_CUDA(cudaMemsetAsync(s->d_cnt, 0, (size_t)s->vote_cap * 4, s->stream));
    vote_count_kernel<<<148 * 4, 256, 0, s->stream>>>(dev_own_first_map, s->d_in + npx, npx, s->d_keys, s->d_cnt, s->vote_cap - 1);
    vote_best_kernel<<<(s->vote_cap + 255) / 256, 256, 0, s->stream>>>(s->d_keys, s->d_cnt, s->vote_cap, s->d_best, n_ids);
    s->launches += 2;
  }
  vote_table_kernel<<<(n_ids + 255) / 256, 256, 0, s->stream>>>(s->d_best, n_ids, (long long)id_offset, dev_table_out);
  ++s->launches;
  SH_CUDA(cudaGetLastError());
  SH_CUDA(cudaStreamSynchronize(s->stream));
  return VSB200_OK;
}

// table applied to an int32 id buffer on the device (id images; ids outside [0, n_ids) are left alone)
int vsb200_shard_relabel(vsb200_shard* s, int32_t* dev_ids, size_t n, const int32_t* dev_table, int n_ids) {
  if (!s || !dev_ids || !dev_table) { set_error("shard_relabel: bad arguments"); return VSB200_ERR_INVALID; }
  SH_CUDA(cudaSetDevice(s->device));
  relabel_ids_kernel<<<148 * 4, 256, 0, s->stream>>>(dev_ids, n, dev_table, n_ids);
  ++s->launches;
  SH_CUDA(cudaGetLastError());
  SH_CUDA(cudaStreamSynchronize(s->stream));
  return VSB200_OK;
}

void vsb200_shard_stats(vsb200_shard* s, double out[3]) {
  if (!s || !out) return;
  out[0] = s->exchange_ms; out[1] = (double)s->exchanges; out[2] = (double)s->launches;
}

void vsb200_shard_destroy(vsb200_shard* s) { delete s; }

}  // extern "C"
