// results.cuh -- launch wrappers of results.cu (label volume -> runs, neighbour pairs).
#pragma once
#include "common.cuh"

namespace vsb {

struct RunRec { int slice, y, left_x, right_x, id; };

int launch_n4(int* labels, int w, int h, int n_slices, const int* dev_slice_ids, int* size_adjust, cudaStream_t s);
int launch_rle_count(const int* labels, int w, int h, const int* dev_slice_ids, int n_slices, unsigned* row_counts,
                     cudaStream_t s);
int launch_scan_u32(const unsigned* in, unsigned* out_exclusive, unsigned* total, int n, cudaStream_t s);
int launch_rle_write(const int* labels, int w, int h, const int* dev_slice_ids, int n_slices,
                     const unsigned* row_offsets, RunRec* runs, cudaStream_t s);
int launch_gather_region_info(const int* ids, int n, const RegionRec* rec, const int* size_adjust, int2* out,
                              cudaStream_t s);
int launch_relabel(const RunRec* runs, int n, int w, int h, int* node_labels, cudaStream_t s);
int launch_neighbor_pairs(const int* roots, const int* labels, int w, int h, int slots, const float* flows, int virtual_slot0,
                          unsigned long long* table, unsigned table_cap_pow2, unsigned long long* out,
                          unsigned long long* out_count, unsigned long long out_cap, cudaStream_t s);
int launch_fill_i32(int* p, int value, long long n, cudaStream_t s);
int launch_init_hull(NodeScratch* hull, long long n_nodes, cudaStream_t s);
int launch_init_iota(int* p, long long n, cudaStream_t s);

}  // namespace vsb
