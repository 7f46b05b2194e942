// shape.cuh -- launch wrappers of shape.cu (components of the runs, stable pair sort, run groups with shape moments).
#pragma once
#include "results.cuh"

namespace vsb {

// One run of equal keys in a sorted run order: a connected component (pass 0, tag = label) or one region in one frame
// of the result (pass 1, tag = key).  first / count index the interval array written by the same pass.
struct RunGroup { int first, count, tag, slice, area; float mean_x, mean_y, xx, xy, yy; };

int launch_run_components(const RunRec* runs, unsigned n_runs, const unsigned* row_offsets, int h, int slice0, int* parent,
                          unsigned* keys, unsigned* vals, cudaStream_t s);
int launch_sort_pairs(unsigned* keys, unsigned* vals, unsigned* keys_alt, unsigned* vals_alt, unsigned n, int bits,
                      unsigned* hist, unsigned* scratch_total, unsigned** out_keys, unsigned** out_vals, cudaStream_t s);
int launch_group_runs(const unsigned* keys, const unsigned* vals, unsigned n, const RunRec* runs, int pass, unsigned* tile_counts,
                      unsigned* tile_bases, unsigned* head_pos, unsigned* n_groups, RunGroup* groups, int* group_of_run,
                      int3* intervals, cudaStream_t s);
int launch_result_keys(const RunRec* runs, unsigned n_runs, const int* group_of_run, const int* rank_of_group, int slice0,
                       unsigned n_ranks, unsigned* keys, unsigned* vals, cudaStream_t s);
int launch_relabel_groups(const RunRec* runs, unsigned n_runs, const int* group_of_run, const int* label_of_group, int w, int h,
                          int* node_labels, cudaStream_t s);

}  // namespace vsb
