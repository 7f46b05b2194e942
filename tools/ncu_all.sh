#!/bin/bash
# Dev tool (GPU box, under gpurun): launch list of a short bench run plus one `ncu --set full` capture per kernel of the
# path (B200_PROFILING.md).  Reports land in gpurun_out/ncu/; summarise them with tools/ncu_summarise.py.
set -u
mkdir -p gpurun_out/ncu
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 1500 --csv --log-file gpurun_out/ncu/r02_bench_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu/bench_under_ncu.log 2>&1
for k in minmax_u8_kernel build_lut_kernel bilateral_u8_kernel edge_build_tma_pipe_kernel hist_kernel scatter_kernel rowscan_kernel init_nodes_kernel \
         flatten_kernel n4_kernel rle_count_kernel rle_write_kernel neighbor_pairs_rows_kernel mcr_visits_kernel region_hist_kernel flow_hist_kernel \
         finish_slots_kernel slot_distance_kernel merge_slots_kernel paint_runs_kernel; do
  timeout 300 $NCU --set full --import-source on -k regex:^$k -s $( [[ $k == bilateral_u8_kernel || $k == edge_build_tma_pipe_kernel || $k == init_nodes_kernel || $k == region_hist_kernel || $k == flow_hist_kernel || $k == slot_distance_kernel || $k == merge_slots_kernel || $k == paint_runs_kernel ]] && echo 3 || echo 0 ) -c 1 -f -o gpurun_out/ncu/r02_$k \
      python tools/profile_workload.py > gpurun_out/ncu/$k.log 2>&1
done
# the merge kernel: one launch is ~0.1 s on a 640x480 chunk and is replayed ~40 times
PROFILE_REGION=0 timeout 900 $NCU --set full --import-source on -k regex:merge_kernel -c 1 -f -o gpurun_out/ncu/r02_merge_kernel \
    python tools/profile_workload.py 640 480 22 > gpurun_out/ncu/merge_kernel.log 2>&1
ls -la gpurun_out/ncu | tail -40
