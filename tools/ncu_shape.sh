#!/bin/bash
# Dev tool (GPU box, under gpurun): the result stage after the merge -- launch list (durations) of its kernels over a
# free and a constrained 1080p chunk, then one `ncu --set full` capture of each kernel of csrc/shape.cu.
set -u
mkdir -p gpurun_out/ncu
NCU="ncu --clock-control none"
export PROFILE_REGION=0
$NCU --metrics gpu__time_duration.sum -k 'regex:run_cc|radix|group_|result_keys|relabel_groups|scan_u32|rle_|n4_|flatten|neighbor|gather_' \
    --csv --log-file gpurun_out/ncu/r02_shape_launches.csv python tools/profile_workload.py 1920 1080 22 > gpurun_out/ncu/shape_launches.log 2>&1
for k in run_cc_link_kernel run_cc_name_kernel radix_hist_kernel radix_scatter_kernel group_heads_count_kernel group_heads_write_kernel group_runs_kernel result_keys_kernel; do
  timeout 300 $NCU --set full --import-source on -k regex:^$k -c 1 -f -o gpurun_out/ncu/r02_$k \
      python tools/profile_workload.py 1920 1080 22 > gpurun_out/ncu/$k.log 2>&1
done
ls -la gpurun_out/ncu | tail -12
