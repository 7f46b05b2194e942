#!/bin/bash
# Dev tool (GPU box, under gpurun): re-capture of the two group kernels of csrc/shape.cu after the warp-per-group rewrite,
# plus the launch list of the result stage.
set -u
mkdir -p gpurun_out/ncu
NCU="ncu --clock-control none"
export PROFILE_REGION=0
$NCU --metrics gpu__time_duration.sum -k 'regex:run_cc|radix|group_|result_keys|relabel_groups|scan_u32|rle_|n4_|flatten|neighbor|gather_' \
    --csv --log-file gpurun_out/ncu/r02_shape_launches.csv python tools/profile_workload.py 1920 1080 22 > gpurun_out/ncu/shape_launches.log 2>&1
for k in group_heads_write_kernel group_runs_kernel; do
  timeout 300 $NCU --set full --import-source on -k regex:^$k -c 1 -f -o gpurun_out/ncu/r02_$k \
      python tools/profile_workload.py 1920 1080 22 > gpurun_out/ncu/$k.log 2>&1
done
ls -la gpurun_out/ncu | grep -i "group\|shape"
