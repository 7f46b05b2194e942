set -u
mkdir -p gpurun_out/ncu
NCU="ncu --clock-control none"
for k in region_hist_kernel flow_hist_kernel finish_slots_kernel slot_distance_kernel merge_slots_kernel paint_runs_kernel; do
  timeout 300 $NCU --set full --import-source on -k regex:^$k -s 3 -c 1 -f -o gpurun_out/ncu/r02_$k python tools/profile_workload.py 640 360 4 > gpurun_out/ncu/$k.log 2>&1
done
for k in scatter_kernel hist_kernel; do
  PROFILE_REGION=0 timeout 300 $NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum -k regex:^$k -c 1 python tools/profile_workload.py 2>&1 | grep -E "gpu__time|dram__bytes" | sed "s/^/$k /"
done
ls gpurun_out/ncu/*.ncu-rep | wc -l
python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err; tail -c 2500 gpurun_out/r02_bench_1gpu.json
