"""Dev tool (CPU, here): reads the .ncu-rep files tools/ncu_all.sh brought back (gpurun_out/ncu/) with `ncu -i ... --page raw
--csv`, writes the per-kernel raw page to profiles/r02_<kernel>_ncu.csv (selected metrics only: the full page is 2 000
columns) and a summary table to profiles/r02_kernels.json.
usage: python tools/ncu_summarise.py [gpurun_out/ncu] [profiles]"""
import csv, glob, io, json, os, subprocess, sys

src = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/ncu"
dst = sys.argv[2] if len(sys.argv) > 2 else "profiles"
KEEP = ["Kernel Name", "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio"]
summary = {}
for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
    name = os.path.basename(rep)[:-len(".ncu-rep")]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        print("skip", rep, "(empty)")
        continue
    head, units = rows[0], rows[1]
    cols = [i for i, h in enumerate(head) if h in KEEP]
    with open(os.path.join(dst, name + "_ncu.csv"), "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + ["launch %d" % k for k in range(len(rows) - 2)])
        for i in cols:
            w.writerow([head[i], units[i]] + [r[i] for r in rows[2:]])
    r = rows[2]
    get = lambda k: (r[head.index(k)], units[head.index(k)]) if k in head else (None, None)
    def num(k):
        v, u = get(k)
        try:
            return float(v.replace(",", "")), u
        except Exception:
            return None, u
    def to_bytes(k):
        v, u = num(k)
        if v is None:
            return None
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    def to_us(k):
        v, u = num(k)
        if v is None:
            return None
        return v * {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1, "msecond": 1e3, "second": 1e6}.get(u, 1)
    us = to_us("gpu__time_duration.sum")
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    summary[name] = {
        "kernel": get("Kernel Name")[0], "grid": get("Grid Size")[0], "block": get("Block Size")[0], "duration_us": us,
        "dram_read_MB": None if rd is None else round(rd / 1e6, 3), "dram_write_MB": None if wr is None else round(wr / 1e6, 3),
        "dram_GBps": None if not us or rd is None else round((rd + wr) / us / 1e3, 1),
        "dram_pct_of_peak": num("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[0] or num("dram__throughput.avg.pct_of_peak_sustained_elapsed")[0],
        "sm_pct_of_peak": num("sm__throughput.avg.pct_of_peak_sustained_elapsed")[0],
        "warps_active_pct": num("sm__warps_active.avg.pct_of_peak_sustained_active")[0],
        "registers": num("launch__registers_per_thread")[0],
    }
    print(name, summary[name])
json.dump(summary, open(os.path.join(dst, "r02_kernels.json"), "w"), indent=1)
