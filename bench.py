#!/usr/bin/env python
"""bench.py -- frames/s of dense video over-segmentation at 1920x1080 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W              # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...     # the reference's own CPU path (oracle/_ref)

A "step" is one chunk of the streaming hot path: 19 new 1080p frames pushed through the
DenseSegmentationUnit mirror (preprocess -> edge build -> bucket sort -> merge -> labels / N4 /
RLE -> region bookkeeping), i.e. exactly what the reference outputs per chunk
(dense_segmentation.cpp:281-432).  Rank 0 prints ONE JSON line (see the task contract):
  value   : frames/s with the u8 frames already resident in HBM (vsb200_dense_push_device)
  e2e     : frames/s through the public host API (host frames in, host SegmentationDesc arrays
            out): every step copies its 19 frames H2D from pinned staging and reads the
            rasterisation runs / neighbour pairs back
  roofline: edge-build kernel, algorithmic bytes / CUDA-event time of its launches inside the
            timed region, against MEASURED_PEAKS.json's HBM copy bandwidth
  cpu_baseline : the reference's own DenseSegmentation pipeline compiled unmodified into
            oracle/_ref/libref_results.so (kind "reference"; the oracle port if that prebuilt
            file is missing, kind "port") timed on the same box's host cores on a bounded sample
            (rank 0, N = 1 only)
Multi GPU (torchrun): frame-chunk groups of ONE synthetic video are sharded over the ranks
(weak scaling, fixed frames per GPU); the only data-path exchange is the seam hand-over of the
two overlap frames' region-id maps (NCCL send/recv over NVLink, C1) and the all-gather of the
groups' region-id counts (C2); both run inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "frames/sec 1080p dense over-seg"
FRAMES_PER_STEP = 19          # new frames per chunk (chunk_size 20, 1 virtual + 1 constrained overlap)
UNIQUE_FRAMES = 39            # generated frames per rank; longer runs ping-pong over them
WORKLOAD = "BASELINE config 3 (dense half): {w}x{h} synthetic stream, dense over-segmentation, 20-slot constrained chunks"


def edge_build_bytes(w, h):
    n = w * h
    es = (w - 1) * h + w * (h - 1) + 2 * (w - 1) * (h - 1)
    et = (3 * w - 2) * (3 * h - 2)
    return 24 * n + 4 * (es + et)          # 12N + 12N read, 4 (Es + Et) written


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, indices):
        super().__init__(daemon=True)
        self.indices, self.rows, self.stop_flag = set(indices), [], False

    def run(self):
        # ONE sampler per job (rank 0), one query for all GPUs of the job: eight ranks polling nvidia-smi next to each
        # other serialise on the driver and show up in everybody's step time
        q = ("index,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                for line in out.splitlines():
                    f = [x.strip() for x in line.split(",")]
                    if f and f[0].isdigit() and int(f[0]) in self.indices:
                        self.rows.append(f[1:])
            except Exception:
                pass
            time.sleep(0.5)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.rows)}


def frame_index(k):
    """Ping-pong over the generated frames keeps the video temporally continuous."""
    period = 2 * (UNIQUE_FRAMES - 1)
    k %= period
    return k if k < UNIQUE_FRAMES else period - k


def cpu_engine():
    """The CPU implementation of the path the CPU legs time: the reference's own DenseSegmentation pipeline
    (oracle/_ref/libref_results.so, built from /root/reference where that is mounted and shipped as a prebuilt
    file) or, without it, the oracle port.  Returns (kind, factory(w, h), threads)."""
    import oracle_binding as ob
    import reference_binding as rb
    cores = os.cpu_count() or 1
    ok = rb.available(build=False)
    if ok:
        try:
            rb.lib()
        except OSError as e:      # a prebuilt file this box cannot load: say so and time the port instead
            print(f"bench: {rb.LIB_PATH} not loadable ({e}); timing the oracle port", file=sys.stderr)
            ok = False
    if ok:
        # threads: the reference parallelises graph construction with one std::thread per frame
        # (FLAGS_parallel_graph_construction) and the bilateral filter with OpenMP over 8 row blocks
        return "reference", (lambda w, h: rb.ReferenceDense(w, h)), cores
    return "port", (lambda w, h: ob.OracleDense(w, h, num_threads=cores)), cores


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (its DenseSegmentation::ProcessFrame
    stream compiled unmodified, oracle/Makefile; seg_tree_sample's decode / hierarchy / writer stages are outside the
    path) on the box's host cores, on the SAME workload and step as the GPU arm: one continuous stream of the same
    synthetic clip, a step = one constrained 20-slot chunk = 19 new frames.  The first (unconstrained, 20-frame) chunk
    and min(W, 1) constrained chunks are warm-up; K constrained chunks are timed."""
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)          # torchrun pins it to 1; the reference's OpenMP loops want the cores
    from video_segment_b200.synth import synth
    w, h = args.width, args.height
    kind, make, cores = cpu_engine()
    frames = list(synth(3, w, h, UNIQUE_FRAMES))
    o = make(w, h)
    k = 0
    def step():
        nonlocal k
        got = 0
        while got < FRAMES_PER_STEP:
            got += len(o.push(frames[frame_index(k)]))
            k += 1
        return got
    for _ in range(1 + min(args.warmup, 1)):
        step()
    series = []
    t0 = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        t1 = time.perf_counter()
        total += step()
        series.append(round(1000.0 * (time.perf_counter() - t1), 1))
    dt = time.perf_counter() - t0
    o.flush()           # untimed: the reference joins its graph-construction threads on flush, not in its destructor
    o.close()
    fps = total / dt
    sample = (f"one stream of the {w}x{h} synthetic clip (seed 3, ping-pong over {UNIQUE_FRAMES} frames): {args.steps} constrained chunks of "
              f"{FRAMES_PER_STEP} new frames timed after the free first chunk")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD.format(w=w, h=h), "step": f"one chunk = {FRAMES_PER_STEP} new frames", "threads": cores},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "ms_per_step_series": series,
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--hierarchy", action="store_true",
                    help="N = 1 only: add a leg that runs the whole of config 3 -- dense over-segmentation feeding the hierarchical "
                         "region stage (RegionSegmentationUnit, defaults) -- and report it under \"hierarchy\"")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from video_segment_b200._lib import lib
    from video_segment_b200.synth import synth
    from video_segment_b200.unit import DenseSegmentationUnit

    if not torch.cuda.is_available() or lib().vsb200_device_count() < 1:
        raise SystemExit("bench.py: no B200 / CUDA library -- this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # O(#scan intervals) host shaping threads of the engine: the box's cores are shared by the ranks
    os.environ.setdefault("VSB200_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // max(world, 1)))))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    w, h = args.width, args.height
    n = w * h
    K, W = args.steps, args.warmup
    if W < 3:
        print("bench.py: warning: fewer than 3 warm-up steps", file=sys.stderr)
    from video_segment_b200.shard import group_range
    frames_per_rank = 1 + FRAMES_PER_STEP * (W + K)
    # One video, sharded: rank g owns frames [g * L, (g + 1) * L] (one read-overlap frame).
    start, _ = group_range(rank, world, frames_per_rank)
    host_frames = list(synth(3, w, h, min(UNIQUE_FRAMES, frames_per_rank), start=start))
    dev_frames = [torch.from_numpy(f).cuda() for f in host_frames]
    pinned = [torch.from_numpy(f).pin_memory() for f in host_frames]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # group seams: NCCL in the C++ host layer (csrc/shard.cu); the id travels over torch.distributed's store
    from video_segment_b200.shard import SeamLink, nccl_unique_id
    link = None
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        link = SeamLink(uid[0], rank, world, local_rank, w, h)
    first_map_dev = torch.zeros((h, w), dtype=torch.int32, device="cuda")

    EXCHANGE_EVERY = 5      # chunks per frame group: the seam hand-over runs at every group boundary inside the timed region

    def seam_exchange(unit, have_first_map):
        """Group boundary: C1 (overlap id maps to the successor, ncclSend / ncclRecv), C2 (all-gather of the region-id
        counts), then the seam vote on the device -> relabel table of this group (what a writer applies to its ids)."""
        if link is None:
            return None
        offsets = link.exchange(unit)
        n_ids = max(1, offsets[rank + 1] - offsets[rank])
        return link.relabel_table(first_map_dev.data_ptr() if have_first_map else 0, n_ids, offsets[rank])

    def run_leg(device_resident):
        unit = DenseSegmentationUnit(device=local_rank)
        if not unit.open_streams(w, h):
            raise SystemExit("bench.py: " + lib().vsb200_last_error().decode())
        unit.set_profiling(True)
        k = 0
        def push_next():
            nonlocal k
            i = frame_index(k)
            k += 1
            if device_resident:
                return unit.process_device_frame(dev_frames[i].data_ptr(), w * 3)
            return unit.process_frame(pinned[i].numpy())
        out_frames = 0
        have_first = False           # this group's own segmentation of its first frame (shared with the predecessor) is on the device
        # warm-up: W chunks (the first one takes 20 frames)
        while out_frames < FRAMES_PER_STEP * W:
            res = push_next()
            if res and not have_first and world > 1:
                from video_segment_b200.unit import id_map_from_result
                first_map_dev.copy_(torch.from_numpy(id_map_from_result(res[0])))
                have_first = True
            out_frames += len(res)
        seam_exchange(unit, have_first)  # warm-up of the exchange too (NCCL opens its peer channels on first use)
        st0, io0 = unit.stats(), unit.io_stats()
        link0 = link.stats() if link else None
        visible = [int(x) for x in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if x.strip().isdigit()]
        gpu_ids = [visible[i] if i < len(visible) else i for i in range(world)]      # physical indices of local ranks 0 .. world-1
        sampler = ClockSampler(gpu_ids) if rank == 0 else None
        barrier()
        if sampler:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        timed, chunks, series, regions, t_prev = 0, 0, [], [], t0
        while timed < FRAMES_PER_STEP * K:
            res = push_next()
            if res:
                timed += len(res)
                chunks += 1
                now = time.perf_counter()
                series.append(round(1000.0 * (now - t_prev), 1))
                regions.append(int(len(res[0]["region_id"])))
                t_prev = now
                if chunks % EXCHANGE_EVERY == 0 or timed >= FRAMES_PER_STEP * K:
                    seam_exchange(unit, have_first)
                    t_prev = time.perf_counter()
        torch.cuda.synchronize()
        e1.record()
        barrier()
        wall = time.perf_counter() - t0
        if sampler:
            sampler.stop_flag = True
        ms = e0.elapsed_time(e1)
        st1, io1 = unit.stats(), unit.io_stats()
        link1 = link.stats() if link else None
        unit.close()
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        d = {k2: st1[k2] - st0[k2] for k2 in st1}
        dio = {k2: io1[k2] - io0[k2] for k2 in io1}
        mine = {"rank": rank, "ms": round(ms, 1), "merge_ms": round(d["merge_ms"], 1), "host_shape_ms": round(d["host_shape_ms"], 1),
                "exchange_ms": round(link1["exchange_ms"] - link0["exchange_ms"], 2) if link else 0.0,
                "exchanges": int(link1["exchanges"] - link0["exchanges"]) if link else 0}
        per_rank = [mine]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, mine)
            per_rank = gathered
        return dict(ms=float(t.item()), wall=wall, frames=timed, stats=d, io=dio, clocks=sampler.summary() if sampler else None, series=series,
                    regions=regions, per_rank=per_rank)

    leg_dev = run_leg(True)
    leg_e2e = run_leg(False)

    total_frames = FRAMES_PER_STEP * K * world
    value = total_frames / (leg_dev["ms"] / 1000.0)
    e2e = total_frames / (leg_e2e["ms"] / 1000.0)
    # roofline of the edge-build kernel (dominant HBM kernel named by BASELINE.json)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6.65 TB/s (B200_PROFILING.md)"
    io = leg_dev["io"]
    edge_ms = io["edge_ms"] / max(io["edge_launches"], 1)
    achieved = edge_build_bytes(w, h) / (edge_ms * 1e-3) / 1e9 if edge_ms > 0 else 0.0
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "edge_build_traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": leg_dev["ms"] / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (video_segment_b200.synth seed 3, 39 unique frames per GPU, ping-pong)",
        "config": {"workload": WORKLOAD.format(w=w, h=h),
                   "step": f"one chunk = {FRAMES_PER_STEP} new frames per GPU", "frames_per_gpu": FRAMES_PER_STEP * K,
                   "seam_exchange": f"every {EXCHANGE_EVERY} chunks (csrc/shard.cu: ncclSend/ncclRecv + ncclAllGather, vote + relabel on the device)" if world > 1 else "none (1 GPU)",
                   "parallelism": f"frame-chunk groups x{world}", "l2": "inputs larger than L2: 21 slots x 24.9 MB frames + 2.3 GB edge weights per chunk"},
        "e2e": {"value": e2e, "unit": "frames/s",
                "h2d_bytes_per_step": leg_e2e["io"]["h2d_bytes"] / K, "d2h_bytes_per_step": leg_e2e["io"]["d2h_bytes"] / K},
        "gpu_launches": int(leg_dev["stats"]["kernel_launches"]),
        "clocks": leg_dev["clocks"],
        "roofline": {"bound": "hbm", "kernel": "edge_build_tma_pipe_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "bytes_per_launch": edge_build_bytes(w, h), "ms_per_launch": edge_ms, "launches_timed": int(io["edge_launches"])},
        "stage_ms_per_step": {k2: v / K for k2, v in leg_dev["stats"].items() if k2.endswith("_ms")},
        "merge_rounds_per_step": leg_dev["stats"]["merge_rounds"] / K,
        "ms_per_step_series": leg_dev["series"], "regions_per_step": leg_dev["regions"],
        "per_rank": leg_dev["per_rank"],
    }
    if rank == 0 and world == 1 and args.hierarchy:
        # config 3 end to end: host frames -> DenseSegmentationUnit -> RegionSegmentationUnit (segmentation tree), flushed
        from video_segment_b200.unit import RegionSegmentationUnit
        dense = DenseSegmentationUnit(device=local_rank)
        region = RegionSegmentationUnit(raw_records=True)
        assert dense.open_streams(w, h) and region.open_streams(w, h)
        n_in = 1 + FRAMES_PER_STEP * (W + K)
        fed, out, t_region, levels = 0, 0, 0.0, 0
        def feed(results):
            nonlocal fed, out, t_region, levels
            for r in results:
                t1 = time.perf_counter()
                recs = region.process_frame(r, pinned[frame_index(fed)].numpy())
                t_region += time.perf_counter() - t1
                fed += 1
                out += len(recs)
                for rec in recs:
                    levels = max(levels, int(rec[7]))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for kf in range(n_in):
            feed(dense.process_frame(pinned[frame_index(kf)].numpy()))
        feed(dense.post_process())
        t1 = time.perf_counter()
        out += len(region.post_process())
        t_region += time.perf_counter() - t1
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        line["hierarchy"] = {"value": out / dt, "unit": "frames/s", "frames": out, "seconds": round(dt, 2),
                             "region_stage_ms_per_frame": round(1000.0 * t_region / max(out, 1), 2),
                             "region_stage_share": round(t_region / dt, 3), "hierarchy_levels": levels,
                             "what": "config 3 whole: dense over-segmentation + hierarchical region stage (RegionSegmentationOptions "
                                     "defaults: chunk sets of 6 chunks, overlap 2, appearance descriptor), host frames in, records out, flushed"}
        dense.close(); region.close()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        kind, make, cores = cpu_engine()
        nfr = 1 + 2 * FRAMES_PER_STEP      # the free first chunk and one constrained chunk: ~20-30 s of CPU work at 1080p
        o = make(w, h)
        t0 = time.perf_counter()
        got = 0
        for kf in range(nfr):
            got += len(o.push(host_frames[frame_index(kf)]))
        got += len(o.flush())
        dt = time.perf_counter() - t0
        o.close()
        line["cpu_baseline"] = {"value": got / dt, "unit": "frames/s", "cores": cores, "kind": kind,
                                "sample": f"first {nfr} frames of the same stream: the free 20-frame chunk and one constrained chunk, flushed ({dt:.1f} s)"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
