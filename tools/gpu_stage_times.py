"""GPU bring-up / profiling script (not a test): per-stage times of the streaming engine and the
per-bucket merge profile at the BASELINE geometries.  Usage:
    python tools/gpu_stage_times.py [engine640] [chunk1080] [engine1080]
Writes gpurun_out/stage_times.json and gpurun_out/merge_debug_*.txt."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import torch

from video_segment_b200 import kernels as K
from video_segment_b200.synth import synth
from video_segment_b200.unit import DenseSegmentationUnit

os.makedirs("gpurun_out", exist_ok=True)
what = sys.argv[1:] or ["engine640", "chunk1080", "engine1080"]
out = {}


def engine(w, h, t, seed):
    frames = list(synth(seed, w, h, min(t, 39)))
    u = DenseSegmentationUnit(device=0)
    assert u.open_streams(w, h)
    got = 0
    marks = []
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(t):
        period = 2 * (len(frames) - 1)
        i = k % period
        i = i if i < len(frames) else period - i
        r = u.process_frame(frames[i])
        if r:
            marks.append((k, len(r), time.perf_counter() - t0, dict(u.stats())))
        got += len(r)
    got += len(u.post_process())
    dt = time.perf_counter() - t0
    st = u.stats()
    u.close()
    prev = None
    per_chunk = []
    for k, n, tt, s in marks:
        if prev is not None:
            per_chunk.append({kk: round(s[kk] - prev[kk], 2) for kk in s})
        else:
            per_chunk.append({kk: round(s[kk], 2) for kk in s})
        prev = s
    return dict(frames=got, seconds=dt, fps=got / dt, stats=st, per_chunk=per_chunk)


if "engine640" in what:
    out["engine_640x480x60"] = engine(640, 480, 60, 1)
    print("engine 640x480", json.dumps(out["engine_640x480x60"]), flush=True)

if "chunk1080" in what:
    w, h, t = 1920, 1080, 20
    frames = list(synth(3, w, h, t))
    sm = torch.stack([K.preprocess(torch.from_numpy(f).cuda()) for f in frames])
    minr = int(np.float32(0.01) * w * np.float32(0.01) * h * 20)
    os.environ["VSB200_MERGE_DEBUG"] = "gpurun_out/merge_debug_1080p.txt"
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    lab, stats = K.segment_chunk(sm, minr)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    del os.environ["VSB200_MERGE_DEBUG"]
    nreg = int(torch.unique(lab).numel())
    out["chunk_1920x1080x20"] = dict(seconds=dt, edge_ms=stats[0], sort_ms=stats[1], merge_ms=stats[2], rounds=stats[3], regions=nreg)
    print("chunk 1080p", json.dumps(out["chunk_1920x1080x20"]), flush=True)
    del sm, lab
    torch.cuda.empty_cache()

if "engine1080" in what:
    out["engine_1920x1080x58"] = engine(1920, 1080, 58, 3)
    print("engine 1080p", json.dumps(out["engine_1920x1080x58"]), flush=True)

json.dump(out, open("gpurun_out/stage_times.json", "w"), indent=1)
