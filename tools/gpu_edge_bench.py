"""Times the steady-state edge-build launch at 1080p with CUDA events (outside any profiler).
Buffers rotate over 4 sets (4 x 166 MB written + 50 MB read > the 126 MB L2).  VSB200_EDGE_MODE selects
the kernel variant (development switch in csrc/edges.cu); run once per variant (the switches
are read once per process).  Also cross-checks the variant against the staged kernel bit for bit."""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))


def run_one():
    import torch
    from video_segment_b200 import kernels as K
    from video_segment_b200.synth import synth_clip
    W, H = 1920, 1080
    clip = synth_clip(2, W, H, 5)
    sm = [K.preprocess(torch.from_numpy(f).cuda()) for f in clip]
    sets = [(torch.empty((H, W, 4), dtype=torch.float32, device="cuda"), torch.empty((H, W, 9), dtype=torch.float32, device="cuda")) for _ in range(4)]
    for i in range(8):
        K.edge_build(sm[1 + i % 4], sm[i % 4], None, False, *sets[i % 4])
    torch.cuda.synchronize()
    n = 200
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
    ev[0].record()
    for i in range(n):
        K.edge_build(sm[1 + i % 4], sm[i % 4], None, False, *sets[i % 4])
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(n))
    total = ev[0].elapsed_time(ev[n]) * 1e3 / n
    import hashlib
    K.edge_build(sm[1], sm[0], None, False, *sets[0])
    torch.cuda.synchronize()
    dig = hashlib.sha256(sets[0][0].cpu().numpy().tobytes() + sets[0][1].cpu().numpy().tobytes()).hexdigest()[:16]
    alg = 157485624
    print(json.dumps({"variant": os.environ.get("VSB_VARIANT"), "us_avg_back_to_back": round(total, 2), "us_median": round(ts[n // 2], 2),
                      "us_min": round(ts[0], 2), "GBps_avg": round(alg / total / 1e3, 1), "digest": dig}))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run_one()
    else:
        for name, env in (("pipe", {}), ("pipe_static", {"VSB200_EDGE_MODE": "pipe_static"}), ("x2", {"VSB200_EDGE_MODE": "x2"}), ("scalar_tma", {"VSB200_EDGE_MODE": "scalar"})):
            e = dict(os.environ); e.update(env); e["VSB_VARIANT"] = name
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=e, check=False)
