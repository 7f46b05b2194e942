"""Development sweep (not a test): builds libvsb200 variants with different merge constants on the CPU side
(`python tools/gpu_merge_sweep_constants.py build`), then times three 1080p chunks per variant on the GPU
(`python tools/gpu_merge_sweep_constants.py run`) and prints merge ms per chunk plus a checksum of the region counts."""
import json, os, subprocess, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
VARIANTS = {                       # -D overrides of the constants at the top of csrc/merge.cu
    "base": [],
    "grp1024": ["-DVSB_GROUP_SCAN_MIN=1024"],
    "grp32": ["-DVSB_GROUP_SCAN_MIN=32"],
    "split8k": ["-DVSB_RESIDUAL_SPLIT=8192"],
    "win20": ["-DVSB_WINDOW_TARGET=(1ull<<20)"],
}
OUT = os.path.join(ROOT, "tests", "build", "sweep_libs")


def build():
    from video_segment_b200 import build as B
    os.makedirs(OUT, exist_ok=True)
    objs = [os.path.join(B.HERE, "build", s.replace(".cu", ".o")) for s in B.SOURCES if s != "merge.cu"]
    procs = []
    for name, defs in VARIANTS.items():
        obj = os.path.join(OUT, f"merge_{name}.o")
        procs.append((name, obj, subprocess.Popen([B._nvcc()] + B.NVCC_FLAGS + defs + ["-c", os.path.join(B.CSRC, "merge.cu"), "-o", obj])))
    for name, obj, p in procs:
        assert p.wait() == 0, name
        subprocess.check_call([B._nvcc(), "-shared", "-o", os.path.join(OUT, f"libvsb200_{name}.so")] + objs + [obj, "-gencode", "arch=compute_100a,code=sm_100a"])
    print("built", list(VARIANTS))


def one():
    import numpy as np
    import torch
    from video_segment_b200.synth import synth
    from video_segment_b200.unit import DenseSegmentationUnit
    w, h = 1920, 1080
    cache = "/tmp/vsb_sweep_frames.npy"
    if os.path.exists(cache):
        frames = list(np.load(cache))
    else:
        frames = list(synth(3, w, h, 39))
        np.save(cache, np.stack(frames))
    dev = [torch.from_numpy(f).cuda() for f in frames]
    u = DenseSegmentationUnit(device=0)
    assert u.open_streams(w, h)
    merges, regions, prev = [], [], 0.0
    for k in range(20 + 19 * 3):
        i = k % 76
        i = i if i < 39 else 76 - i
        r = u.process_device_frame(dev[i].data_ptr(), w * 3)
        if r:
            m = u.stats()["merge_ms"]
            merges.append(round(m - prev, 1)); prev = m
            regions.append(int(sum(len(x["region_id"]) for x in r)))
    u.close()
    print(json.dumps({"variant": os.environ.get("VSB_VARIANT"), "merge_ms": merges, "regions": regions}), flush=True)


if __name__ == "__main__":
    if sys.argv[1:] == ["build"]:
        build()
    elif sys.argv[1:] == ["one"]:
        one()
    else:
        for name in (sys.argv[1:] or list(VARIANTS)):
            e = dict(os.environ); e["VSB_VARIANT"] = name; e["VSB200_LIB"] = os.path.join(OUT, f"libvsb200_{name}.so")
            subprocess.run([sys.executable, os.path.abspath(__file__), "one"], env=e, check=False, timeout=120)
