"""Times the fused Lab + per-region histogram pass (csrc/region_hist.cu) at 1080p with CUDA events:
algorithmic bytes = 3 B (BGR) + 4 B (region id) per pixel."""
import ctypes as C, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import torch
from video_segment_b200 import kernels as K
from video_segment_b200._lib import lib
from video_segment_b200.synth import synth_clip

W, H, NR = 1920, 1080, 2000
clip = synth_clip(3, W, H, 4)
rng = np.random.default_rng(1)
ids = [np.kron(rng.integers(0, NR, size=(H // 40, W // 40)).astype(np.int32), np.ones((40, 40), np.int32)) for _ in range(4)]
d_bgr = [torch.from_numpy(f).cuda() for f in clip]
d_ids = [torch.from_numpy(m).cuda() for m in ids]
sb = lib().vsb200_region_hist_scratch_bytes(NR, 10, 20)
scratch = torch.empty(sb, dtype=torch.uint8, device="cuda")
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
lib().vsb200_region_hist_reset(C.c_void_p(scratch.data_ptr()), NR, 10, 20, st)
def add(i):
    lib().vsb200_region_hist_add(C.c_void_p(d_bgr[i % 4].data_ptr()), W * 3, C.c_void_p(d_ids[i % 4].data_ptr()), W, H, NR, 10, 20,
                                 C.c_void_p(scratch.data_ptr()), st)
for i in range(4):
    add(i)
torch.cuda.synchronize()
n = 40
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(n):
    add(i)
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / n
print(json.dumps({"kernel": "region_hist_kernel", "workload": "1920x1080, 2000 regions of 40x40 blocks, 10x20x20 bins", "us_per_frame": round(us, 1),
                  "algorithmic_GBps": round(W * H * 7 / us / 1e3, 1), "scratch_MB": round(sb / 1e6, 1)}))
