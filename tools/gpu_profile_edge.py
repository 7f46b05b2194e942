"""Small driver for ncu: a few steady-state edge-build launches at 1080p (plus preprocess)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from video_segment_b200 import kernels as K
from video_segment_b200.synth import synth_clip
W, H = 1920, 1080
clip = synth_clip(2, W, H, 2)
d = [torch.from_numpy(f).cuda() for f in clip]
sm = [K.preprocess(x) for x in d]
sp = torch.empty((H, W, 4), dtype=torch.float32, device="cuda")
tp = torch.empty((H, W, 9), dtype=torch.float32, device="cuda")
for _ in range(6):
    K.edge_build(sm[1], sm[0], None, False, sp, tp)
torch.cuda.synchronize()
print("ok")
