"""Development probe (not a test): one TMA edge-build launch at a small size, compared with the staged-load kernel."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from video_segment_b200 import kernels as K
h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 48, int(sys.argv[2]) if len(sys.argv) > 2 else 64
rng = np.random.default_rng(0)
a = torch.from_numpy(rng.random((h, w, 3), dtype=np.float32)).cuda()
b = torch.from_numpy(rng.random((h, w, 3), dtype=np.float32)).cuda()
sp, tp = K.edge_build(a, b)
torch.cuda.synchronize()
print("tma launch ok", float(sp.sum()), float(tp.sum()))
