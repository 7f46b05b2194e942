import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle_binding as ob
from helpers import partition_equal
from video_segment_b200 import kernels as K
from video_segment_b200.synth import synth_clip
# dirty the allocator like the preceding tests do
junk = [torch.full((1 << 22,), 0x5A5A5A5A, dtype=torch.int32, device='cuda') for _ in range(64)]
del junk
torch.cuda.empty_cache()
clip = synth_clip(11, 160, 120, 4)
sm = [ob.preprocess(f) for f in clip]
for (w,h,t) in [(96,72,6)]:
    c = synth_clip(12, w, h, t)
    smn = np.stack([ob.preprocess(f, threads=8) for f in c])
    minr = max(1,int(np.float32(0.01)*w*np.float32(0.01)*h*20))
    ref = ob.segment_chunk_labels(smn, minr)
    print('start', (w,h,t), flush=True)
    lab, stats = K.segment_chunk(torch.from_numpy(smn).cuda(), minr)
    torch.cuda.synchronize()
    print('done', (w,h,t), [round(x,2) for x in stats], partition_equal(ref, lab.cpu().numpy()), flush=True)
