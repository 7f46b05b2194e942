import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import oracle_binding as ob
from helpers import partition_equal
from video_segment_b200 import kernels as K
from video_segment_b200.synth import synth_clip
for (w,h,t) in [(96,72,6),(640,480,20),(1280,720,10)]:
    c = synth_clip(12, w, h, t)
    sm = np.stack([ob.preprocess(f, threads=8) for f in c])
    minr = max(1,int(np.float32(0.01)*w*np.float32(0.01)*h*20))
    ref = ob.segment_chunk_labels(sm, minr)
    print('start', (w,h,t), flush=True)
    t0=time.time()
    lab, stats = K.segment_chunk(torch.from_numpy(sm).cuda(), minr)
    torch.cuda.synchronize()
    print('done', (w,h,t), round(time.time()-t0,3), [round(x,2) for x in stats], partition_equal(ref, lab.cpu().numpy()), flush=True)
