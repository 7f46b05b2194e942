"""First GPU bring-up script (not a test): parity spot checks + kernel timings, verbose."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import oracle_binding as ob
from helpers import overseg_iou, partition_equal
from video_segment_b200 import kernels as K
from video_segment_b200.synth import synth_clip

def ev_time(fn, n=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(True), torch.cuda.Event(True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n

out = {}
print(torch.cuda.get_device_name(0))
clip = synth_clip(2, 1920, 1080, 3)
d = [torch.from_numpy(f).cuda() for f in clip]
t = ev_time(lambda: K.preprocess(d[0]))
out['preprocess_ms_1080p'] = t; print('preprocess ms', t)
sm = [K.preprocess(x) for x in d]
W, H = 1920, 1080; N = W*H
sp = torch.empty((H, W, 4), dtype=torch.float32, device='cuda'); tp = torch.empty((H, W, 9), dtype=torch.float32, device='cuda')
t = ev_time(lambda: K.edge_build(sm[1], sm[0], None, False, sp, tp), n=20)
Es = (W-1)*H + W*(H-1) + 2*(W-1)*(H-1); Et = (3*W-2)*(3*H-2)
alg = 24*N + 4*(Es+Et)
out['edge_ms_1080p'] = t; out['edge_GBps_alg'] = alg/t/1e6; print('edge ms', t, 'alg GB/s', alg/t/1e6)
ref = ob.preprocess(clip[0], threads=8)
print('preprocess max err 1080p', float((sm[0].cpu().numpy()-ref).__abs__().max()), 'neq frac', float((sm[0].cpu().numpy()!=ref).mean()))
# merge bring-up at growing sizes
for (w,h,tt) in [(64,48,4),(160,120,6),(320,240,8),(640,480,20)]:
    c = synth_clip(3, w, h, tt)
    smn = np.stack([ob.preprocess(f, threads=8) for f in c])
    minr = int(np.float32(0.01)*w*np.float32(0.01)*h*20)
    t0 = time.time(); refl = ob.segment_chunk_labels(smn, minr); t_ref = time.time()-t0
    torch.cuda.synchronize(); t0 = time.time()
    lab, stats = K.segment_chunk(torch.from_numpy(smn).cuda(), minr)
    torch.cuda.synchronize(); t_gpu = time.time()-t0
    lab = lab.cpu().numpy()
    iou = min(overseg_iou(refl[k], lab[k]) for k in range(tt))
    print((w,h,tt), 'oracle s', round(t_ref,2), 'gpu s', round(t_gpu,3), 'stats', [round(x,2) for x in stats], 'iou', iou, 'equal', partition_equal(refl, lab), 'regions', len(np.unique(refl)), len(np.unique(lab)))
    out[f'merge_{w}x{h}x{tt}'] = dict(oracle_s=t_ref, gpu_s=t_gpu, stats=stats, iou=iou)
os.makedirs('gpurun_out', exist_ok=True)
json.dump(out, open('gpurun_out/first_run.json','w'), indent=1)
