"""Development probe: region sizes reported in the hierarchy (device union-find records) must equal the voxel
counts of the id maps (self-consistency, no oracle).  usage: gpu_size_check.py [real|WxHxT] [repeats]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests")); sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from video_segment_b200.synth import synth_clip
from video_segment_b200.unit import DenseSegmentationUnit
which = sys.argv[1] if len(sys.argv) > 1 else "real"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which == "real":
    clip = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "real_clip_136x240x24.npz"))["frames"][:12]
else:
    w, h, t = map(int, which.split("x"))
    clip = synth_clip(3, w, h, t)
h, w = clip[0].shape[:2]
bad_total = 0
for rep in range(reps):
    u = DenseSegmentationUnit(want_id_maps=True)
    assert u.open_streams(w, h)
    out = []
    for f in clip:
        out += u.process_frame(f)
    out += u.post_process()
    u.close()
    ids = np.stack([o["id_map"] for o in out])
    uid, cnt = np.unique(ids, return_counts=True)
    true = dict(zip(uid.tolist(), cnt.tolist()))
    comp = out[0]["compound"]
    bad = [(int(r[0]), int(r[1]), true.get(int(r[0]), 0)) for r in comp if true.get(int(r[0]), 0) != int(r[1])]
    bad_total += len(bad)
    print(f"rep {rep}: regions {len(comp)} size mismatches {len(bad)} {bad[:5]}", flush=True)
print("SIZE_CHECK", "OK" if bad_total == 0 else "FAIL", os.environ.get("VSB200_MERGE_FLAGS", "0"))
