"""Dev tool (GPU box): every constrained chunk of a clip started from the ORACLE's hand-over state (import_halo) and
compared with the oracle's chunk, under merge dev-flag settings.
usage: python tools/gpu_chunk_probe.py W H T seed flags[,flags...]"""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np
import torch
import oracle_binding as ob
from helpers import overseg_iou
from video_segment_b200.synth import synth_clip
from video_segment_b200.unit import DenseSegmentationUnit

w, h, t, seed = [int(x) for x in sys.argv[1:5]]
flags = [int(x) for x in sys.argv[5].split(",")]
clip = synth_clip(seed, w, h, t)
o = ob.OracleDense(w, h, num_threads=16)
ref, handover = [], []
for f in clip:
    r = o.push(f)
    if r:
        maps, state = o.last_overlap_state()
        handover.append((len(ref) + len(r), maps, state))
    ref += r
ref += o.flush()
o.close()
for fl in flags:
    os.environ["VSB200_MERGE_FLAGS"] = str(fl)
    per_chunk, counts = [], []
    for out_so_far, maps, state in handover:
        u = DenseSegmentationUnit(want_id_maps=True)
        assert u.open_streams(w, h)
        halo = torch.from_numpy(maps).cuda()
        u.import_halo(halo[0].data_ptr(), halo[1].data_ptr(), state)
        got, k = [], out_so_far
        while not got and k < len(clip):
            got += u.process_frame(clip[k], pts=k); k += 1
        if not got:
            got += u.post_process()
        u.close()
        want = ref[out_so_far:out_so_far + len(got)]
        ious = [overseg_iou(ob.id_map_from_result(r), g["id_map"]) for g, r in zip(got, want)]
        per_chunk.append(round(min(ious), 4))
        counts.append((len(got[0]["region_id"]), len(want[0]["region_id"])))
    print(json.dumps({"flags": fl, "min_iou_per_chunk": per_chunk, "regions_first_frame_gpu_ref": counts}), flush=True)
