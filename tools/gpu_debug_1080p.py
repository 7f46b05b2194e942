"""Development probe (not a test): one flushed 1080p chunk against the oracle under the merge kernel's
development switches (VSB200_MERGE_FLAGS: 1 no hub-pair certificates, 2 fixed window target, 4 no block-0
rounds, 16 no group-parallel scans).  Prints min IoU / partition-exact frames per setting."""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE))
import numpy as np
import oracle_binding as ob
from helpers import overseg_iou, partition_equal
from video_segment_b200.synth import synth_clip
from video_segment_b200.unit import DenseSegmentationUnit

w, h, t = [int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (1920, 1080, 8))]
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 3
flags = [int(x) for x in sys.argv[5].split(",")] if len(sys.argv) > 5 else [17, 0, 16, 1]
clip = synth_clip(seed, w, h, t)
o = ob.OracleDense(w, h, num_threads=16)
ref = []
for f in clip:
    ref += o.push(f)
ref += o.flush()
ref_maps = [ob.id_map_from_result(r) for r in ref]
for fl in flags:
    os.environ["VSB200_MERGE_FLAGS"] = str(fl)
    u = DenseSegmentationUnit(want_id_maps=True)
    assert u.open_streams(w, h)
    got = []
    for f in clip:
        got += u.process_frame(f)
    got += u.post_process()
    st = u.stats()
    u.close()
    ious = [overseg_iou(a, g["id_map"]) for a, g in zip(ref_maps, got)]
    same = sum(partition_equal(a, g["id_map"]) for a, g in zip(ref_maps, got))
    print(json.dumps({"size": [w, h, t], "flags": fl, "min_iou": round(min(ious), 5), "exact_frames": same, "regions_gpu": len(got[0]["region_id"]),
                      "regions_ref": len(ref[0]["region_id"]), "merge_ms": round(st["merge_ms"], 1)}), flush=True)
