"""Dev tool (GPU box): multi-chunk clip through the engine under merge dev-flag settings, every frame compared
with the oracle; per-chunk merge ms and the per-bucket / per-phase profile of every chunk.
usage: python tools/gpu_merge_probe.py W H T seed flags[,flags..] [tag]"""
import json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np
import oracle_binding as ob
from helpers import overseg_iou, partition_equal
from video_segment_b200.synth import synth_clip
from video_segment_b200.unit import DenseSegmentationUnit

w, h, t, seed = [int(x) for x in sys.argv[1:5]]
flags = [int(x) for x in sys.argv[5].split(",")]
tag = sys.argv[6] if len(sys.argv) > 6 else "probe"
os.makedirs("gpurun_out", exist_ok=True)
clip = synth_clip(seed, w, h, t)
ref_maps = None
if os.environ.get("PROBE_NO_ORACLE") is None:
    t0 = time.time()
    o = ob.OracleDense(w, h, num_threads=16)
    ref = []
    for f in clip:
        ref += o.push(f)
    ref += o.flush()
    ref_maps = [ob.id_map_from_result(r) for r in ref]
    print("oracle %.1fs" % (time.time() - t0), flush=True)
for fl in flags:
    os.environ["VSB200_MERGE_FLAGS"] = str(fl)
    if os.environ.get("PROBE_NO_DEBUG") is None:
        os.environ["VSB200_MERGE_DEBUG"] = "gpurun_out/%s_f%d_merge" % (tag, fl)
    u = DenseSegmentationUnit(want_id_maps=True)
    assert u.open_streams(w, h)
    got, per_chunk, prev = [], [], 0.0
    for f in clip:
        r = u.process_frame(f)
        if r:
            m = u.stats()["merge_ms"]; per_chunk.append(round(m - prev, 1)); prev = m
        got += r
    got += u.post_process()
    m = u.stats()["merge_ms"]; per_chunk.append(round(m - prev, 1))
    st = u.stats()
    u.close()
    rec = {"size": [w, h, t], "flags": fl, "merge_ms_per_chunk": per_chunk, "stats": {k: round(v, 1) for k, v in st.items()}}
    if ref_maps is not None:
        ious = [overseg_iou(a, g["id_map"]) for a, g in zip(ref_maps, got)]
        same = [bool(partition_equal(a, g["id_map"])) for a, g in zip(ref_maps, got)]
        chunk_ids = [g["chunk_id"] for g in got]
        per_chunk_iou = {}
        for c, v in zip(chunk_ids, ious):
            per_chunk_iou[c] = min(per_chunk_iou.get(c, 1.0), v)
        rec.update(min_iou_per_chunk=[round(per_chunk_iou[c], 4) for c in sorted(per_chunk_iou)])
        rec.update(min_iou=round(min(ious), 6), exact_frames=int(sum(same)), frames=len(got),
                   first_inexact=(same.index(False) if False in same else -1),
                   regions_gpu=[len(g["region_id"]) for g in got][::6], regions_ref=[len(r["region_id"]) for r in ref][::6])
    print(json.dumps(rec), flush=True)
    with open("gpurun_out/%s.jsonl" % tag, "a") as fo:
        fo.write(json.dumps(rec) + "\n")
