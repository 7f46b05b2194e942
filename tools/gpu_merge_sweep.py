"""Dev tool (GPU box): sweep of the merge window parameters (VSB200_WINDOW_TARGET / _RESIDUAL_SPLIT / _SEGMENT_MIN) on a
multi-chunk 1080p clip; prints merge ms per chunk and a digest of the id maps (all settings must give the same digest).
usage: python tools/gpu_merge_sweep.py T 'wt,rs,sm;wt,rs,sm;...'"""
import hashlib, json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "tests")); sys.path.insert(0, os.path.join(HERE, ".."))
import numpy as np
from video_segment_b200.synth import synth_clip
from video_segment_b200.unit import DenseSegmentationUnit

t = int(sys.argv[1])
clip = synth_clip(2, 1920, 1080, t)
for cfg in sys.argv[2].split(";"):
    wt, rs, sm = [int(x) for x in cfg.split(",")]
    os.environ["VSB200_WINDOW_TARGET"] = str(wt); os.environ["VSB200_RESIDUAL_SPLIT"] = str(rs); os.environ["VSB200_SEGMENT_MIN"] = str(sm)
    u = DenseSegmentationUnit(want_id_maps=True)
    assert u.open_streams(1920, 1080)
    per_chunk, prev, h = [], 0.0, hashlib.sha256()
    def take(res):
        global prev
        if res:
            m = u.stats()["merge_ms"]; per_chunk.append(round(m - prev, 1)); prev = m
        for r in res:
            h.update(r["id_map"].tobytes())
    for f in clip:
        take(u.process_frame(f))
    take(u.post_process())
    u.close()
    print(json.dumps({"window_target": wt, "residual_split": rs, "segment_min": sm, "merge_ms_per_chunk": per_chunk,
                      "sum": round(sum(per_chunk), 1), "digest": h.hexdigest()[:12]}), flush=True)
