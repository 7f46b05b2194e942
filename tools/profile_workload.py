"""Dev tool (GPU box): a short pass of the whole path for the profiler -- one free 1080p chunk plus the start of a
constrained one through DenseSegmentationUnit, then the region stage on a small clip.
usage: python tools/profile_workload.py [W H T]"""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..")); sys.path.insert(0, os.path.join(HERE, "..", "tests"))
import numpy as np
from video_segment_b200.synth import synth, synth_flow
from video_segment_b200.unit import (DenseSegmentationOptions, DenseSegmentationUnit, RegionSegmentationOptions,
                                     RegionSegmentationUnit)

w, h, t = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (1920, 1080, 22)
u = DenseSegmentationUnit()
assert u.open_streams(w, h)
n = 0
for f in synth(3, w, h, t):
    n += len(u.process_frame(f))
n += len(u.post_process())
u.close()
print("dense frames", n)
if os.environ.get("PROFILE_REGION", "1") == "1":
    pairs = list(synth_flow(5, 640, 360, 30))
    d = DenseSegmentationUnit(dense_seg_options=DenseSegmentationOptions(chunk_size=8))
    assert d.open_streams(640, 360, flow_stream_present=True)
    r = RegionSegmentationUnit(RegionSegmentationOptions(chunk_set_size=3, chunk_set_overlap=1), raw_records=True)
    assert r.open_streams(640, 360, flow_stream_present=True)
    k, m = 0, 0
    def feed(res):
        global k, m
        for x in res:
            m += len(r.process_frame(x, pairs[k][0], None if k == 0 else pairs[k][1])); k += 1
    for i, (f, fl) in enumerate(pairs):
        feed(d.process_frame(f, fl))
    feed(d.post_process())
    m += len(r.post_process())
    print("region frames", m)
