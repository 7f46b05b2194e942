"""Dev tool (GPU box): config 3 whole (dense over-segmentation feeding the hierarchical region stage) at 1080p under cProfile --
where the host time of the region half goes.  usage: python tools/profile_hierarchy.py [frames]"""
import cProfile, os, pstats, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
from video_segment_b200.synth import synth
from video_segment_b200.unit import DenseSegmentationUnit, RegionSegmentationUnit

n = int(sys.argv[1]) if len(sys.argv) > 1 else 134
w, h = 1920, 1080
frames = list(synth(3, w, h, 39))
idx = lambda k: (k % 76) if (k % 76) < 39 else 76 - (k % 76)
dense, region = DenseSegmentationUnit(), RegionSegmentationUnit(raw_records=True)
assert dense.open_streams(w, h) and region.open_streams(w, h)
over = []
for k in range(n):
    over += dense.process_frame(frames[idx(k)])
over += dense.post_process()
dense.close()
def run():
    out = 0
    for k, r in enumerate(over):
        out += len(region.process_frame(r, frames[idx(k)]))
    out += len(region.post_process())
    return out
t0 = time.perf_counter()
pr = cProfile.Profile(); pr.enable(); out = run(); pr.disable()
print("region stage: %d frames in %.2f s" % (out, time.perf_counter() - t0))
pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
