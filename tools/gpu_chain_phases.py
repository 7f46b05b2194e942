"""Dev tool (GPU box): the bench stream (39 unique 1080p frames, ping-pong) through one handle for N chunks with the merge
debug taps on -- per-chunk merge ms, region count, and the phase split of the chunks named on the command line.
usage: python tools/gpu_chain_phases.py [n_chunks] [chunk,chunk,...]   (writes gpurun_out/chain_phases.txt)"""
import json, os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, ".."))
n_chunks = int(sys.argv[1]) if len(sys.argv) > 1 else 21
show = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1, n_chunks - 1]
os.makedirs("gpurun_out", exist_ok=True)
os.environ["VSB200_MERGE_DEBUG"] = "gpurun_out/chain_merge"
from video_segment_b200.synth import synth
from video_segment_b200.unit import DenseSegmentationUnit

w, h = 1920, 1080
frames = list(synth(3, w, h, 39))
idx = lambda k: (k % 76) if (k % 76) < 39 else 76 - (k % 76)
u = DenseSegmentationUnit()
assert u.open_streams(w, h)
rows, prev, k = [], 0.0, 0
while len(rows) < n_chunks:
    r = u.process_frame(frames[idx(k)]); k += 1
    if r:
        m = u.stats()["merge_ms"]
        rows.append({"chunk": len(rows), "merge_ms": round(m - prev, 1), "regions_first_frame": len(r[0]["region_id"])})
        prev = m
u.close()
with open("gpurun_out/chain_phases.txt", "w") as f:
    f.write("bench stream, 1080p, one handle, merge debug taps on (taps add ~5 %%)\n")
    for row in rows:
        f.write(json.dumps(row) + "\n")
    for c in show:
        p = "gpurun_out/chain_merge.chunk%d" % c
        if os.path.exists(p):
            f.write("---- chunk %d ----\n" % c)
            f.write("".join(l for l in open(p) if not l[0].isdigit()))
print(open("gpurun_out/chain_phases.txt").read()[-3000:])
