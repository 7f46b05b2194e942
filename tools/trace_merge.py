"""Dev tool (CPU): per-bucket statistics of the reference's sequential merge (oracle, VSO_TRACE_MERGE).
usage: python tools/trace_merge.py W H T [seed]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
os.environ["VSO_TRACE_MERGE"] = "1"
import oracle_binding as ob
from video_segment_b200.synth import synth
W, H, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 2
o = ob.OracleDense(W, H, num_threads=8)
for f in synth(seed, W, H, T):
    o.push(f)
o.flush()
