"""Dev tool (CPU): runs the oracle with and without the stage-0 simulation (oracle/vso_graph.cpp,
VSO_SIM_STAGE0) on a synthetic clip and compares every frame's partition.
usage: python tools/sim_stage0.py W H T [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import oracle_binding as ob
from helpers import partition_equal, overseg_iou
from video_segment_b200.synth import synth

W, H, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
seed = int(sys.argv[4]) if len(sys.argv) > 4 else 2
frames = list(synth(seed, W, H, T))

def run(sim):
    if sim:
        os.environ[os.environ.get("SIM_VAR", "VSO_SIM_STAGE0")] = "1"
    else:
        os.environ.pop(os.environ.get("SIM_VAR", "VSO_SIM_STAGE0"), None)
    o = ob.OracleDense(W, H, num_threads=8)
    out = []
    t0 = time.time()
    for f in frames:
        out += o.push(f)
    out += o.flush()
    print("sim" if sim else "ref", "%.1fs" % (time.time() - t0), "stage s", [round(x, 2) for x in o.stage_seconds()], flush=True)
    return out

a = run(False)
b = run(True)
bad = 0
for i, (x, y) in enumerate(zip(a, b)):
    ma, mb = ob.id_map_from_result(x), ob.id_map_from_result(y)
    eq = partition_equal(ma, mb)
    same_ids = np.array_equal(ma, mb)
    if not eq:
        bad += 1
        print("frame", i, "chunk", x["chunk_id"], "regions", x["n_regions"] if "n_regions" in x else len(x["region_id"]), len(y["region_id"]), "iou", overseg_iou(ma, mb))
    elif not same_ids:
        print("frame", i, "partition equal, ids differ")
print("frames", len(a), "partition mismatches", bad, "regions/frame", [len(x["region_id"]) for x in a][::5])
