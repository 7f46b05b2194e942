"""2-GPU test of the sharded run (-m gpu, needs two devices; skipped on one): one clip segmented by ONE handle and by
TWO ranks (frame groups with one shared frame, csrc/shard.cu: NCCL send/recv of the overlap id maps, all-gather of
the id counts, vote + relabel on the device).  Bars: the predecessor's chunks are identical to the single run's, ids
included; ids are consistent across the seam wherever the two partitions of the shared frame agree (measured on a B200
pair: 80.6 % of its pixels carry the same id on both sides; >= 75 % asserted); ids born after the seam never collide.
The successor's chain is an independent over-segmentation of its group (it never sees the predecessor's map -- that is
what makes the seam parallel), so its partitions are compared with the sequential chain's for the record only
(measured IoU 0.71-0.75 on this clip, >= 0.65 asserted as a regression guard); the seam that reproduces the sequential
chain bit for bit is the pipelined one (test_seam_import_halo_continues_the_chain_exactly).  The successor starts one
chunk early and discards it (warm-up overlap), so that its first kept chunk is constrained like every chunk of a chain
(seam id agreement 75.4 % without the warm-up, 80.6 % with it)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))

WORKER = r'''
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.join(r"%(root)s")); sys.path.insert(0, os.path.join(r"%(root)s", "tests"))
import torch
from video_segment_b200.synth import synth
from video_segment_b200.unit import DenseSegmentationOptions, DenseSegmentationUnit
from video_segment_b200.shard import SeamLink, group_range, nccl_unique_id, relabel_results
rank, world, idfile, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
W, H, L = 320, 240, 28                      # 28 frames per group: one 10-frame chunk + two 9-frame constrained chunks; groups share one frame
torch.cuda.set_device(rank)
if rank == 0:
    uid = nccl_unique_id()
    open(idfile + ".tmp", "wb").write(uid); os.rename(idfile + ".tmp", idfile)
else:
    import time
    while not os.path.exists(idfile): time.sleep(0.05)
    uid = open(idfile, "rb").read()
link = SeamLink(uid, rank, world, rank, W, H)
start, count = group_range(rank, world, L)
# a successor group starts one chunk (9 frames) early and discards that chunk: its first kept chunk is then constrained
# by its own history, like every chunk of the sequential chain, instead of starting free
WARM = 9 if rank > 0 else 0
frames = list(synth(11, W, H, count + WARM, start=start - WARM))
u = DenseSegmentationUnit(dense_seg_options=DenseSegmentationOptions(chunk_size=10), want_id_maps=True, device=rank)
assert u.open_streams(W, H)
res = []
for f in frames: res += u.process_frame(f)
# group boundary: the unit stands right after its third chunk boundary (28 = 10 + 9 + 9 frames pushed)
offsets = link.exchange(u)
res += u.post_process()
res = res[WARM:]
n_ids = max(int(r["region_id"].max()) for r in res) + 1
first = torch.from_numpy(np.ascontiguousarray(res[0]["id_map"])).cuda()
table = link.relabel_table(first.data_ptr(), n_ids, offsets[rank])
if rank > 0:
    relabel_results(res, table)
np.savez(out, maps=np.stack([r["id_map"] for r in res]), offsets=np.asarray(offsets), stats=json.dumps(link.stats()))
link.close(); u.close()
'''


def test_two_rank_sharded_run_matches_single_run(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from helpers import overseg_iou
    from video_segment_b200.synth import synth
    from video_segment_b200.unit import DenseSegmentationOptions, DenseSegmentationUnit
    root = os.path.dirname(HERE)
    script = tmp_path / "worker.py"
    script.write_text(WORKER % dict(root=root))
    idfile = str(tmp_path / "nccl_id")
    procs = [subprocess.Popen([sys.executable, str(script), str(r), "2", idfile, str(tmp_path / f"rank{r}.npz")]) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    a = np.load(tmp_path / "rank0.npz")
    b = np.load(tmp_path / "rank1.npz")
    W, H, L = 320, 240, 28
    # single handle over the whole clip (2 * 28 - 1 = 55 frames)
    u = DenseSegmentationUnit(dense_seg_options=DenseSegmentationOptions(chunk_size=10), want_id_maps=True)
    assert u.open_streams(W, H)
    single = []
    for f in synth(11, W, H, 2 * L - 1):
        single += u.process_frame(f)
    single += u.post_process()
    u.close()
    single = np.stack([r["id_map"] for r in single])
    assert len(a["maps"]) == L and len(b["maps"]) == L
    # rank 0's chunks that closed before its group ended are the single run's, ids included (its flush differs)
    assert np.array_equal(a["maps"][:18], single[:18])
    # ids consistent across the seam: the shared frame (rank 0's last, rank 1's first)
    same = float(np.mean(a["maps"][L - 1] == b["maps"][0]))
    ious = [overseg_iou(single[L - 1 + k], b["maps"][k]) for k in range(L)]
    report = {"seam_id_agreement": same, "successor_iou_min": float(min(ious)), "successor_iou_mean": float(np.mean(ious)),
              "successor_iou_first6": [round(float(x), 4) for x in ious[:6]], "link0": str(a["stats"]), "link1": str(b["stats"])}
    print("sharded run vs single handle:", json.dumps(report))
    os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
    with open(os.path.join(root, "gpurun_out", "r02_shard_2gpu_test.json"), "w") as f:
        json.dump(report, f)
    # The successor's partition of the shared frame is its own (a parallel seam: it never saw the predecessor's map), so
    # agreement is bounded by how much the two partitions differ; what must hold is that the ids it shares are the
    # predecessor's, the rest are fresh, and its partitions stay close to the sequential chain's.
    assert same >= 0.75
    assert min(ious) >= 0.65
    assert int(b["offsets"][1]) > 0 and int(a["offsets"][0]) == 0
    # ids the successor creates later do not collide with the predecessor's
    born_later = set(np.unique(b["maps"][-1]).tolist()) - set(np.unique(a["maps"]).tolist())
    assert all(x >= int(b["offsets"][1]) for x in born_later)
