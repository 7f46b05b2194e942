"""GPU parity of the hierarchical region stage (-m gpu): DenseSegmentationUnit -> RegionSegmentationUnit (both through
the C ABI) against the oracle's two stages (oracle/vso_engine.cpp -> oracle/vso_hier.cpp, the latter pinned word for
word by the compiled reference, tests/test_hier_oracle_cpu.py).
What is exact: stream shape, over-segmentation rasters and ids, level sizes, tree structure, region sizes.  What is
held to a tolerance: WHICH regions merge -- descriptor sums run in another order on the device, so an edge
next to a bucket boundary may merge earlier or later; the partitions of every level are compared by IoU."""
import os

import numpy as np
import pytest

import oracle_binding as ob
import reference_hierarchy as rh
from helpers import overseg_iou

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _streams(case):
    n, flow_sigma, dense_chunk, set_size, set_overlap, min_regions, cutoff = rh.CASES[case]
    clip = np.load(os.path.join(GOLDEN, "real_clip_136x240x24.npz"))["frames"][:n]
    t, h, w, _ = clip.shape
    flows = None
    if flow_sigma is not None:
        flows = np.random.default_rng(3).normal(0, flow_sigma, (t, h, w, 2)).astype(np.float32)
    return clip, flows, dict(dense_chunk=dense_chunk, chunk_set_size=set_size, chunk_set_overlap=set_overlap,
                             min_region_num=min_regions, level_cutoff_fraction=cutoff)


def _oracle_dense(clip, flows, p):
    """The over-segmentation stream both region stages are fed with (the oracle's dense stage: identical input, so that
    the comparison isolates the region stage; the product's own dense stage is covered by tests/test_gpu_engine.py)."""
    h, w = clip[0].shape[:2]
    dense = ob.OracleDense(w, h, use_flow=flows is not None, chunk_size=p["dense_chunk"], num_threads=8)
    out = []
    for k, f in enumerate(clip):
        out += dense.push(f, None if flows is None or k == 0 else flows[k])
    out += dense.flush()
    dense.close()
    return out


def _run_gpu(overseg, clip, flows, p):
    from video_segment_b200.unit import RegionSegmentationOptions, RegionSegmentationUnit
    h, w = clip[0].shape[:2]
    region = RegionSegmentationUnit(RegionSegmentationOptions(chunk_set_size=p["chunk_set_size"], chunk_set_overlap=p["chunk_set_overlap"],
                                                              min_region_num=p["min_region_num"], level_cutoff_fraction=p["level_cutoff_fraction"]))
    assert region.open_streams(w, h, flow_stream_present=flows is not None)
    out = []
    for k, r in enumerate(overseg):
        out += region.process_frame(r, clip[k], None if flows is None or k == 0 else flows[k])
    out += region.post_process()
    st = region.stats()
    region.close()
    return out, st


def _run_oracle(overseg, clip, flows, p):
    h, w = clip[0].shape[:2]
    hier = ob.OracleHier(w, h, use_flow=flows is not None, chunk_set_size=p["chunk_set_size"], chunk_set_overlap=p["chunk_set_overlap"],
                         min_region_num=p["min_region_num"], level_cutoff_fraction=p["level_cutoff_fraction"])
    out = []
    for k, r in enumerate(overseg):
        out += hier.push(r, clip[k], None if flows is None or k == 0 else flows[k])
    out += hier.flush()
    return [rh.parse(r) for r in out]


def _level_maps(frames, level):
    """Label volume (T, H, W) of hierarchy level `level` of a chunk set: every over-segmentation region carries the id of
    its ancestor at that level."""
    levels = frames[0]["levels"]
    up = {c["id"]: c["id"] for c in levels[0]}
    for l in range(1, level + 1):
        parent = {c["id"]: c["parent_id"] for c in levels[l - 1]}
        up = {k: parent[v] for k, v in up.items()}
    vol = []
    for f in frames:
        img = ob.id_map_from_result(f)
        lut = np.full(max(up) + 2, -1, np.int64)
        for k, v in up.items():
            lut[k] = v
        vol.append(lut[img])
    return np.stack(vol)


@pytest.mark.parametrize("case", ["real_one_chunk_set", "real_flow_one_chunk_set", "real_coarse_levels"])
def test_region_stage_matches_oracle(case):
    clip, flows, p = _streams(case)
    overseg = _oracle_dense(clip, flows, p)
    got, st = _run_gpu(overseg, clip, flows, p)
    ref = _run_oracle(overseg, clip, flows, p)
    assert len(got) == len(ref) and st["kernel_launches"] > 0
    for g, r in zip(got, ref):
        for k in ("width", "height", "chunk_id", "chunk_size", "overlap_start", "hierarchy_frame_idx"):
            assert g[k] == r[k], k
        # the over-segmentation under the region stage's ids: identical rasters, ids and shape moments
        assert np.array_equal(g["region_id"], r["region_id"])
        assert np.array_equal(g["intervals"], r["intervals"]) and np.array_equal(g["interval_offset"], r["interval_offset"])
        assert np.array_equal(g["shape_moments"], r["shape_moments"])
        assert len(g["levels"]) == len(r["levels"])
    gl, rl = got[0]["levels"], ref[0]["levels"]
    assert [len(l) for l in gl] == [len(l) for l in rl]                   # level sizes follow the cut-off fraction exactly
    assert [(c["id"], c["size"], c["start_frame"], c["end_frame"], c["neighbors"]) for c in gl[0]] == \
           [(c["id"], c["size"], c["start_frame"], c["end_frame"], c["neighbors"]) for c in rl[0]]
    for k in range(1, len(gl)):                                           # a tree with consistent sizes and frame bounds
        below = {c["id"]: c for c in gl[k - 1]}
        assert sorted(x for c in gl[k] for x in c["children"]) == sorted(below)
        for c in gl[k]:
            assert c["size"] == sum(below[x]["size"] for x in c["children"])
            assert all(below[x]["parent_id"] == c["id"] for x in c["children"])
            assert c["start_frame"] == min(below[x]["start_frame"] for x in c["children"])
            assert c["end_frame"] == max(below[x]["end_frame"] for x in c["children"])
    # which regions merged: partitions per level against the oracle
    ious, same = [], 0
    for level in range(1, len(gl)):
        a, b = _level_maps(ref, level), _level_maps(got, level)
        ious.append(overseg_iou(a, b))
        same += int(sorted(tuple(sorted(c["children"])) for c in gl[level]) == sorted(tuple(sorted(c["children"])) for c in rl[level]))
    print(case, "levels", len(gl), "identical levels", same, "IoU per level", [round(x, 4) for x in ious])
    assert ious[0] >= 0.95 and np.mean(ious) >= 0.85, ious


def test_region_stage_errors():
    from video_segment_b200.unit import RegionSegmentationOptions, RegionSegmentationUnit
    assert not RegionSegmentationUnit(RegionSegmentationOptions(chunk_set_size=1)).open_streams(64, 48)
    assert not RegionSegmentationUnit(RegionSegmentationOptions(chunk_set_overlap=6)).open_streams(64, 48)
    assert not RegionSegmentationUnit(RegionSegmentationOptions(compute_vectorization=True)).open_streams(64, 48)
    assert not RegionSegmentationUnit().open_streams(64, 48, pixel_format="RGB24")
    u = RegionSegmentationUnit()
    assert u.open_streams(64, 48)
    assert u.post_process() == []
    u.close()


def test_dense_and_region_units_chain_end_to_end():
    """DenseSegmentationUnit -> RegionSegmentationUnit, both on the device path, on a synthetic clip with flow: stream
    shape, trees, and the base of the hierarchy is the dense stage's partition."""
    from helpers import partition_equal
    from video_segment_b200.synth import synth_flow
    from video_segment_b200.unit import (DenseSegmentationOptions, DenseSegmentationUnit, RegionSegmentationOptions,
                                         RegionSegmentationUnit, id_map_from_result)
    pairs = list(synth_flow(5, 192, 128, 34))
    clip = [p[0] for p in pairs]
    flows = [p[1] for p in pairs]
    dense = DenseSegmentationUnit(dense_seg_options=DenseSegmentationOptions(chunk_size=8), want_id_maps=True)
    assert dense.open_streams(192, 128, flow_stream_present=True)
    region = RegionSegmentationUnit(RegionSegmentationOptions(chunk_set_size=3, chunk_set_overlap=1))
    assert region.open_streams(192, 128, flow_stream_present=True)
    out, dense_maps, fed = [], [], [0]

    def feed(results):
        got = []
        for r in results:
            k = fed[0]
            dense_maps.append(r["id_map"])
            got += region.process_frame(r, clip[k], None if k == 0 else flows[k])
            fed[0] += 1
        return got

    for k, f in enumerate(clip):
        out += feed(dense.process_frame(f, flows[k]))
    out += feed(dense.post_process())
    out += region.post_process()
    dense.close(); region.close()
    assert len(out) == len(clip)
    sets = sorted(set(f["hierarchy_frame_idx"] for f in out))
    assert len(sets) >= 2
    for f, dm in zip(out, dense_maps):
        assert partition_equal(ob.id_map_from_result(f), dm)
    for f in out:
        levels = f["levels"]
        for k in range(1, len(levels)):
            below = {c["id"]: c for c in levels[k - 1]}
            assert sorted(x for c in levels[k] for x in c["children"]) == sorted(below)
            for c in levels[k]:
                assert c["size"] == sum(below[x]["size"] for x in c["children"])
