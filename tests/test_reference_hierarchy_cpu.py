"""Golden vectors for the NEXT row of SURVEY section 8f (N1, hierarchical region merge), produced by the reference's own
two-stage pipeline (DenseSegmentation -> RegionSegmentation) compiled into oracle/_ref/libref_hier.so (oracle/Makefile:
one documented build-time edit; 8-bit Lab = the oracle's cv2-identical restatement).  Nothing in the product implements
this stage yet; these tests keep the vectors honest (the library reproduces them, they are well-formed hierarchies, and
their base level IS the dense stage the product already matches) so that the hierarchical oracle and kernels can be
built against them."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import reference_cases as rc
import reference_hierarchy as rh
from helpers import partition_equal

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_hierarchy.json")


def _need_lib():
    if not rh.available():
        pytest.skip("oracle/_ref/libref_hier.so not built (needs /root/reference)")


@pytest.mark.parametrize("case", sorted(rh.CASES))
def test_compiled_reference_reproduces_hierarchy_golden(case):
    _need_lib()
    gold = json.load(open(GOLD))[case]
    recs, batches = rh.run_case(case)
    first = rh.first_chunk_set(recs)
    assert len(recs) == gold["frames"] and batches == gold["batches"]
    assert rh.digest(first) == gold["sha256_first_chunk_set"]
    frames = [rh.parse(r) for r in first]
    assert [len(f["region_id"]) for f in frames] == gold["regions_per_frame"]
    hier = [f for f in frames if f["levels"]]
    assert [[len(l) for l in f["levels"]] for f in hier] == [h["level_region_counts"] for h in gold["hierarchies"]]


@pytest.mark.parametrize("case", ["real_one_chunk_set", "real_coarse_levels"])
def test_hierarchy_golden_is_a_tree_over_the_dense_stage(case):
    """Each level's regions partition the level below through child / parent ids with sizes adding up, level counts
    shrink by the cut-off fraction down to min_region_num, and the base of the hierarchy is the dense stage: the
    over-segmentation in the hierarchical results equals the oracle's dense stream up to the ids the region stage
    re-assigns."""
    _need_lib()
    n, _, dense_chunk, _, _, min_regions, cutoff = rh.CASES[case]
    recs, _ = rh.run_case(case)
    frames = [rh.parse(r) for r in recs]
    levels = frames[0]["levels"]
    assert all(not f["levels"] for f in frames[1:])                       # one chunk set: hierarchy on its first frame
    assert set(int(i) for f in frames for i in f["region_id"]) == set(c["id"] for c in levels[0])
    assert sum(c["size"] for c in levels[0]) == frames[0]["width"] * frames[0]["height"] * n
    for k in range(1, len(levels)):
        below = {c["id"]: c for c in levels[k - 1]}
        assert sorted(x for c in levels[k] for x in c["children"]) == sorted(below)
        for c in levels[k]:
            assert c["size"] == sum(below[x]["size"] for x in c["children"])
            assert all(below[x]["parent_id"] == c["id"] for x in c["children"])
            assert c["start_frame"] == min(below[x]["start_frame"] for x in c["children"])
            assert c["end_frame"] == max(below[x]["end_frame"] for x in c["children"])
        assert len(levels[k]) <= int(np.ceil(len(levels[k - 1]) * cutoff)) + 1
    assert len(levels[-1]) <= max(min_regions, int(len(levels[-2]) * cutoff) + 1) and all(c["parent_id"] == -1 for c in levels[-1])
    clip = np.load(os.path.join(os.path.dirname(GOLD), "real_clip_136x240x24.npz"))["frames"][:n]
    dense = rc.run_stream(ob.OracleDense, clip, None, dict(chunk_size=dense_chunk))
    for f, d in zip(frames, dense):
        assert partition_equal(ob.id_map_from_result(f), ob.id_map_from_result(d))
