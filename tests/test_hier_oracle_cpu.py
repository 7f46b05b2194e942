"""The oracle's hierarchical region stage (oracle/vso_hier.cpp) against the REFERENCE's own two stages compiled into
oracle/_ref/libref_hier.so (DenseSegmentation -> RegionSegmentation), word for word on the flat result records, and
against the committed golden digests of those records (tests/golden/reference_hierarchy.json) where /root/reference is
absent.  First chunk set of every case: the reference's constrained chunk sets are not run-to-run deterministic
(tests/reference_hierarchy.py)."""
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
import reference_hierarchy as rh

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_hierarchy.json")


def run_oracle_case(name):
    n, flow_sigma, dense_chunk, set_size, set_overlap, min_regions, cutoff = rh.CASES[name]
    clip = np.load(os.path.join(os.path.dirname(GOLD), "real_clip_136x240x24.npz"))["frames"][:n]
    t, h, w, _ = clip.shape
    flows = None
    if flow_sigma is not None:
        flows = np.random.default_rng(3).normal(0, flow_sigma, (t, h, w, 2)).astype(np.float32)
    dense = ob.OracleDense(w, h, use_flow=flows is not None, chunk_size=dense_chunk, num_threads=4)
    hier = ob.OracleHier(w, h, use_flow=flows is not None, chunk_set_size=set_size, chunk_set_overlap=set_overlap,
                         min_region_num=min_regions, level_cutoff_fraction=cutoff)
    out, batches, fed = [], [], [0]

    def feed(results):
        got = []
        for r in results:
            k = fed[0]
            fl = None if flows is None or k == 0 else flows[k]
            got += hier.push(r, clip[k], fl)
            fed[0] += 1
        return got

    for k, f in enumerate(clip):
        fl = None if flows is None or k == 0 else flows[k]
        got = feed(dense.push(f, fl))
        batches.append(len(got))
        out += got
    got = feed(dense.flush())
    got += hier.flush()
    batches.append(len(got))
    out += got
    return out, batches


# "real_two_chunk_sets" (chunk sets of two chunks with one chunk of overlap): the reference's chunk arithmetic makes the
# very first dense chunk an overlap chunk, so the first records it outputs already come from a CONSTRAINED set -- where
# its AddEdge reads a not yet constructed Region (uninitialised heap, see oracle/vso_hier.cpp) and the result varies from
# process to process.  That case is held to the structural properties only (test_hier_oracle_constrained_sets_are_consistent).
DETERMINISTIC = [c for c in sorted(rh.CASES) if c != "real_two_chunk_sets"]


@pytest.mark.parametrize("case", DETERMINISTIC)
def test_hier_oracle_matches_reference_golden(case):
    gold = json.load(open(GOLD))[case]
    recs, batches = run_oracle_case(case)
    assert len(recs) == gold["frames"] and batches == gold["batches"]
    first = rh.first_chunk_set(recs)
    frames = [rh.parse(r) for r in first]
    assert [len(f["region_id"]) for f in frames] == gold["regions_per_frame"]
    hier = [f for f in frames if f["levels"]]
    assert [[len(l) for l in f["levels"]] for f in hier] == [h["level_region_counts"] for h in gold["hierarchies"]]
    assert rh.digest(first) == gold["sha256_first_chunk_set"]


@pytest.mark.parametrize("case", DETERMINISTIC)
def test_hier_oracle_equals_compiled_reference(case):
    if not rh.available():
        pytest.skip("oracle/_ref/libref_hier.so not built (needs /root/reference)")
    ref, ref_batches = rh.run_case(case)
    got, batches = run_oracle_case(case)
    assert batches == ref_batches and len(got) == len(ref)
    for k, (a, b) in enumerate(zip(rh.first_chunk_set(got), rh.first_chunk_set(ref))):
        assert len(a) == len(b) and np.array_equal(a, b), (case, k, rh.parse(a)["levels"][:1] if len(a) == len(b) else (len(a), len(b)))


def test_hier_oracle_constrained_sets_are_consistent():
    """Chunk sets constrained by their predecessor: same stream shape as the reference (frames, batches), every
    hierarchy a tree over the over-segmentation, region ids persistent across the set boundary."""
    case = "real_two_chunk_sets"
    gold = json.load(open(GOLD))[case]
    recs, batches = run_oracle_case(case)
    assert len(recs) == gold["frames"] and batches == gold["batches"]
    frames = [rh.parse(r) for r in recs]
    assert [len(f["region_id"]) for f in frames[:len(gold["regions_per_frame"])]] == gold["regions_per_frame"]
    for f in frames:
        levels = f["levels"]
        for k in range(1, len(levels)):
            below = {c["id"]: c for c in levels[k - 1]}
            assert sorted(x for c in levels[k] for x in c["children"]) == sorted(below)
            for c in levels[k]:
                assert c["size"] == sum(below[x]["size"] for x in c["children"])
                assert all(below[x]["parent_id"] == c["id"] for x in c["children"])
    sets = sorted(set(f["hierarchy_frame_idx"] for f in frames))
    assert len(sets) >= 2
    a = [f for f in frames if f["hierarchy_frame_idx"] == sets[0]][-1]
    b = [f for f in frames if f["hierarchy_frame_idx"] == sets[1]][0]
    shared = set(int(i) for i in a["region_id"]) & set(int(i) for i in b["region_id"])
    assert len(shared) > 0.3 * len(a["region_id"])          # constrained ids carry over the boundary (the compiled reference: 111 of 266)
