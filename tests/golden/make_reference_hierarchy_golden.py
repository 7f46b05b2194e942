"""Writes tests/golden/reference_hierarchy.json: what the REFERENCE's two-stage pipeline (dense over-segmentation ->
hierarchical region segmentation) produces on the cases of tests/reference_hierarchy.py, from
oracle/_ref/libref_hier.so (oracle/Makefile: the reference's sources with one documented build-time edit, 8-bit Lab =
the oracle's cv2-identical restatement).  These are the golden vectors the hierarchical merge is built
against next: per-level region counts and sizes in the clear, everything else (ids, rasters, parents, children,
neighbours, frame ranges) under a SHA-256 of the flat records.

    make -C oracle _ref && python tests/golden/make_reference_hierarchy_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import reference_hierarchy as rh   # noqa: E402

assert rh.available(), "oracle/_ref/libref_hier.so missing"
out = {"_source": "videosegmentation/video_segment @ c930c455: DenseSegmentation -> RegionSegmentation, oracle/Makefile target _ref/libref_hier.so"}
for name in rh.CASES:
    recs, batches = rh.run_case(name)
    first = rh.first_chunk_set(recs)
    frames = [rh.parse(r) for r in first]
    out[name] = {
        "sha256_first_chunk_set": rh.digest(first), "frames": len(recs), "frames_first_chunk_set": len(first), "batches": batches,
        "regions_per_frame": [int(len(f["region_id"])) for f in frames],
        "hierarchies": [{"frame": k, "chunk_id": f["chunk_id"], "level_region_counts": [len(l) for l in f["levels"]],
                         "top_level_sizes": sorted((c["size"] for c in f["levels"][-1]), reverse=True)}
                        for k, f in enumerate(frames) if f["levels"]],
    }
    print(name, out[name]["sha256_first_chunk_set"][:16], batches[-3:], [h["level_region_counts"][:4] for h in out[name]["hierarchies"]])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "reference_hierarchy.json"), "w"), indent=1)
