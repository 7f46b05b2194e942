"""Writes tests/golden/reference_results.json: digests of the per-frame results of the REFERENCE ITSELF (its own
DenseSegmentation pipeline compiled unmodified into oracle/_ref/libref_results.so, see oracle/Makefile) on the cases of
tests/reference_cases.py.  Run in the container that mounts /root/reference:

    make -C oracle _ref && python tests/golden/make_reference_golden.py

tests/test_oracle_cpu.py::test_oracle_matches_reference_golden holds the oracle to these digests wherever it runs;
test_oracle_equals_compiled_reference additionally compares field by field where the library is present."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import reference_binding as rb   # noqa: E402
import reference_cases as rc     # noqa: E402

assert rb.available(), "oracle/_ref/libref_results.so missing: run `make -C oracle _ref` where /root/reference is mounted"
out = {"_source": "videosegmentation/video_segment @ c930c455, DenseSegmentation::ProcessFrame compiled unmodified (oracle/Makefile target _ref/libref_results.so)"}
for name in rc.CASES:
    clip, flows, opts = rc.load_case(name)
    res = rc.run_stream(rb.ReferenceDense, clip, flows, opts)
    out[name] = {"sha256": rc.digest(res), "frames": len(res), "regions_per_frame": [int(r["region_id"].size) for r in res],
                 "compound_regions_first_frame": int(res[0]["compound"].shape[0])}
    print(name, out[name]["sha256"][:16], out[name]["regions_per_frame"][:4])
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "reference_results.json"), "w"), indent=1)
