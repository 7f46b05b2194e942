"""Runs segmentation::B200DenseSegmentation (the C++ host class, through tests/host_check_wrap.cpp) on a clip of
tests/reference_cases.py and compares its SegmentationDesc stream with the compiled reference's, field by field.
A separate process because the class reports setup / device errors like the reference does, through CHECK -> abort.
Usage: python tests/gpu_host_class_probe.py <case>   (needs a B200; prints OK or the first difference)"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import reference_binding as rb   # noqa: E402
import reference_cases as rc     # noqa: E402


def main(case):
    clip, flows, opts = rc.load_case(case)
    ref = rc.run_stream(rb.ReferenceDense, clip, flows, opts)
    h, w = clip[0].shape[:2]
    e = rb.B200HostDense(w, h, use_flow=flows is not None, **opts)
    got = []
    for k, f in enumerate(clip):
        got += e.push(f, None if flows is None or k == 0 else flows[k])
    got += e.flush()
    launches = e.kernel_launches()
    e.close()
    diff = rc.first_difference(ref, got)
    if launches <= 0:
        diff = "no kernel launches"
    print("OK" if diff is None else f"DIFF {diff}", launches, flush=True)
    return 0 if diff is None else 1


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "real_single_chunk"))
