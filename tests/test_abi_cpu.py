"""CPU tests (-m "not gpu"): the C-ABI library loads and exports every symbol include/vsb200.h
declares; without a GPU the product refuses to run (no CPU fallback); host-side unit logic."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from video_segment_b200._lib import EXPORTED_SYMBOLS, LIB_PATH, lib
    hdr = open(os.path.join(ROOT, "include", "vsb200.h")).read()
    declared = sorted(set(re.findall(r"\b(vsb200_[a-z0-9_]+)\s*\(", hdr)))
    assert declared, "no declarations found"
    assert sorted(EXPORTED_SYMBOLS) == declared
    L = lib()
    raw = C.CDLL(LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert L.vsb200_bucket_index(0.0) == 0 and L.vsb200_bucket_index(1.0) == 2047 and L.vsb200_bucket_index(1e10) == 2048


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from video_segment_b200._lib import lib
    from video_segment_b200.unit import DenseSegmentationUnit
    assert lib().vsb200_device_count() == 0
    u = DenseSegmentationUnit()
    assert not u.open_streams(64, 48)          # fails loudly (status + message), never computes on the CPU
    assert b"no CPU fallback" in lib().vsb200_last_error()
    with pytest.raises(RuntimeError):
        u.process_frame(np.zeros((48, 64, 3), np.uint8))
    # the region stage and the kernel-level entries refuse as well
    from video_segment_b200.unit import RegionSegmentationUnit
    assert not RegionSegmentationUnit().open_streams(64, 48)
    import ctypes as C
    n = C.c_int(-1)
    rc = lib().vsb200_label_components(None, 64, 48, 1, None, None, 0, C.byref(n), None)
    assert rc != 0 and n.value == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "video_segment_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle_binding" not in txt and "libvso" not in txt and "vso.h" not in txt, f


def test_default_opts_match_reference_defaults():
    from video_segment_b200._lib import DenseOpts, lib
    o = DenseOpts()
    lib().vsb200_dense_default_opts(C.byref(o))
    assert (o.presmoothing, o.chunk_size, o.num_constraint_frames) == (2, 20, 1)
    assert abs(o.frac_min_region_size - 0.01) < 1e-9 and abs(o.chunk_overlap_ratio - 0.2) < 1e-7
    assert (o.enforce_n4_connectivity, o.enforce_spatial_connectedness, o.color_distance) == (1, 1, 1)


def test_synth_is_deterministic():
    from video_segment_b200.synth import synth_clip
    a = synth_clip(3, 80, 60, 3)
    b = synth_clip(3, 80, 60, 3)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (3, 60, 80, 3)
    assert not np.array_equal(a[0], a[1])


def test_cpp_example_refuses_to_run_without_gpu(tmp_path):
    """examples/over_segment_b200.cpp (the dense half of seg_tree_sample in C++ over B200DenseSegmentation + the C ABI's
    container writer, built against the reference's headers by `make -C oracle _ref`): without an sm_100 device it
    aborts with the C ABI's 'no CPU fallback' message instead of computing anything."""
    import struct
    import subprocess
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "over_segment_b200")
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/over_segment_b200 not built (needs /root/reference)")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    clip = np.load(os.path.join(ROOT, "tests", "golden", "real_clip_136x240x24.npz"))["frames"][:3]
    t, h, w, _ = clip.shape
    src = tmp_path / "in.bgr"
    src.write_bytes(struct.pack("<iii", w, h, t) + clip.tobytes())
    p = subprocess.run([exe, str(src), str(tmp_path / "out.pb")], capture_output=True, text=True, timeout=120)
    assert p.returncode != 0
    assert "no CPU fallback" in p.stderr


def test_reference_arm_prints_one_line_under_torchrun():
    """bench.py --impl reference launched like the driver launches it for N > 1: rank 0 alone runs the CPU reference
    and prints ONE JSON line with the contract's keys, the other rank exits 0 without work."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29613", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0", "--width", "160", "--height", "120"]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-800:]
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["unit"] == "frames/s" and line["value"] > 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["higher_is_better"] is True and "160x120" in line["config"]["workload"]
    assert len(line["ms_per_step_series"]) == 1


def _units_check_input(tmp_path, frames, flows=None):
    import struct
    t, h, w, _ = frames.shape
    src = tmp_path / "in.bgr"
    blob = struct.pack("<iiii", w, h, t, 1 if flows is not None else 0) + frames.tobytes()
    if flows is not None:
        blob += np.ascontiguousarray(flows, np.float32).tobytes()
    src.write_bytes(blob)
    return src


def test_cpp_video_units_build_against_reference_framework_and_fail_loudly_without_gpu(tmp_path):
    """oracle/_ref/b200_units_check = the product's B200DenseSegmentationUnit / B200RegionSegmentationUnit compiled
    against the reference's video_framework (video_unit.cpp unmodified) and run as the tree seg_tree_sample builds.
    Without a GPU OpenStreams must return false (no CPU fallback) and the pipeline must not run."""
    import subprocess
    import torch
    exe = os.path.join(ROOT, "oracle", "_ref", "b200_units_check")
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200_units_check not built (needs /root/reference)")
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    clip = np.load(os.path.join(ROOT, "tests", "golden", "real_clip_136x240x24.npz"))["frames"][:3]
    src = _units_check_input(tmp_path, clip)
    p = subprocess.run([exe, str(src), str(tmp_path / "out.bin")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 1
    assert "no CPU fallback" in p.stderr
