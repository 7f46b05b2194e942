"""GPU parity (-m gpu) against the REFERENCE ITSELF: the CUDA path behind DenseSegmentationUnit versus the
reference's own DenseSegmentation pipeline compiled unmodified into oracle/_ref/libref_results.so (built where
/root/reference is mounted, shipped to the GPU box as a prebuilt file; nothing here reads /root/reference)."""
import numpy as np
import pytest

import os

import oracle_binding as ob
import reference_binding as rb
import reference_cases as rc
from helpers import overseg_iou
from test_gpu_engine import _run_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", ["real_default_chunk20", "real_single_chunk", "real_l1_single_chunk_plain"])
def test_stream_matches_compiled_reference(case):
    """Header fields identical on every frame, region-id maps within the IoU bar (BASELINE.json: >= 0.99 up to a label
    permutation) on every frame; on the frames of the first (unconstrained) chunk the partition is the reference's, and
    with it the region ids, scan intervals, shape moments, hierarchy level 0 and neighbour lists."""
    if not rb.available(build=False):
        pytest.skip("oracle/_ref/libref_results.so not shipped")
    clip, flows, opts = rc.load_case(case)
    ref = rc.run_stream(rb.ReferenceDense, clip, flows, opts)
    kw = {k: (bool(v) if k.startswith("enforce") else v) for k, v in opts.items()}
    got, batches, st = _run_gpu(clip, flows, **kw)
    assert st["kernel_launches"] > 0
    assert len(got) == len(ref)
    ious = []
    for t, (g, r) in enumerate(zip(got, ref)):
        for k in ("width", "height", "chunk_id", "chunk_size", "overlap_start", "hierarchy_frame_idx", "connectedness"):
            assert g[k] == r[k], (t, k)
        ious.append(overseg_iou(ob.id_map_from_result(r), g["id_map"]))
    assert min(ious) >= 0.99, ious
    if flows is None:
        first = ref[0]["chunk_size"]            # frames output by chunk 0
        for g, r in list(zip(got, ref))[:first]:
            assert np.array_equal(g["region_id"], r["region_id"])
            assert np.array_equal(g["intervals"], r["intervals"]) and np.array_equal(g["interval_offset"], r["interval_offset"])
            assert np.array_equal(g["shape_moments"], r["shape_moments"])
        assert np.array_equal(got[0]["compound"], ref[0]["compound"])
        assert np.array_equal(got[0]["neighbor_offset"], ref[0]["neighbor_offset"])
        assert np.array_equal(got[0]["neighbor_id"], ref[0]["neighbor_id"])


def test_segmentation_file_written_from_the_engine(tmp_path, real_clip):
    """Result container (csrc/pb_io.cu): frames appended from the engine's wire encoder as they are popped; the file
    read back holds exactly the per-frame protobuf bytes and pts, and the reference's own reader (where shipped)
    reads the same."""
    from proto_schema import segmentation_desc_class
    from video_segment_b200.segio import SegmentationReader, SegmentationWriter
    from video_segment_b200.unit import DenseSegmentationUnit
    clip = real_clip[:8]
    h, w = clip[0].shape[:2]
    path = str(tmp_path / "seg.pb")
    writer = SegmentationWriter(path)
    assert writer.open_file([1, 0])
    u = DenseSegmentationUnit(want_proto=True)
    u.segmentation_writer = writer
    assert u.open_streams(w, h)
    got = []
    for k, f in enumerate(clip):
        got += u.process_frame(f, pts=1000 * k)
    got += u.post_process()
    u.close()
    writer.write_term_header_and_close()
    r = SegmentationReader(path)
    assert r.open_file_and_read_headers()
    assert r.get_header_flags() == [1, 0] and r.num_frames() == len(clip)
    assert r.time_stamps() == [1000 * k for k in range(len(clip))]
    frames = [r.read_next_frame_binary() for _ in range(len(clip))]
    r.close_file()
    assert frames == [g["proto"] for g in got]
    m = segmentation_desc_class()()
    m.ParseFromString(frames[0])
    assert m.frame_width == w and m.frame_height == h and [x.id for x in m.region] == list(got[0]["region_id"])
    if rb.host_available(build=False):
        assert rb.ref_io_read(path) == ([1, 0], frames, [1000 * k for k in range(len(clip))])


def test_cpp_example_program_writes_the_reference_container(tmp_path, real_clip):
    """examples/over_segment_b200 (C++: B200DenseSegmentation -> wire encoder -> container writer) on a B200: the file
    it writes holds, frame for frame, the messages of the compiled reference."""
    import os
    import struct
    import subprocess
    from video_segment_b200.segio import SegmentationReader, encode_frame_proto
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "over_segment_b200")
    if not os.path.exists(exe) or not rb.available(build=False):
        pytest.skip("oracle/_ref/over_segment_b200 / libref_results.so not shipped")
    clip = np.ascontiguousarray(real_clip[:12])
    t, h, w, _ = clip.shape
    src, dst = tmp_path / "in.bgr", tmp_path / "out.pb"
    src.write_bytes(struct.pack("<iii", w, h, t) + clip.tobytes())
    p = subprocess.run([exe, str(src), str(dst)], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-800:]
    r = SegmentationReader(str(dst))
    assert r.open_file_and_read_headers() and r.num_frames() == t and r.get_header_flags() == [1, 0]
    frames = [r.read_next_frame_binary() for _ in range(t)]
    r.close_file()
    ref = rc.run_stream(rb.ReferenceDense, clip, None, {})
    assert frames == [encode_frame_proto(rb.result_struct(d)) for d in ref]


def test_cpp_host_class_matches_compiled_reference():
    """segmentation::B200DenseSegmentation (video_segment_b200/host: the C++ class with DenseSegmentation's ProcessFrame
    signature over the C ABI, compiled against the reference's headers) on a flushed single chunk: the
    SegmentationDesc objects it returns equal the reference's in every field.  Runs in a child process
    (tests/gpu_host_class_probe.py): the class aborts through CHECK on errors, like the reference."""
    import os
    import subprocess
    import sys
    if not rb.host_available(build=False):
        pytest.skip("oracle/_ref/libb200_host_check.so not shipped")
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "gpu_host_class_probe.py")
    p = subprocess.run([sys.executable, probe, "real_single_chunk"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and p.stdout.startswith("OK"), (p.returncode, p.stdout[-400:], p.stderr[-800:])


def test_cpp_video_units_in_the_reference_framework(tmp_path):
    """B200DenseSegmentationUnit -> B200RegionSegmentationUnit as VideoUnits inside the reference's own video_framework
    (oracle/_ref/b200_units_check: video_unit.cpp compiled unmodified, the tree seg_tree_sample builds), with the
    reference's --chunk_size flag override, against the Python mirrors of the same two units: identical hierarchical
    records, word for word."""
    import struct
    import subprocess
    from video_segment_b200.unit import (DenseSegmentationOptions, DenseSegmentationUnit, RegionSegmentationOptions,
                                         RegionSegmentationUnit)
    exe = os.path.join(ROOT, "oracle", "_ref", "b200_units_check")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/b200_units_check not built")
    clip = np.load(os.path.join(ROOT, "tests", "golden", "real_clip_136x240x24.npz"))["frames"][:20]
    t, h, w, _ = clip.shape
    flows = np.random.default_rng(3).normal(0, 1.5, (t, h, w, 2)).astype(np.float32)
    for use_flow in (False, True):
        src = tmp_path / f"in{int(use_flow)}.bgr"
        blob = struct.pack("<iiii", w, h, t, int(use_flow)) + clip.tobytes() + (flows.tobytes() if use_flow else b"")
        src.write_bytes(blob)
        out = tmp_path / f"out{int(use_flow)}.bin"
        p = subprocess.run([exe, str(src), str(out), "8"], capture_output=True, text=True, timeout=600)
        assert p.returncode == 0, p.stderr[-600:]
        raw = out.read_bytes()
        n = struct.unpack_from("<q", raw, 0)[0]
        pos, recs = 8, []
        for _ in range(n):
            ln = struct.unpack_from("<q", raw, pos)[0]
            recs.append(np.frombuffer(raw, np.int32, ln, pos + 8).copy())
            pos += 8 + 4 * ln
        # the same two units through the Python mirrors
        dense = DenseSegmentationUnit(dense_seg_options=DenseSegmentationOptions(chunk_size=8))
        assert dense.open_streams(w, h, flow_stream_present=use_flow)
        region = RegionSegmentationUnit(RegionSegmentationOptions(), raw_records=True)
        assert region.open_streams(w, h, flow_stream_present=use_flow)
        want, fed = [], [0]

        def feed(results):
            for r in results:
                k = fed[0]
                want.extend(region.process_frame(r, clip[k], flows[k] if use_flow and k > 0 else None))
                fed[0] += 1
        for k, f in enumerate(clip):
            feed(dense.process_frame(f, flows[k] if use_flow else None))
        feed(dense.post_process())
        want.extend(region.post_process())
        dense.close(); region.close()
        assert n == t == len(want)
        for a, b in zip(recs, want):
            assert len(a) == len(b) and np.array_equal(a, b)
