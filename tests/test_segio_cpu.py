"""CPU tests of the result container (csrc/pb_io.cu, video_segment_b200/segio.py): host-only code,
so it runs without a GPU.  Checked byte for byte against the reference's own SegmentationWriter / SegmentationReader /
StripToEssentials (segment_util/segmentation_io.cpp compiled unmodified into oracle/_ref, where that is present) and
against the published layout otherwise."""
import os
import struct

import numpy as np
import pytest

import oracle_binding as ob
import reference_binding as rb
import reference_cases as rc
from video_segment_b200.segio import SegmentationReader, SegmentationWriter, strip_to_essentials


def _payloads(seed, n):
    rng = np.random.default_rng(seed)
    return [rng.integers(0, 256, int(rng.integers(0, 3000)), dtype=np.uint8).tobytes() for _ in range(n)]


def _write(path, entries, payloads, pts, chunk_every):
    w = SegmentationWriter(path)
    assert w.open_file(entries)
    for k, (p, t) in enumerate(zip(payloads, pts)):
        w.add_segmentation_data_to_chunk(p, t)
        if chunk_every and (k + 1) % chunk_every == 0:
            w.write_chunk()
    w.write_term_header_and_close()


def _read(path):
    r = SegmentationReader(path)
    assert r.open_file_and_read_headers()
    out = []
    while r.remaining_frames():
        out.append(r.read_next_frame_binary())
    flags, ts = r.get_header_flags(), r.time_stamps()
    r.close_file()
    return flags, out, ts


@pytest.mark.parametrize("n,chunk_every,entries", [(0, 0, [1, 0]), (1, 0, []), (7, 0, [1, 0]), (20, 5, [1, 0]), (23, 10, [3]), (6, 1, [1, 0, 7])])
def test_container_layout_and_roundtrip(tmp_path, n, chunk_every, entries):
    """segmentation_io.cpp:46-155 restated with struct: HEAD | CHNK ... SEGD ... | TERM, absolute offsets."""
    payloads, pts = _payloads(n, n), [1000 * k - 5 for k in range(n)]
    path = str(tmp_path / "a.pb")
    _write(path, entries, payloads, pts, chunk_every)
    raw = open(path, "rb").read()
    pos = 0
    assert raw[:4] == b"HEAD" and struct.unpack_from("<i", raw, 4)[0] == len(entries)
    assert list(struct.unpack_from(f"<{len(entries)}i", raw, 8)) == entries
    pos = 8 + 4 * len(entries)
    groups = [list(range(i, min(i + chunk_every, n))) for i in range(0, n, chunk_every)] if chunk_every else [list(range(n))]
    if chunk_every and n % chunk_every == 0 and n:
        pass                                  # the last WriteChunk left nothing pending: close writes no extra chunk
    elif not chunk_every and n == 0:
        groups = []                           # WriteTermHeaderAndClose skips the chunk when nothing is buffered
    chunk_id = 0
    for g in groups:
        assert raw[pos:pos + 4] == b"CHNK"
        cid, nf = struct.unpack_from("<ii", raw, pos + 4)
        assert (cid, nf) == (chunk_id, len(g))
        offs = struct.unpack_from(f"<{nf}q", raw, pos + 12)
        ts = struct.unpack_from(f"<{nf}q", raw, pos + 12 + 8 * nf)
        nxt = struct.unpack_from("<q", raw, pos + 12 + 16 * nf)[0]
        assert list(ts) == [pts[k] for k in g]
        p = pos + 12 + 16 * nf + 8
        for k, o in zip(g, offs):
            assert o == p and raw[p:p + 4] == b"SEGD" and struct.unpack_from("<i", raw, p + 4)[0] == len(payloads[k])
            assert raw[p + 8:p + 8 + len(payloads[k])] == payloads[k]
            p += 8 + len(payloads[k])
        assert nxt == p
        pos, chunk_id = p, chunk_id + 1
    assert raw[pos:pos + 4] == b"TERM" and struct.unpack_from("<i", raw, pos + 4)[0] == chunk_id and pos + 8 == len(raw)
    flags, got, ts = _read(path)
    assert flags == entries and got == payloads and ts == pts


@pytest.mark.parametrize("n,chunk_every", [(0, 0), (9, 0), (20, 5), (23, 10), (4, 1)])
def test_container_is_byte_identical_to_the_reference_writer(tmp_path, n, chunk_every):
    if not rb.host_available():
        pytest.skip("oracle/_ref/libb200_host_check.so not built (needs /root/reference)")
    payloads, pts = _payloads(100 + n, n), [40 * k for k in range(n)]
    mine, ref = str(tmp_path / "mine.pb"), str(tmp_path / "ref.pb")
    _write(mine, [1, 0], payloads, pts, chunk_every)
    rb.ref_io_write(ref, [1, 0], payloads, pts, chunk_every)
    assert open(mine, "rb").read() == open(ref, "rb").read()
    # each reader reads the other's file
    assert rb.ref_io_read(mine) == ([1, 0], payloads, pts)
    assert _read(ref) == ([1, 0], payloads, pts)


def test_reader_rejects_garbage(tmp_path):
    p = tmp_path / "bad.pb"
    p.write_bytes(b"HEAD\x00\x00\x00\x00JUNKJUNKJUNK")
    assert not SegmentationReader(str(p)).open_file_and_read_headers()
    p.write_bytes(b"HEAD\x00\x00\x00\x00CHNK")        # truncated
    assert not SegmentationReader(str(p)).open_file_and_read_headers()
    assert not SegmentationReader(str(tmp_path / "missing.pb")).open_file_and_read_headers()
    assert not SegmentationWriter(str(tmp_path / "no_such_dir" / "x.pb")).open_file([1, 0])


@pytest.mark.parametrize("moments", [False, True])
def test_strip_to_essentials_matches_reference(moments):
    """StripToEssentials(desc, false, moments) of the reference (compiled unmodified) on the messages of an oracle
    stream vs vsb200_strip_to_essentials on the same frame results."""
    if not rb.host_available():
        pytest.skip("oracle/_ref/libb200_host_check.so not built (needs /root/reference)")
    clip, flows, opts = rc.load_case("real_chunk8")
    res = rc.run_stream(ob.OracleDense, clip[:10], None, opts)
    assert any(len(d["compound"]) for d in res) and any(not len(d["compound"]) for d in res)
    for d in res:
        mine = strip_to_essentials(rb.result_struct(d), moments)
        assert mine == rb.ref_io_strip(d, moments)
        w, h, nreg = struct.unpack_from("<iii", mine, 0)
        assert (w, h, nreg) == (d["width"], d["height"], len(d["region_id"]))


def _message_from_result(Desc, d):
    """The SegmentationDesc protobuf (real protobuf runtime, schema restated in tests/proto_schema.py) holding the fields
    RetrieveSegmentation3D / SegmentAndOutputChunk set for a frame (segmentation.cpp:458-533, dense_segmentation.cpp:385-387)."""
    m = Desc()
    m.frame_width, m.frame_height, m.chunk_id = d["width"], d["height"], d["chunk_id"]
    m.connectedness = d["connectedness"]
    m.chunk_size, m.overlap_start, m.hierarchy_frame_idx = d["chunk_size"], d["overlap_start"], d["hierarchy_frame_idx"]
    off = d["interval_offset"]
    for k, rid in enumerate(d["region_id"]):
        r = m.region.add()
        r.id = int(rid)
        r.raster.SetInParent()
        for y, lx, rx in d["intervals"][off[k]:off[k + 1]]:
            s = r.raster.scan_inter.add()
            s.y, s.left_x, s.right_x = int(y), int(lx), int(rx)
        sm = d["shape_moments"][k]
        r.shape_moments.size, r.shape_moments.mean_x, r.shape_moments.mean_y = float(sm[0]), float(sm[1]), float(sm[2])
        r.shape_moments.moment_xx, r.shape_moments.moment_xy, r.shape_moments.moment_yy = float(sm[3]), float(sm[4]), float(sm[5])
    if len(d["compound"]):
        h = m.hierarchy.add()
        no = d["neighbor_offset"]
        for k, (cid, size, start, end) in enumerate(d["compound"]):
            c = h.region.add()
            c.id, c.size, c.start_frame, c.end_frame = int(cid), int(size), int(start), int(end)
            c.neighbor_id.extend(int(x) for x in d["neighbor_id"][no[k]:no[k + 1]])
    return m


@pytest.mark.parametrize("case", ["real_chunk8", "synth_flow", "synth_one_frame"])
def test_wire_encoder_equals_protobuf_serialisation(case, tmp_path):
    """vsb200_encode_frame_proto (the engine's wire encoder, host code) writes exactly the bytes the protobuf runtime
    serialises for the same SegmentationDesc, on whole oracle streams (= the reference's messages, test_oracle_cpu);
    a file of such frames parses back to the same messages."""
    from proto_schema import segmentation_desc_class
    from video_segment_b200.segio import encode_frame_proto
    Desc = segmentation_desc_class()
    clip, flows, opts = rc.load_case(case)
    res = rc.run_stream(ob.OracleDense, clip, flows, opts)
    path = str(tmp_path / "stream.pb")
    w = SegmentationWriter(path)
    assert w.open_file([1, 0])
    want = []
    for k, d in enumerate(res):
        mine = encode_frame_proto(rb.result_struct(d))
        m = _message_from_result(Desc, d)
        assert mine == m.SerializeToString(deterministic=True), (case, k)
        back = Desc()
        back.ParseFromString(mine)
        assert back == m
        w.add_segmentation_data_to_chunk(mine, 40 * k)
        want.append(mine)
    w.write_term_header_and_close()
    assert _read(path) == ([1, 0], want, [40 * k for k in range(len(res))])


def test_strip_and_wire_encoder_on_random_streams():
    """The two serialisers of csrc (vsb200_strip_to_essentials, vsb200_encode_frame_proto) on the result streams of 24
    random cases (tests/reference_cases.random_case): byte identical to the reference's StripToEssentials (where the
    compiled reference is present) and to the protobuf runtime."""
    from proto_schema import segmentation_desc_class
    from video_segment_b200.segio import encode_frame_proto
    Desc = segmentation_desc_class()
    have_ref = rb.host_available()
    frames = 0
    for seed in range(24):
        clip, flows, opts = rc.random_case(seed)
        for d in rc.run_stream(ob.OracleDense, clip, flows, opts):
            s = rb.result_struct(d)
            assert encode_frame_proto(s) == _message_from_result(Desc, d).SerializeToString(deterministic=True), seed
            if have_ref:
                for moments in (False, True):
                    assert strip_to_essentials(s, moments) == rb.ref_io_strip(d, moments), (seed, moments)
            frames += 1
    assert frames > 100
