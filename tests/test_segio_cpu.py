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
