"""CPU tests (-m "not gpu"): the oracle against the committed golden vectors, its own
invariants (the reference ships no tests: SURVEY.md section 4 lists the properties the source
promises), and the chunk arithmetic of dense_segmentation.cpp:281-432."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle_binding as ob
from video_segment_b200.synth import synth_clip

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_thirdparty_arithmetic_matches_cv2():
    g = np.load(os.path.join(GOLD, "cv2_thirdparty.npz"))
    conv = ob.convert_u8(g["src"])
    assert np.array_equal(conv, g["convert"])            # cv::Mat::convertTo(CV_32FC3, 1/255)
    assert float(conv.min()) == g["minmax"][0] and float(conv.max()) == g["minmax"][1]   # cv::minMaxLoc
    # BORDER_REPLICATE: the oracle's padded tile must equal cv::copyMakeBorder
    pad = np.pad(conv, ((4, 4), (4, 4), (0, 0)), mode="edge")
    assert np.array_equal(pad, g["border"])


def test_bilateral_matches_direct_numpy_restatement():
    rng = np.random.default_rng(0)
    img = rng.random((20, 24, 3), dtype=np.float32)
    out, lut, scale = ob.bilateral(img, want_lut=True)
    # independent float32 evaluation of image_filter.cpp:130-167 for a few pixels
    pad = np.pad(img, ((4, 4), (4, 4), (0, 0)), mode="edge")
    taps = [(i, j) for i in range(-4, 5) for j in range(-4, 5) if i * i + j * j <= 16]
    assert len(taps) == 49
    sc = np.float32(-0.5) / (np.float32(3.0) * np.float32(3.0))
    for (y, x) in [(0, 0), (5, 7), (19, 23), (10, 0)]:
        c = pad[y + 4, x + 4]
        ws = np.float32(0)
        acc = np.zeros(3, np.float32)
        for (i, j) in taps:
            l = pad[y + 4 + i, x + 4 + j]
            d = c - l
            idx = int((d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * np.float32(scale))
            wgt = np.float32(np.exp(np.float64(sc * np.float32(i * i + j * j)))) * lut[idx]
            ws = np.float32(ws + wgt)
            acc = (acc + l * wgt).astype(np.float32)
        inv = np.float32(1.0 / np.float64(ws))
        assert np.array_equal((acc * inv).astype(np.float32), out[y, x])


def test_bucket_index_edges():
    assert ob.bucket_index(0.0) == 0
    assert ob.bucket_index(1.0) == 2047          # w <= 1 never reaches the virtual bucket
    assert ob.bucket_index(1e10) == 2048         # ConstantPixelDistance(1e10) -> virtual bucket
    assert ob.bucket_index(0.5 / 2048) == 0 and ob.bucket_index(1.5 / 2048) == 1


def test_edge_weight_layout_and_counts():
    rng = np.random.default_rng(1)
    a = rng.random((7, 9, 3), dtype=np.float32)
    b = rng.random((7, 9, 3), dtype=np.float32)
    sp = ob.spatial_weights(a)
    tp = ob.temporal_weights(a, b)
    h, w = 7, 9
    assert (sp >= 0).sum() == (w - 1) * h + w * (h - 1) + 2 * (w - 1) * (h - 1)      # E_s, SURVEY 8
    assert (tp >= 0).sum() == (3 * w - 2) * (3 * h - 2)                               # E_t
    d = a[2, 3] - a[3, 4]
    assert sp[3, 2, 3] == np.sqrt(np.float32((d[0] * d[0] + d[1] * d[1] + d[2] * d[2]) * np.float32(1 / 3)))
    # flow displacement truncates toward zero and clamps (dense_segmentation_graph.h:1126-1130)
    flow = np.zeros((h, w, 2), np.float32)
    flow[..., 0] = 0.9
    assert np.array_equal(ob.temporal_weights(a, b, flow), tp)
    flow[..., 0] = -100.0
    t2 = ob.temporal_weights(a, b, flow)
    assert (t2[0] < 0).all()   # centre clamped to x = 0 -> no "left" neighbours


def test_oracle_pins_real_clip(real_clip):
    pins = json.load(open(os.path.join(GOLD, "oracle_pins.json")))
    o = ob.OracleDense(136, 240)
    res = []
    for f in real_clip:
        res += o.push(f)
    res += o.flush()
    assert len(res) == len(real_clip)
    assert [int(r["region_id"].size) for r in res] == pins["real_clip_regions_per_frame"]
    h = hashlib.sha256()
    for r in res:
        h.update(ob.id_map_from_result(r).tobytes())
    assert h.hexdigest() == pins["real_clip_id_maps_sha256"]
    assert hashlib.sha256(ob.preprocess(real_clip[0]).tobytes()).hexdigest() == pins["real_clip_frame0_smoothed_sha256"]


@pytest.mark.parametrize("threads", [1, 4])
def test_stream_contract_and_invariants(threads):
    """Chunk arithmetic (dense_segmentation.cpp:330-432) and the properties segmentation.proto
    promises: every pixel in exactly one region, sorted scan intervals, sorted ids when
    constrained, N4-connected slices, hierarchy only on a chunk's first frame."""
    W, H, T = 64, 48, 45
    clip = synth_clip(5, W, H, T)
    o = ob.OracleDense(W, H, num_threads=threads)
    got, batches = [], []
    for f in clip:
        r = o.push(f)
        if r:
            batches.append(len(r))
        got += r
    r = o.flush()
    batches.append(len(r))
    got += r
    assert batches == [19, 19, 7]                      # 19 new frames per chunk, rest on flush
    assert [g["pts"] for g in got] == list(range(T))
    min_region = int(np.float32(0.01) * W * np.float32(0.01) * H * 20)
    for t, g in enumerate(got):
        img = ob.id_map_from_result(g)
        assert (img >= 0).all()
        assert int((g["intervals"][:, 2] - g["intervals"][:, 1] + 1).sum()) == W * H   # disjoint cover
        off = g["interval_offset"]
        for k in range(g["region_id"].size):
            iv = g["intervals"][off[k]:off[k + 1]]
            key = iv[:, 0].astype(np.int64) * 100000 + iv[:, 1]
            assert (np.diff(key) > 0).all()
            assert g["shape_moments"][k, 0] == (iv[:, 2] - iv[:, 1] + 1).sum()
        if g["chunk_id"] > 0:
            assert (np.diff(g["region_id"]) > 0).all()
        first_of_chunk = t in (0, 19, 38)
        assert (g["compound"].shape[0] > 0) == first_of_chunk
        assert g["hierarchy_frame_idx"] == (0, 19, 38)[g["chunk_id"]]
        assert g["chunk_size"] == (19, 19, 7)[g["chunk_id"]]
    # region ids persist across the chunk boundary (constraints): frame 18 -> 19
    a, b = ob.id_map_from_result(got[18]), ob.id_map_from_result(got[19])
    common = set(np.unique(a)) & set(np.unique(b))
    assert len(common) > 0.25 * len(np.unique(a)), (len(common), len(np.unique(a)))
    assert min_region > 0


def test_threaded_oracle_equals_serial(real_clip):
    outs = []
    for th in (1, 8):
        o = ob.OracleDense(136, 240, num_threads=th)
        res = []
        for f in real_clip[:21]:
            res += o.push(f)
        outs.append(np.stack([ob.id_map_from_result(r) for r in res]))
    assert np.array_equal(outs[0], outs[1])
