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


# ---- region stage: appearance descriptor (oracle/vso_region.cpp) ----

def _lab_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lab_bgr2lab_cv2.npz"))


def test_bgr2lab_matches_cv2_golden_vectors_and_full_cube_checksum():
    """Third-party arithmetic (cv::cvtColor(CV_BGR2Lab), region_descriptor.cpp:73): pinned by cv2 4.13 answers."""
    import hashlib
    g = _lab_golden()
    got = ob.bgr2lab(g["sample_bgr"].reshape(-1, 1, 3)).reshape(-1, 3)
    assert np.array_equal(got, g["sample_lab"])
    r, gg, b = np.meshgrid(np.arange(256), np.arange(256), np.arange(256), indexing="ij")
    cube = np.stack([b, gg, r], -1).astype(np.uint8).reshape(65536, 256, 3)
    digest = hashlib.sha256(ob.bgr2lab(cube).tobytes()).digest()
    assert digest == g["cube_sha256"].tobytes()
    # known answers: black, white, primaries (OpenCV 8-bit Lab: L * 255 / 100, a + 128, b + 128)
    px = np.uint8([[[0, 0, 0]], [[255, 255, 255]], [[0, 0, 255]], [[0, 255, 0]], [[255, 0, 0]]])
    assert ob.bgr2lab(px).reshape(-1, 3).tolist() == [[0, 128, 128], [255, 128, 128], [136, 208, 195], [224, 42, 211], [82, 207, 20]]


def _numpy_region_hist(lab, ids, n_regions, lum_bins, color_bins):
    """Independent float64 restatement of AddValueInterpolated (histograms.cpp:140-204) with numpy scatter-adds."""
    total = lum_bins * color_bins * color_bins
    hist = np.zeros((n_regions, total))
    ids = ids.reshape(-1)
    px = lab.reshape(-1, 3)
    ok = (ids >= 0) & (ids < n_regions)
    f = np.float32
    xb = px[:, 0].astype(f) * f(1.0 / 255.0) * f(lum_bins - 1)
    yb = px[:, 1].astype(f) * f(1.0 / 255.0) * f(color_bins - 1)
    zb = px[:, 2].astype(f) * f(1.0 / 255.0) * f(color_bins - 1)
    ix, iy, iz = xb.astype(np.int32), yb.astype(np.int32), zb.astype(np.int32)
    dx, dy, dz = xb - ix.astype(f), yb - iy.astype(f), zb - iz.astype(f)
    for a in range(2):
        for b in range(2):
            for c in range(2):
                bx = ix + (a & (dx >= f(1e-6))); by = iy + (b & (dy >= f(1e-6))); bz = iz + (c & (dz >= f(1e-6)))
                wv = ((dx if a else f(1) - dx) * (dy if b else f(1) - dy)) * (dz if c else f(1) - dz)
                np.add.at(hist, (ids[ok], (bx * color_bins * color_bins + by * color_bins + bz)[ok]), wv[ok].astype(np.float64))
    cnt = np.bincount(ids[ok], minlength=n_regions).astype(np.float64)
    return hist, cnt


def test_region_hist_against_independent_restatement():
    rng = np.random.default_rng(5)
    h, w, nr = 40, 56, 7
    lab = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    lab[:4] = 255                                   # top of the range: last bin, zero fraction
    lab[4:8, :, 1] = 0
    ids = rng.integers(-1, nr + 1, size=(h, w)).astype(np.int32)   # -1 and nr: pixels outside every region
    for lum, col in ((10, 20), (4, 3)):
        ref, cnt = _numpy_region_hist(lab, ids, nr, lum, col)
        exact, wsum = ob.region_hist([lab], [ids], nr, lum, col, exact=True)
        assert np.array_equal(wsum, cnt)
        assert np.abs(exact - ref / np.maximum(cnt, 1)[:, None]).max() <= 1e-7
        like_ref, _ = ob.region_hist([lab], [ids], nr, lum, col, exact=False)
        assert np.abs(like_ref - exact).max() <= 1e-5      # float accumulation noise of the reference itself
        assert np.abs(like_ref.sum(1) - (cnt > 0)).max() <= 1e-4      # L1 norm 1 (NormalizeToOne)
    # two frames accumulate into the same descriptor (3-D regions)
    two, w2 = ob.region_hist([lab, lab[::-1]], [ids, ids[::-1]], nr, exact=True)
    one, w1 = ob.region_hist([lab], [ids], nr, exact=True)
    assert np.array_equal(w2, 2 * w1) and np.abs(two - one).max() <= 1e-7


def test_hist_chisquare_properties():
    rng = np.random.default_rng(6)
    hist = rng.random((5, 4000)).astype(np.float32)
    hist[:, rng.random(4000) < 0.7] = 0            # sparse like real descriptors
    hist /= hist.sum(1, keepdims=True)
    hist[4] = 0                                     # an empty region
    pairs = np.int32([[0, 1], [1, 0], [2, 2], [3, 4], [0, 3]])
    d = ob.hist_chisquare(hist, pairs)
    a, b = hist[pairs[:, 0]].astype(np.float64), hist[pairs[:, 1]].astype(np.float64)
    with np.errstate(invalid="ignore", divide="ignore"):
        ref = 0.5 * np.where(a + b > 1e-12, (a - b) ** 2 / (a + b), 0).sum(1)
    assert np.abs(d - ref).max() <= 1e-6
    assert d[0] == d[1] and d[2] == 0 and abs(d[3] - 0.5) <= 1e-6 and np.all(d <= 1.0 + 1e-6)


# ---- the reference's own ColorHistogram (oracle/_ref, compiled unmodified from /root/reference) ----

def _ref_hist_lib():
    import ctypes as C
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "libref_hist.so")
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(root, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_hist.so not built (needs /root/reference)")
    L = C.CDLL(path)
    L.ref_color_hist.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.ref_color_hist.restype = C.c_double
    L.ref_color_hist_chisquare.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
    L.ref_color_hist_chisquare.restype = C.c_float
    return L


@pytest.mark.parametrize("bins", [(10, 20), (4, 3), (16, 16)])
def test_region_hist_oracle_equals_reference_color_histogram(bins):
    """Pins oracle/vso_region.cpp to the reference itself: histograms.cpp compiled as it lies (glog stand-in only)."""
    L = _ref_hist_lib()
    lum, col = bins
    total = lum * col * col
    rng = np.random.default_rng(17)
    frame = synth_clip(15, 96, 64, 1)[0]
    lab = ob.bgr2lab(frame)
    lab[:2] = 255
    lab[2:4] = 0
    ids = np.kron(rng.integers(0, 5, size=(8, 8)).astype(np.int32), np.ones((8, 12), np.int32))
    mine, wsum = ob.region_hist([lab], [ids], 5, lum, col, exact=False)
    sets = []
    for r in range(5):
        px = np.ascontiguousarray(lab[ids == r])                    # raster order == scan-interval order
        sets.append(px)
        out = np.zeros(total, np.float32)
        w = L.ref_color_hist(px.ctypes.data, len(px), lum, col, out.ctypes.data)
        assert w == wsum[r] == len(px)
        assert np.array_equal(out, mine[r]), (r, np.abs(out - mine[r]).max())
    pairs = np.int32([[0, 1], [1, 2], [3, 3], [4, 0]])
    d = ob.hist_chisquare(mine, pairs)
    for (a, b), dv in zip(pairs, d):
        ref_dense = L.ref_color_hist_chisquare(sets[a].ctypes.data, len(sets[a]), sets[b].ctypes.data, len(sets[b]), lum, col, 0)
        ref_sparse = L.ref_color_hist_chisquare(sets[a].ctypes.data, len(sets[a]), sets[b].ctypes.data, len(sets[b]), lum, col, 1)
        assert dv == ref_dense
        assert abs(dv - ref_sparse) <= 1e-7          # the hash map walks the bins in another order (double sum)


# ---- the reference's own FastSegmentationGraph + pixel distances (oracle/_ref/libref_graph.so) ----

def _ref_graph_lib():
    import ctypes as C
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "libref_graph.so")
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(root, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_graph.so not built (needs /root/reference)")
    L = C.CDLL(path)
    L.ref_segment_chunk_labels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ref_bucket_index.argtypes = [C.c_float]
    return L


def _ref_segment(L, frames, min_region_size, l1=False, force_constraints=True):
    t, h, w, _ = frames.shape
    frames = np.ascontiguousarray(frames, np.float32)
    labels = np.empty((t, h, w), np.int32)
    sp = np.empty((t, h, w, 4), np.float32)
    tp = np.empty((t, h, w, 9), np.float32)
    L.ref_segment_chunk_labels(frames.ctypes.data, w, h, t, int(l1), int(min_region_size), int(force_constraints),
                               labels.ctypes.data, sp.ctypes.data, tp.ctypes.data)
    return labels, sp, tp


def _same_partition(a, b):
    a, b = a.reshape(-1), b.reshape(-1)
    _, ia = np.unique(a, return_inverse=True)
    _, ib = np.unique(b, return_inverse=True)
    pairs = np.unique(np.stack([ia, ib], 1), axis=0)
    return len(pairs) == ia.max() + 1 == ib.max() + 1


@pytest.mark.parametrize("case", ["real", "synth", "synth_l1", "tiny_regions"])
def test_merge_oracle_equals_reference_segmentation_graph(case, real_clip):
    """Pins the oracle's merge (vso_graph.cpp) and edge weights to the reference's own code: FastSegmentationGraph::
    SegmentGraph / MergeRegions / ColorMeanDescriptorTraits and the CvMat distance walkers, compiled as they lie."""
    L = _ref_graph_lib()
    l1 = case == "synth_l1"
    if case == "real":
        clip = real_clip[:10]
    elif case == "tiny_regions":
        clip = synth_clip(23, 40, 32, 5)
    else:
        clip = synth_clip(22, 96, 72, 6)
    frames = np.stack([ob.preprocess(f) for f in clip])
    t, h, w, _ = frames.shape
    mins = 4 if case == "tiny_regions" else int(0.01 * w * 0.01 * h * t) or 1    # dense_segmentation.cpp: frac^2 * w * h * frames
    ref_labels, sp, tp = _ref_segment(L, frames, mins, l1)
    # edge weights: bit identical to the reference's walkers
    for k in range(t):
        assert np.array_equal(ob.spatial_weights(frames[k], l1), sp[k].transpose(2, 0, 1))
        if k:
            assert np.array_equal(ob.temporal_weights(frames[k], frames[k - 1], None, l1), tp[k].transpose(2, 0, 1))
    for wv in np.concatenate([sp.reshape(-1)[::97], np.float32([0, 1, 0.5, 1e10])]):
        if wv >= 0:
            assert ob.bucket_index(float(wv)) == L.ref_bucket_index(float(wv))
    # the merge: identical partition of the voxels
    mine = ob.segment_chunk_labels(frames, mins, l1)
    assert _same_partition(mine, ref_labels)
    assert len(np.unique(ref_labels)) > 1


# ---- the reference's own BilateralFilter (oracle/_ref/libref_filter.so) ----

def _ref_filter_lib():
    import ctypes as C
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    path = os.path.join(root, "oracle", "_ref", "libref_filter.so")
    if os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(root, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libref_filter.so not built (needs /root/reference)")
    L = C.CDLL(path)
    L.ref_preprocess_bilateral.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    L.ref_bilateral_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p]
    return L


@pytest.mark.parametrize("case", ["real", "synth", "flat", "two_level", "tiny", "odd"])
def test_preprocess_oracle_equals_reference_bilateral_filter(case, real_clip):
    """Pins the oracle's PreprocessFeatures (vso_preprocess.cpp: convert, replicate border, exp LUT, 49-tap bilateral
    filter) bit for bit to the reference's own imagefilter::BilateralFilter compiled as it lies (image_filter.cpp:184-277)."""
    L = _ref_filter_lib()
    rng = np.random.default_rng(5)
    if case == "real":
        frames = [real_clip[0], real_clip[13]]
    elif case == "synth":
        frames = list(synth_clip(31, 96, 72, 2))
    elif case == "flat":              # max == min: the 1e-3 floor of diff_range, every LUT index 0
        frames = [np.full((20, 24, 3), 77, np.uint8)]
    elif case == "two_level":         # narrow value range: scale is large, LUT reaches its zero tail
        frames = [(rng.integers(0, 2, (33, 47, 3)) * 3 + 100).astype(np.uint8)]
    elif case == "tiny":              # about the filter radius: the window is mostly replicated border.  (The reference
        # itself never returns for height < 8: its ParallelFor grain is height / 8 = 0, image_filter.cpp:255.)
        frames = [rng.integers(0, 256, (8, 5, 3), dtype=np.uint8), rng.integers(0, 256, (9, 1, 3), dtype=np.uint8)]
    else:
        frames = [rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)]
    for f in frames:
        f = np.ascontiguousarray(f)
        h, w, _ = f.shape
        ref = np.empty((h, w, 3), np.float32)
        L.ref_preprocess_bilateral(f.ctypes.data, w, h, 3 * w, 3.0, 0.25, ref.ctypes.data)
        mine = ob.preprocess(f)
        assert np.array_equal(mine.view(np.uint32), ref.view(np.uint32))
        # the filter alone, other sigmas (radius 2 and 6)
        img = ob.convert_u8(f)
        for ss, sc in ((1.5, 0.1), (4.0, 0.5)):
            ref2 = np.empty_like(img)
            L.ref_bilateral_f32(img.ctypes.data, w, h, 3, ss, sc, ref2.ctypes.data)
            assert np.array_equal(ob.bilateral(img, ss, sc).view(np.uint32), ref2.view(np.uint32))


# ---- the reference's own streaming over-segmentation (oracle/_ref/libref_results.so) ----

import reference_cases as rc      # noqa: E402


@pytest.mark.parametrize("case", sorted(rc.CASES))
def test_oracle_matches_reference_golden(case):
    """The oracle engine against digests of the REFERENCE's own per-frame results (tests/golden/reference_results.json,
    written by tests/golden/make_reference_golden.py from the unmodified DenseSegmentation pipeline): every header
    field, region id, scan interval, shape moment (float bits) and hierarchy-level-0 entry with neighbours."""
    gold = json.load(open(os.path.join(GOLD, "reference_results.json")))[case]
    clip, flows, opts = rc.load_case(case)
    res = rc.run_stream(ob.OracleDense, clip, flows, opts)
    assert len(res) == gold["frames"]
    assert [int(r["region_id"].size) for r in res] == gold["regions_per_frame"]
    assert rc.digest(res) == gold["sha256"]


@pytest.mark.parametrize("case", sorted(rc.CASES))
def test_oracle_equals_compiled_reference(case):
    """Same cases, field by field against the compiled reference where oracle/_ref/libref_results.so exists, and the
    committed digests against what the library produces now (guards the fixture against a stale library)."""
    import reference_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libref_results.so not built (needs /root/reference)")
    clip, flows, opts = rc.load_case(case)
    ref = rc.run_stream(rb.ReferenceDense, clip, flows, opts)
    assert rc.digest(ref) == json.load(open(os.path.join(GOLD, "reference_results.json")))[case]["sha256"]
    assert rc.first_difference(ref, rc.run_stream(ob.OracleDense, clip, flows, opts)) is None
    # the oracle's reference-style threading (parallel graph construction) must not change a bit either
    if case in ("real_chunk8", "synth_flow"):
        assert rc.first_difference(ref, rc.run_stream(ob.OracleDense, clip, flows, dict(opts, num_threads=4))) is None


@pytest.mark.parametrize("case", ["real_chunk8", "real_l1_single_chunk_plain", "synth_flow", "synth_one_frame"])
def test_cpp_host_side_rebuilds_reference_messages(case):
    """video_segment_b200/host: FrameResultToSegmentationDesc (the C++ host side's array -> protobuf step, compiled
    against the reference's own headers) applied to the flattened results of the compiled reference reproduces the
    reference's SegmentationDesc objects exactly, values and presence bits."""
    import reference_binding as rb
    if not rb.host_available():
        pytest.skip("oracle/_ref/libb200_host_check.so not built (needs /root/reference)")
    clip, flows, opts = rc.load_case(case)
    bad, msg = rb.host_check_desc_vs_reference(clip, flows, **opts)
    assert bad == 0, msg


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_product_host_shape_helpers_equal_reference_functions(seed):
    """video_segment_b200/csrc/shape_math.hpp (the moment accumulator shared by the device kernel of shape.cu and the host,
    shape descriptors, oriented boxes) and csrc/region_raster.hpp (raster union of the hierarchical stage) against the
    reference's own segment_util/segmentation_util.cpp functions, compiled unmodified, on random rasters; floats by bits."""
    import reference_binding as rb
    if not rb.host_available():
        pytest.skip("oracle/_ref/libb200_host_check.so not built (needs /root/reference)")
    bad, msg = rb.host_shape_check(seed, 400)
    assert bad == 0, msg


def test_oracle_equals_compiled_reference_at_1080p():
    """BASELINE config C geometry: a flushed 3-frame 1920x1080 chunk (20 M edges) through the compiled reference and
    the oracle (reference-style threading), identical in every field."""
    import reference_binding as rb
    from video_segment_b200.synth import synth
    if not rb.available():
        pytest.skip("oracle/_ref/libref_results.so not built (needs /root/reference)")
    clip = np.stack(list(synth(3, 1920, 1080, 3)))
    ref = rc.run_stream(rb.ReferenceDense, clip, None, {})
    mine = rc.run_stream(ob.OracleDense, clip, None, dict(num_threads=os.cpu_count() or 1))
    assert rc.first_difference(ref, mine) is None
    assert len(ref[0]["region_id"]) > 50


@pytest.mark.parametrize("block", range(4))
def test_oracle_equals_compiled_reference_on_random_cases(block):
    """Differential test: 4 x 12 random small clips (noise, gradients, moving blocks, speckles; random chunk size,
    overlap, constraint frames, colour distance, N4 / connectedness switches, min region size, random flow) through the
    compiled reference and the oracle, identical in every field.  (4 000 seeds were run once by hand: no difference.)"""
    import reference_binding as rb
    if not rb.available():
        pytest.skip("oracle/_ref/libref_results.so not built (needs /root/reference)")
    for seed in range(12 * block, 12 * block + 12):
        clip, flows, opts = rc.random_case(seed)
        ref = rc.run_stream(rb.ReferenceDense, clip, flows, opts)
        mine = rc.run_stream(ob.OracleDense, clip, flows, opts)
        assert rc.first_difference(ref, mine) is None, (seed, clip.shape, opts)


@pytest.mark.parametrize("seed,with_flow", [(1, False), (2, False), (3, True), (4, True)])
def test_product_tube_split_equals_oracle(seed, with_flow):
    """video_segment_b200/csrc/tubes.hpp (the host half of EnforceSpatialConnectedness: tube matching, folds, joins, on
    the moment accumulator of csrc/shape_math.hpp) against the oracle's restatement of dense_segmentation_graph.h:666-904
    on random label volumes (moving blobs that break apart and rejoin, specks), with and without flow: the same
    regions in the same order on every volume."""
    bad, msg = ob.host_tubes_check(seed, 200, with_flow)
    assert bad == 0, msg
    assert "regions split" in msg and int(msg.split(",")[1].split()[0]) > 100, msg      # the volumes do exercise the split
