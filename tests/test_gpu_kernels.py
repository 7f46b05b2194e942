"""GPU parity tests (-m gpu): every hand-written kernel, called through the C ABI, against
the CPU oracle on the same seeded inputs.  Integer / index results must be bit exact; float
results are compared bit exact as well where the arithmetic is replicated op by op
(no FMA), with the 1e-5 bar of BASELINE.json as the documented tolerance."""
import os

import numpy as np
import pytest

import oracle_binding as ob
from helpers import overseg_iou, partition_equal
from video_segment_b200.synth import synth_clip

pytestmark = pytest.mark.gpu

EDGE_TOL = 1e-5   # BASELINE.json: fp32 edge weights within 1e-5


@pytest.fixture(scope="module")
def K():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from video_segment_b200 import kernels
    from video_segment_b200._lib import lib
    assert lib().vsb200_device_count() >= 1, "no sm_100 device: the CUDA path cannot run"
    return kernels


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape", [(48, 64), (37, 53), (240, 136), (480, 640)])
def test_preprocess_bilateral_parity(K, shape, real_clip):
    h, w = shape
    if shape == (240, 136):
        frame = real_clip[3]
    else:
        frame = synth_clip(7, w, h, 1)[0]
    ref = ob.preprocess(frame)
    got = K.preprocess(_dev(frame)).cpu().numpy()
    err = np.abs(got - ref).max()
    assert err <= 1e-6, err
    # op-by-op replication: expected to be bit identical except where the device exp() of the
    # LUT rounds differently from glibc (never observed)
    assert (got != ref).mean() < 1e-4


def test_preprocess_none_and_constant_frame(K):
    frame = synth_clip(8, 40, 30, 1)[0]
    assert np.array_equal(K.preprocess(_dev(frame), presmoothing=0).cpu().numpy(), ob.preprocess(frame, presmoothing=0))
    flat = np.full((30, 40, 3), 77, np.uint8)          # max == min -> diff_range clamps to 1e-3
    assert np.array_equal(K.preprocess(_dev(flat)).cpu().numpy(), ob.preprocess(flat))


@pytest.mark.parametrize("shape,l1", [((48, 64), False), ((37, 53), False), ((37, 53), True), ((480, 640), False)])
def test_edge_weights_parity(K, shape, l1):
    h, w = shape
    clip = synth_clip(9, w, h, 2)
    a, b = ob.preprocess(clip[1]), ob.preprocess(clip[0])
    sp_ref, tp_ref = ob.spatial_weights(a, l1), ob.temporal_weights(a, b, None, l1)
    sp, tp = K.edge_build(_dev(a), _dev(b), None, l1)
    sp = sp.cpu().numpy().transpose(2, 0, 1)
    tp = tp.cpu().numpy().transpose(2, 0, 1)
    assert np.array_equal(sp < 0, sp_ref < 0) and np.array_equal(tp < 0, tp_ref < 0)
    assert np.abs(sp - sp_ref).max() <= EDGE_TOL and np.abs(tp - tp_ref).max() <= EDGE_TOL
    assert np.array_equal(sp, sp_ref) and np.array_equal(tp, tp_ref)     # bit exact in practice
    # spatial only (first frame of a chunk)
    sp0, none = K.edge_build(_dev(a), None, None, l1)
    assert none is None and np.array_equal(sp0.cpu().numpy().transpose(2, 0, 1), sp_ref)


@pytest.mark.parametrize("shape", [(24, 128), (37, 68), (17, 192)])
def test_edge_weights_flat_patches_parity(K, shape):
    """Identical neighbouring pixels (zero colour distance) take the exact-sqrt slow path of the packed /
    scalar TMA kernels; widths cover the packed kernel (w % 64 == 0) and the scalar TMA kernel (w % 4 == 0)."""
    h, w = shape
    rng = np.random.default_rng(21)
    a = rng.random((h, w, 3), dtype=np.float32)
    b = rng.random((h, w, 3), dtype=np.float32)
    a[3:9, 5:40] = a[3, 5]                      # flat patch inside frame t
    b[2:12, 20:70] = a[3, 5]                    # and the same colour in frame t-1
    a[h - 4:, w - 30:] = 0.0
    b[h - 4:, w - 30:] = 0.0
    a[10:14, :] *= np.float32(1e-20)            # tiny but non-zero differences (below the fast-path range)
    sp_ref, tp_ref = ob.spatial_weights(a), ob.temporal_weights(a, b, None)
    sp, tp = K.edge_build(_dev(a), _dev(b), None, False)
    assert np.array_equal(sp.cpu().numpy().transpose(2, 0, 1), sp_ref)
    assert np.array_equal(tp.cpu().numpy().transpose(2, 0, 1), tp_ref)


def test_edge_weights_flow_parity(K):
    h, w = 45, 70
    clip = synth_clip(10, w, h, 2)
    a, b = ob.preprocess(clip[1]), ob.preprocess(clip[0])
    rng = np.random.default_rng(3)
    flow = rng.uniform(-6, 6, size=(h, w, 2)).astype(np.float32)
    tp_ref = ob.temporal_weights(a, b, flow)
    sp, tp = K.edge_build(_dev(a), _dev(b), _dev(flow))
    assert np.array_equal(tp.cpu().numpy().transpose(2, 0, 1), tp_ref)
    assert np.array_equal(sp.cpu().numpy().transpose(2, 0, 1), ob.spatial_weights(a))


def test_bucket_index_parity(K):
    rng = np.random.default_rng(4)
    ws = np.concatenate([rng.random(2000, dtype=np.float32), np.float32([0, 1, 1e10]),
                         (np.arange(1, 2048, dtype=np.float32) / np.float32(2048 / (1.0 + 1e-6)))])
    for w in ws:
        assert K.bucket_index(float(w)) == ob.bucket_index(float(w))


def _ref_sorted_codes(lists_np, w, h):
    """numpy restatement of the reference traversal order: bucket asc, list asc, insertion order."""
    n = w * h
    codes, buckets = [], []
    for q, t in enumerate(lists_np):
        if t is None:
            continue
        flat = t.reshape(-1)                       # [pixel][dir]
        nd = t.shape[-1]
        idx = np.nonzero(flat >= 0)[0]
        pix, d = idx // nd, idx % nd
        nd = 9 if (q & 1) else 4
        codes.append(((q // 2) * 13 * n + (4 * n if (q & 1) else 0) + pix * nd + d).astype(np.uint32))
        scale = np.float32(2048) / (np.float32(1.0) + np.float32(1e-6))
        buckets.append(np.minimum(np.float32(2048), flat[idx] * scale).astype(np.int32))
    codes = np.concatenate(codes)
    buckets = np.concatenate(buckets)
    order = np.argsort(buckets, kind="stable")
    return codes[order], np.bincount(buckets, minlength=2048)


@pytest.mark.parametrize("shape,t", [((37, 53), 3), ((120, 160), 4)])
def test_sort_edges_stable_parity(K, shape, t):
    h, w = shape
    clip = synth_clip(11, w, h, t)
    sm = [ob.preprocess(f) for f in clip]
    lists_np, lists_dev = [None] * (2 * t - 1), [None] * (2 * t - 1)
    for s in range(t):
        sp, tp = K.edge_build(_dev(sm[s]), _dev(sm[s - 1]) if s else None)
        lists_dev[2 * s] = sp
        lists_np[2 * s] = sp.cpu().numpy()
        if s:
            lists_dev[2 * s - 1] = tp
            lists_np[2 * s - 1] = tp.cpu().numpy()
    codes, bstart = K.sort_edges(lists_dev, w, h)
    ref_codes, ref_counts = _ref_sorted_codes(lists_np, w, h)
    bstart = bstart.cpu().numpy()
    assert bstart[0] == 0 and bstart[-1] == ref_codes.size
    assert np.array_equal(np.diff(bstart), ref_counts)
    assert np.array_equal(codes.cpu().numpy().view(np.uint32)[:ref_codes.size], ref_codes)


def _merge_case(K, frames_u8, min_region):
    sm = np.stack([ob.preprocess(f) for f in frames_u8])
    ref = ob.segment_chunk_labels(sm, min_region)
    lab, stats = K.segment_chunk(_dev(sm), min_region)
    return ref, lab.cpu().numpy(), stats


def test_merge_parity_synthetic(K):
    clip = synth_clip(12, 96, 72, 6)
    ref, got, stats = _merge_case(K, clip, int(np.float32(0.01) * 96 * np.float32(0.01) * 72 * 20))
    iou = min(overseg_iou(ref[t], got[t]) for t in range(ref.shape[0]))
    assert iou >= 0.99, (iou, stats)
    assert partition_equal(ref, got), (iou, stats)


def test_merge_parity_real_clip(K, real_clip):
    ref, got, stats = _merge_case(K, real_clip[:8], int(np.float32(0.01) * 136 * np.float32(0.01) * 240 * 20))
    iou = min(overseg_iou(ref[t], got[t]) for t in range(ref.shape[0]))
    assert iou >= 0.99, (iou, stats)
    assert partition_equal(ref, got), (iou, stats)


def test_merge_min_region_size_property(K):
    """Size-independent property at a larger size: no final region is smaller than
    min_region_size unless it never met a later edge."""
    clip = synth_clip(13, 320, 240, 5)
    minr = 150
    sm = np.stack([ob.preprocess(f) for f in clip])
    lab, stats = K.segment_chunk(_dev(sm), minr)
    lab = lab.cpu().numpy()
    _, counts = np.unique(lab, return_counts=True)
    assert (counts < minr).mean() < 0.02, (counts < minr).mean()


# ---- region stage: appearance descriptor (csrc/region_hist.cu) ----

def test_bgr2lab_bit_exact_over_the_colour_cube(K):
    """Integer path: bit exact against the oracle (itself pinned to cv2 over the same cube, CPU suite)."""
    r, g, b = np.meshgrid(np.arange(0, 256, 3), np.arange(256), np.arange(256), indexing="ij")     # 86 x 256 x 256 colours
    cube = np.stack([b, g, r], -1).astype(np.uint8).reshape(-1, 256, 3)
    got = K.bgr2lab(_dev(cube)).cpu().numpy()
    assert np.array_equal(got, ob.bgr2lab(cube))
    frame = synth_clip(12, 70, 45, 1)[0]                       # odd width: unaligned rows
    assert np.array_equal(K.bgr2lab(_dev(frame)).cpu().numpy(), ob.bgr2lab(frame))


@pytest.mark.parametrize("bins", [(10, 20), (4, 3)])
def test_region_hist_parity(K, bins):
    """Floating point: the kernel sums 2^-26 fixed-point weights (order independent), the reference sums floats in
    raster order.  Tolerances: 1e-7 against the oracle's exact (double) accumulation, 2e-5 against its float
    accumulation (the reference's own rounding noise), on L1-normalised bins in [0, 1]."""
    import torch
    lum, col = bins
    clip = synth_clip(13, 160, 120, 3)
    rng = np.random.default_rng(8)
    nr = 23
    blocks = rng.integers(-1, nr + 1, size=(3, 8, 10)).astype(np.int32)       # coarse region maps; -1 / nr = no region
    ids = [np.kron(bm, np.ones((15, 16), np.int32)) for bm in blocks]
    labs = [ob.bgr2lab(f) for f in clip]
    exact, wsum = ob.region_hist(labs, ids, nr, lum, col, exact=True)
    like_ref, _ = ob.region_hist(labs, ids, nr, lum, col, exact=False)
    hist, w = K.region_hist([_dev(f) for f in clip], [_dev(m) for m in ids], nr, lum, col)
    hist, w = hist.cpu().numpy(), w.cpu().numpy()
    assert np.array_equal(w, wsum.astype(np.float32))
    assert np.abs(hist - exact).max() <= 1e-7
    assert np.abs(hist - like_ref).max() <= 2e-5
    assert np.array_equal(hist > 0, exact > 0) or np.abs(hist - exact)[(hist > 0) != (exact > 0)].max() <= 2e-8
    # empty region: all-zero histogram, zero weight
    hist2, w2 = K.region_hist([_dev(clip[0])], [_dev(np.zeros((120, 160), np.int32))], 2, lum, col)
    assert float(w2[1]) == 0 and float(hist2[1].abs().sum()) == 0 and abs(float(hist2[0].sum()) - 1) <= 1e-5
    # chi-square distances between neighbouring regions
    pairs = np.int32([[0, 1], [1, 0], [2, 2], [3, 22], [5, 9]])
    d = K.hist_chisquare(_dev(hist), _dev(pairs)).cpu().numpy()
    assert np.abs(d - ob.hist_chisquare(hist, pairs)).max() <= 1e-6
    assert d[2] == 0 and abs(d[0] - d[1]) <= 1e-7


def test_region_hist_on_the_engine_output(K):
    """End of the dense stage -> start of the region stage: id maps from the streaming engine feed the descriptor."""
    from video_segment_b200.unit import DenseSegmentationUnit
    clip = synth_clip(14, 160, 120, 6)
    u = DenseSegmentationUnit(want_id_maps=True)
    assert u.open_streams(160, 120)
    out = []
    for f in clip:
        out += u.process_frame(f)
    out += u.post_process()
    u.close()
    nr = 1 + max(int(o["id_map"].max()) for o in out)
    ids = [o["id_map"].astype(np.int32) for o in out]
    hist, w = K.region_hist([_dev(f) for f in clip], [_dev(m) for m in ids], nr)
    exact, wsum = ob.region_hist([ob.bgr2lab(f) for f in clip], ids, nr, exact=True)
    assert np.array_equal(w.cpu().numpy(), wsum.astype(np.float32)) and int(wsum.sum()) == 6 * 160 * 120
    assert np.abs(hist.cpu().numpy() - exact).max() <= 1e-7


def _moments_f32(ys, lxs, rxs):
    """ShapeMomentsFromRasterization (segment_util/segmentation_util.cpp:652-693) in float32, interval by interval."""
    f = np.float32
    area = sx = sy = sxx = syy = sxy = f(0)
    for y, lx, rx in zip(ys, lxs, rxs):
        m, n, cy = f(lx), f(rx), f(y)
        ln = f(n - m + f(1))
        area = f(area + ln)
        cx = f(f(n + m) * f(0.5))
        row_x, row_y = f(cx * ln), f(cy * ln)
        sx, sy = f(sx + row_x), f(sy + row_y)
        sxy, syy = f(sxy + f(cy * row_x)), f(syy + f(cy * row_y))
        poly = f(f(f(f(-m + f(f(f(2) * m) * m)) + n) + f(f(f(2) * m) * n)) + f(f(f(2) * n) * n))
        sxx = f(sxx + f(f(ln * poly) / f(6)))
    inv = f(f(1) / area)
    return [f(sx * inv), f(sy * inv), f(sxx * inv), f(sxy * inv), f(syy * inv)]


@pytest.mark.parametrize("shape,n_labels,seed", [((3, 40, 56), 4, 1), ((2, 97, 131), 7, 2), ((1, 8, 2100), 3, 3), ((4, 33, 33), 2, 4)])
def test_label_components_and_moments(K, shape, n_labels, seed):
    """K11 + K10 (csrc/shape.cu): components of every label in every frame against scipy's 4-connected labelling, numbered
    in the order of their first scan interval; area, interval count and the float32 moments of every component
    against an interval-by-interval restatement of ShapeMomentsFromRasterization (bit exact).  The engine tests
    compare the same numbers, as fields of the output message, with the oracle and the reference."""
    from scipy import ndimage
    rng = np.random.default_rng(seed)
    s, h, w = shape
    coarse = rng.integers(0, n_labels, (s, (h + 5) // 6, (w + 5) // 6))
    labels = np.kron(coarse, np.ones((1, 6, 6), np.int64))[:, :h, :w]
    noise = rng.random((s, h, w)) < 0.08                         # specks: many small components
    labels = np.where(noise, rng.integers(0, n_labels, (s, h, w)), labels).astype(np.int32)
    comp, rec = K.label_components(_dev(labels))
    comp = comp.cpu().numpy()
    # expected numbering: components of all labels, ordered by (slice, first pixel in raster order)
    expected = np.full((s, h, w), -1, np.int64)
    firsts = []
    for k in range(s):
        for lab in range(n_labels):
            lab_map, n = ndimage.label(labels[k] == lab, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])
            for c in range(1, n + 1):
                first = np.flatnonzero(lab_map.ravel() == c)[0]
                firsts.append((k, first, lab, c))
    firsts.sort()
    per_slice_maps = {}
    for idx, (k, first, lab, c) in enumerate(firsts):
        if (k, lab) not in per_slice_maps:
            per_slice_maps[(k, lab)] = ndimage.label(labels[k] == lab, structure=[[0, 1, 0], [1, 1, 1], [0, 1, 0]])[0]
        expected[k][per_slice_maps[(k, lab)] == c] = idx
    assert len(firsts) == len(rec["label"])
    assert np.array_equal(comp, expected)
    assert np.array_equal(rec["label"], [f[2] for f in firsts]) and np.array_equal(rec["slice"], [f[0] for f in firsts])
    for idx in range(len(firsts)):
        mask = expected[rec["slice"][idx]] == idx
        assert rec["area"][idx] == mask.sum()
        ys, lxs, rxs = [], [], []
        for y in np.flatnonzero(mask.any(axis=1)):
            row = np.flatnonzero(np.diff(np.concatenate([[0], mask[y].astype(np.int8), [0]])))
            for a, b in zip(row[0::2], row[1::2]):
                ys.append(y); lxs.append(a); rxs.append(b - 1)
        assert rec["count"][idx] == len(ys)
        want = np.array(_moments_f32(ys, lxs, rxs), np.float32)
        assert np.array_equal(rec["moments"][idx].view(np.uint32), want.view(np.uint32)), (idx, rec["moments"][idx], want)
