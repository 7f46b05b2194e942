"""ctypes binding of the CPU oracle (oracle/libvso.so).  TEST INFRASTRUCTURE ONLY:
imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; never by the product package."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_LIB_PATH = os.path.join(_ROOT, "oracle", "libvso.so")


class DenseOpts(C.Structure):
    _fields_ = [
        ("presmoothing", C.c_int32), ("frac_min_region_size", C.c_float),
        ("chunk_size", C.c_int32), ("chunk_overlap_ratio", C.c_float),
        ("num_constraint_frames", C.c_int32), ("enforce_n4_connectivity", C.c_int32),
        ("enforce_spatial_connectedness", C.c_int32), ("color_distance", C.c_int32),
        ("num_threads", C.c_int32),
    ]


class FrameResult(C.Structure):
    """Layout shared by vso_frame_result and vsb200_frame_result."""
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("chunk_id", C.c_int32),
        ("chunk_size", C.c_int32), ("overlap_start", C.c_int32),
        ("hierarchy_frame_idx", C.c_int32), ("connectedness", C.c_int32),
        ("n_regions", C.c_int32),
        ("region_id", C.POINTER(C.c_int32)), ("interval_offset", C.POINTER(C.c_int32)),
        ("intervals", C.POINTER(C.c_int32)), ("shape_moments", C.POINTER(C.c_float)),
        ("n_compound", C.c_int32),
        ("compound", C.POINTER(C.c_int32)), ("neighbor_offset", C.POINTER(C.c_int32)),
        ("neighbor_id", C.POINTER(C.c_int32)),
        ("pts", C.c_int64),
    ]


def result_to_dict(r: FrameResult) -> dict:
    """Deep-copies a popped frame result into numpy arrays."""
    n = r.n_regions
    def arr(p, cnt, dt):
        if cnt == 0:
            return np.zeros(0, dt)
        return np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True)
    off = arr(r.interval_offset, n + 1, np.int32)
    nint = int(off[-1]) if n > 0 else 0
    nc = r.n_compound
    noff = arr(r.neighbor_offset, nc + 1, np.int32) if nc > 0 else np.zeros(1, np.int32)
    return dict(
        width=r.width, height=r.height, chunk_id=r.chunk_id, chunk_size=r.chunk_size,
        overlap_start=r.overlap_start, hierarchy_frame_idx=r.hierarchy_frame_idx,
        connectedness=r.connectedness, pts=r.pts,
        region_id=arr(r.region_id, n, np.int32), interval_offset=off,
        intervals=arr(r.intervals, 3 * nint, np.int32).reshape(-1, 3),
        shape_moments=arr(r.shape_moments, 6 * n, np.float32).reshape(-1, 6),
        compound=arr(r.compound, 4 * nc, np.int32).reshape(-1, 4),
        neighbor_offset=noff,
        neighbor_id=arr(r.neighbor_id, int(noff[-1]), np.int32),
    )


def id_map_from_result(d: dict) -> np.ndarray:
    """Renders the per-frame region-id map (SegmentationDescToIdImage semantics)."""
    img = np.full((d["height"], d["width"]), -1, np.int32)
    off = d["interval_offset"]
    for k, rid in enumerate(d["region_id"]):
        for y, lx, rx in d["intervals"][off[k]:off[k + 1]]:
            img[y, lx:rx + 1] = rid
    return img


def build_oracle(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "libvso.so"])
    return _LIB_PATH


def host_tubes_check(seed: int, n_cases: int, with_flow: bool):
    """(#mismatching volumes, first mismatch or a summary) of the product's tube split (csrc/tubes.hpp) against the
    oracle's EnforceSpatialConnectedness on random label volumes (tests/host_tubes_check.cpp, oracle/libtubes_check.so)."""
    path = os.path.join(_ROOT, "oracle", "libtubes_check.so")
    if not os.path.exists(path):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "libvso.so", "libtubes_check.so"])
    L = C.CDLL(path)
    L.host_tubes_check.argtypes = [C.c_uint, C.c_int, C.c_int, C.c_char_p, C.c_int]
    msg = C.create_string_buffer(400)
    bad = L.host_tubes_check(seed, n_cases, int(with_flow), msg, 400)
    return bad, msg.value.decode()


_lib = None


def lib():
    global _lib
    if _lib is None:
        build_oracle()
        L = C.CDLL(_LIB_PATH)
        f32p, u8p, i32p = C.POINTER(C.c_float), C.POINTER(C.c_uint8), C.POINTER(C.c_int32)
        L.vso_default_opts.argtypes = [C.POINTER(DenseOpts)]
        L.vso_convert_u8_to_f32.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vso_bilateral.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_float, C.c_void_p,
                                    C.c_int, C.c_void_p, C.c_void_p]
        L.vso_preprocess.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.vso_spatial_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vso_temporal_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vso_bucket_index.argtypes = [C.c_float]
        L.vso_bucket_index.restype = C.c_int
        L.vso_dense_create.argtypes = [C.POINTER(DenseOpts), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
        L.vso_dense_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int64, C.POINTER(C.c_int)]
        L.vso_dense_flush.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
        L.vso_dense_pop.argtypes = [C.c_void_p, C.POINTER(FrameResult)]
        L.vso_dense_destroy.argtypes = [C.c_void_p]
        L.vso_dense_last_chunk_slots.argtypes = [C.c_void_p]
        L.vso_dense_last_chunk_node_labels.argtypes = [C.c_void_p]
        L.vso_dense_last_chunk_node_labels.restype = C.POINTER(C.c_int32)
        L.vso_dense_last_chunk_id_images.argtypes = [C.c_void_p]
        L.vso_dense_last_chunk_id_images.restype = C.POINTER(C.c_int32)
        L.vso_dense_last_chunk_merge_stats.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
        L.vso_dense_stage_seconds.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.vso_segment_chunk_labels.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vso_bgr2lab.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vso_region_hist_add.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_void_p, C.c_void_p]
        L.vso_hist_normalize.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.vso_hist_chisquare.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def default_opts(**kw) -> DenseOpts:
    o = DenseOpts()
    lib().vso_default_opts(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


def convert_u8(bgr: np.ndarray) -> np.ndarray:
    h, w, _ = bgr.shape
    bgr = np.ascontiguousarray(bgr)
    out = np.empty((h, w, 3), np.float32)
    lib().vso_convert_u8_to_f32(bgr.ctypes.data, w, h, w * 3, out.ctypes.data)
    return out


def bilateral(img: np.ndarray, sigma_space=3.0, sigma_color=0.25, threads=1, want_lut=False):
    h, w, _ = img.shape
    img = np.ascontiguousarray(img, np.float32)
    out = np.empty_like(img)
    lut = np.empty(12288, np.float32)
    scale = C.c_float()
    lib().vso_bilateral(img.ctypes.data, w, h, sigma_space, sigma_color, out.ctypes.data, threads,
                        lut.ctypes.data, C.addressof(scale))
    if want_lut:
        return out, lut, scale.value
    return out


def preprocess(bgr: np.ndarray, presmoothing=2, threads=1) -> np.ndarray:
    h, w, _ = bgr.shape
    bgr = np.ascontiguousarray(bgr)
    out = np.empty((h, w, 3), np.float32)
    lib().vso_preprocess(bgr.ctypes.data, w, h, w * 3, presmoothing, out.ctypes.data, threads)
    return out


def spatial_weights(img: np.ndarray, l1=False) -> np.ndarray:
    h, w, _ = img.shape
    out = np.empty((4, h, w), np.float32)
    lib().vso_spatial_weights(np.ascontiguousarray(img).ctypes.data, w, h, int(l1), out.ctypes.data)
    return out


def temporal_weights(curr, prev, flow=None, l1=False) -> np.ndarray:
    h, w, _ = curr.shape
    out = np.empty((9, h, w), np.float32)
    curr = np.ascontiguousarray(curr, np.float32)
    prev = np.ascontiguousarray(prev, np.float32)
    fl = None if flow is None else np.ascontiguousarray(flow, np.float32)
    lib().vso_temporal_weights(curr.ctypes.data, prev.ctypes.data, None if fl is None else fl.ctypes.data,
                               w, h, int(l1), out.ctypes.data)
    return out


def bucket_index(w: float) -> int:
    return lib().vso_bucket_index(float(w))


def bgr2lab(bgr: np.ndarray) -> np.ndarray:
    """cv::cvtColor(CV_BGR2Lab) on 8-bit data (region_descriptor.cpp:73)."""
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w, _ = bgr.shape
    out = np.empty((h, w, 3), np.uint8)
    lib().vso_bgr2lab(bgr.ctypes.data, w, h, w * 3, out.ctypes.data)
    return out


def region_hist(lab_frames, id_maps, n_regions: int, lum_bins: int = 10, color_bins: int = 20, exact: bool = False):
    """AppearanceDescriptor3D over the frames of a chunk set: returns (normalised histograms float32
    [n_regions, lum * col * col], weight sums float64 [n_regions]).  exact=False accumulates in float like the
    reference, exact=True in double."""
    total = lum_bins * color_bins * color_bins
    hist = np.zeros((n_regions, total), np.float64)
    wsum = np.zeros(n_regions, np.float64)
    for lab, ids in zip(lab_frames, id_maps):
        lab = np.ascontiguousarray(lab, np.uint8)
        ids = np.ascontiguousarray(ids, np.int32)
        h, w = ids.shape
        lib().vso_region_hist_add(lab.ctypes.data, ids.ctypes.data, w, h, n_regions, lum_bins, color_bins, int(exact),
                                  hist.ctypes.data, wsum.ctypes.data)
    out = np.empty((n_regions, total), np.float32)
    lib().vso_hist_normalize(hist.ctypes.data, wsum.ctypes.data, n_regions, total, int(exact), out.ctypes.data)
    return out, wsum


def hist_chisquare(hist: np.ndarray, pairs: np.ndarray) -> np.ndarray:
    """ColorHistogram::ChiSquareDist (histograms.cpp:391-407) for region pairs [n, 2]."""
    hist = np.ascontiguousarray(hist, np.float32)
    pairs = np.ascontiguousarray(pairs, np.int32)
    out = np.empty(len(pairs), np.float32)
    lib().vso_hist_chisquare(hist.ctypes.data, hist.shape[1], pairs.ctypes.data, len(pairs), out.ctypes.data)
    return out


def segment_chunk_labels(frames_f32: np.ndarray, min_region_size: int, l1=False) -> np.ndarray:
    t, h, w, _ = frames_f32.shape
    fr = np.ascontiguousarray(frames_f32, np.float32)
    out = np.empty((t, h, w), np.int32)
    lib().vso_segment_chunk_labels(fr.ctypes.data, w, h, t, int(l1), min_region_size, out.ctypes.data)
    return out


class OracleDense:
    """Streaming oracle == reference DenseSegmentation behind the same call shape
    as video_segment_b200.DenseSegmentationUnit."""

    def __init__(self, width, height, use_flow=False, **opts):
        self.w, self.h = width, height
        self.opts = default_opts(**opts)
        self._h = C.c_void_p()
        rc = lib().vso_dense_create(C.byref(self.opts), width, height, int(use_flow), C.byref(self._h))
        if rc != 0:
            raise ValueError("vso_dense_create failed")
        self._n = 0

    def push(self, bgr, flow=None, pts=None):
        bgr = np.ascontiguousarray(bgr)
        n = C.c_int()
        fl = None if flow is None else np.ascontiguousarray(flow, np.float32)
        lib().vso_dense_push(self._h, bgr.ctypes.data, self.w * 3, None if fl is None else fl.ctypes.data,
                             self.w * 8, self._n if pts is None else pts, C.byref(n))
        self._n += 1
        return self._pop(n.value)

    def flush(self):
        n = C.c_int()
        lib().vso_dense_flush(self._h, C.byref(n))
        return self._pop(n.value)

    def _pop(self, n):
        out = []
        for _ in range(n):
            r = FrameResult()
            assert lib().vso_dense_pop(self._h, C.byref(r)) == 0
            out.append(result_to_dict(r))
        return out

    def last_chunk_node_labels(self):
        s = lib().vso_dense_last_chunk_slots(self._h)
        p = lib().vso_dense_last_chunk_node_labels(self._h)
        return np.ctypeslib.as_array(p, shape=(s, self.h, self.w)).copy()

    def last_chunk_id_images(self):
        s = lib().vso_dense_last_chunk_slots(self._h)
        p = lib().vso_dense_last_chunk_id_images(self._h)
        return np.ctypeslib.as_array(p, shape=(s, self.h, self.w)).copy()

    def last_overlap_state(self):
        """(id maps [2, h, w] int32, [max region id, chunk id, frames output]) of the last chunk boundary."""
        p = C.POINTER(C.c_int32)()
        st = (C.c_int32 * 3)()
        L = lib()
        L.vso_dense_last_overlap_state.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.c_int32)]
        if L.vso_dense_last_overlap_state(self._h, C.byref(p), st) != 0:
            return None, None
        return np.ctypeslib.as_array(p, shape=(2, self.h, self.w)).copy(), [int(v) for v in st]

    def merge_stats(self):
        a = (C.c_int64 * 3)()
        lib().vso_dense_last_chunk_merge_stats(self._h, a)
        return list(a)

    def stage_seconds(self):
        a = (C.c_double * 5)()
        lib().vso_dense_stage_seconds(self._h, a)
        return list(a)

    def close(self):
        if self._h:
            lib().vso_dense_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def dict_to_result(d: dict):
    """FrameResult over numpy arrays (the arrays are returned too: keep them alive while the struct is in use)."""
    keep = dict(
        region_id=np.ascontiguousarray(d["region_id"], np.int32),
        interval_offset=np.ascontiguousarray(d["interval_offset"], np.int32),
        intervals=np.ascontiguousarray(d["intervals"], np.int32).reshape(-1),
        shape_moments=np.ascontiguousarray(d["shape_moments"], np.float32).reshape(-1),
        compound=np.ascontiguousarray(d["compound"], np.int32).reshape(-1),
        neighbor_offset=np.ascontiguousarray(d["neighbor_offset"], np.int32),
        neighbor_id=np.ascontiguousarray(d["neighbor_id"], np.int32),
    )
    r = FrameResult()
    for k in ("width", "height", "chunk_id", "chunk_size", "overlap_start", "hierarchy_frame_idx", "connectedness"):
        setattr(r, k, int(d[k]))
    r.pts = int(d.get("pts", 0))
    r.n_regions = len(keep["region_id"])
    r.n_compound = len(keep["compound"]) // 4
    i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    for k in ("region_id", "interval_offset", "intervals", "compound", "neighbor_offset", "neighbor_id"):
        setattr(r, k, keep[k].ctypes.data_as(i32p))
    r.shape_moments = keep["shape_moments"].ctypes.data_as(f32p)
    return r, keep


class OracleHier:
    """The oracle's hierarchical region stage (oracle/vso_hier.cpp), fed with over-segmentation results (dicts as
    produced by OracleDense or by the product's DenseSegmentationUnit) plus the frames they belong to."""

    def __init__(self, width, height, use_flow=False, chunk_set_size=6, chunk_set_overlap=2, constraint_chunks=1,
                 min_region_num=10, max_region_num=10000, level_cutoff_fraction=0.8, small_region_penalizer=0.25):
        L = lib()
        L.vso_hier_create.restype = C.c_void_p
        L.vso_hier_create.argtypes = [C.c_int] * 8 + [C.c_float, C.c_float]
        L.vso_hier_push.argtypes = [C.c_void_p, C.POINTER(FrameResult), C.c_void_p, C.c_void_p]
        L.vso_hier_flush.argtypes = [C.c_void_p]
        L.vso_hier_pop.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int32))]
        L.vso_hier_pop.restype = C.c_longlong
        L.vso_hier_destroy.argtypes = [C.c_void_p]
        self._h = L.vso_hier_create(width, height, int(use_flow), chunk_set_size, chunk_set_overlap, constraint_chunks,
                                    min_region_num, max_region_num, level_cutoff_fraction, small_region_penalizer)
        if not self._h:
            raise ValueError("vso_hier_create failed")

    def _pop(self, n):
        out = []
        for _ in range(n):
            p = C.POINTER(C.c_int32)()
            nw = lib().vso_hier_pop(self._h, C.byref(p))
            out.append(np.ctypeslib.as_array(p, shape=(nw,)).copy())
        return out

    def push(self, overseg: dict, bgr, flow=None):
        r, keep = dict_to_result(overseg)
        bgr = np.ascontiguousarray(bgr)
        fl = None if flow is None else np.ascontiguousarray(flow, np.float32)
        n = lib().vso_hier_push(self._h, C.byref(r), bgr.ctypes.data, None if fl is None else fl.ctypes.data)
        del keep
        return self._pop(n)

    def flush(self):
        return self._pop(lib().vso_hier_flush(self._h))

    def close(self):
        if self._h:
            lib().vso_hier_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
