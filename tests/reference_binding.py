"""ctypes binding of oracle/_ref/libref_results.so: the REFERENCE's own streaming over-segmentation
(DenseSegmentation / Segmentation / DenseSegmentationGraph / FastSegmentationGraph / BilateralFilter ...) compiled
unmodified from /root/reference by `make -C oracle _ref` (oracle/ref_results_wrap.cpp, stand-ins in oracle/ref_shim/).
TEST INFRASTRUCTURE ONLY: used by tests/ to pin the oracle and by bench.py's CPU legs (cpu_baseline kind "reference");
never by the product package.  The library is built where /root/reference exists and travels to the GPU box as a
prebuilt file; nothing here reads /root/reference at run time."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from oracle_binding import FrameResult, default_opts, result_to_dict

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_ROOT, "oracle", "_ref", "libref_results.so")
_lib = None


def available(build: bool = True) -> bool:
    """True if the compiled reference is there (built on demand where the reference sources are mounted)."""
    if build and os.path.isdir("/root/reference"):
        subprocess.check_call(["make", "-C", os.path.join(_ROOT, "oracle"), "_ref"], stdout=subprocess.DEVNULL)
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(LIB_PATH)
        L.ref_dense_create.restype = C.c_void_p
        L.ref_dense_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_int]
        L.ref_dense_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.ref_dense_flush.argtypes = [C.c_void_p]
        L.ref_dense_pop.argtypes = [C.c_void_p, C.c_void_p]
        L.ref_dense_destroy.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class ReferenceDense:
    """The reference's DenseSegmentation behind the call shape of oracle_binding.OracleDense."""

    def __init__(self, width, height, use_flow=False, **opts):
        self.w, self.h, self.use_flow = width, height, use_flow
        o = default_opts(**opts)
        if o.presmoothing == 1:
            raise ValueError("PRESMOOTH_GAUSSIAN needs cv::GaussianBlur (third party, not compiled in)")
        if o.chunk_size < 3 or min(int(o.chunk_overlap_ratio * o.chunk_size + 0.5), 2) < 2:
            # dense_segmentation.cpp:58-62 takes min(overlap, 2): with fewer than 2 overlap frames the reference reads
            # overlap_segmentations_[1] out of bounds at the first chunk boundary (:304-306) and crashes.
            raise ValueError("the reference needs at least 2 overlap frames (chunk_overlap_ratio * chunk_size >= 1.5)")
        self._h = lib().ref_dense_create(o.presmoothing, o.frac_min_region_size, o.chunk_size, o.chunk_overlap_ratio,
                                         o.num_constraint_frames, o.enforce_n4_connectivity,
                                         o.enforce_spatial_connectedness, o.color_distance, width, height, int(use_flow))
        self._n = 0

    def push(self, bgr, flow=None, pts=None):
        bgr = np.ascontiguousarray(bgr)
        fl = None
        if self.use_flow and self._n > 0:
            fl = np.ascontiguousarray(flow, np.float32)
        n = lib().ref_dense_push(self._h, bgr.ctypes.data, self.w * 3, None if fl is None else fl.ctypes.data, self.w * 8)
        self._n += 1
        return self._pop(n)

    def flush(self):
        return self._pop(lib().ref_dense_flush(self._h))

    def _pop(self, n):
        out = []
        for _ in range(n):
            r = FrameResult()
            assert lib().ref_dense_pop(self._h, C.byref(r)) == 0
            out.append(result_to_dict(r))
        return out

    def close(self):
        if self._h:
            lib().ref_dense_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- the C++ host side (video_segment_b200/host/b200_dense_segmentation.*) through tests/host_check_wrap.cpp ----

HOST_LIB_PATH = os.path.join(_ROOT, "oracle", "_ref", "libb200_host_check.so")
_host = None


def host_available(build: bool = True) -> bool:
    available(build)
    return os.path.exists(HOST_LIB_PATH) and os.path.exists(LIB_PATH)


def host_lib():
    global _host
    if _host is None:
        L = C.CDLL(HOST_LIB_PATH)
        L.host_check_desc_vs_reference.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                                   C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_int]
        L.b200_dense_create.restype = C.c_void_p
        L.b200_dense_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int]
        L.b200_dense_push.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.b200_dense_flush.argtypes = [C.c_void_p]
        L.b200_dense_pop.argtypes = [C.c_void_p, C.c_void_p]
        L.b200_dense_kernel_launches.argtypes = [C.c_void_p]
        L.b200_dense_kernel_launches.restype = C.c_longlong
        L.b200_dense_destroy.argtypes = [C.c_void_p]
        _host = L
    return _host


def host_check_desc_vs_reference(clip, flows, **opts):
    """(#frames whose rebuilt SegmentationDesc differs from the reference's own, first difference)."""
    o = default_opts(**opts)
    clip = np.ascontiguousarray(clip)
    t, h, w, _ = clip.shape
    fl = None if flows is None else np.ascontiguousarray(flows, np.float32)
    msg = C.create_string_buffer(256)
    bad = host_lib().host_check_desc_vs_reference(
        clip.ctypes.data, None if fl is None else fl.ctypes.data, t, w, h, o.presmoothing, o.frac_min_region_size,
        o.chunk_size, o.chunk_overlap_ratio, o.num_constraint_frames, o.enforce_n4_connectivity,
        o.enforce_spatial_connectedness, o.color_distance, msg, 256)
    return bad, msg.value.decode()


class B200HostDense:
    """segmentation::B200DenseSegmentation (the C++ host class over libvsb200.so) behind the OracleDense call shape.
    Needs a B200: construction succeeds anywhere, the first push aborts without a device (no CPU fallback)."""

    def __init__(self, width, height, use_flow=False, **opts):
        self.w, self.h, self.use_flow = width, height, use_flow
        o = default_opts(**opts)
        self._h = host_lib().b200_dense_create(o.presmoothing, o.frac_min_region_size, o.chunk_size, o.chunk_overlap_ratio,
                                               o.num_constraint_frames, o.enforce_n4_connectivity,
                                               o.enforce_spatial_connectedness, o.color_distance, width, height, int(use_flow))
        self._n = 0

    def push(self, bgr, flow=None, pts=None):
        bgr = np.ascontiguousarray(bgr)
        fl = None
        if self.use_flow and self._n > 0:
            fl = np.ascontiguousarray(flow, np.float32)
        n = host_lib().b200_dense_push(self._h, bgr.ctypes.data, self.w * 3, None if fl is None else fl.ctypes.data, self.w * 8)
        self._n += 1
        return self._pop(n)

    def flush(self):
        return self._pop(host_lib().b200_dense_flush(self._h))

    def kernel_launches(self):
        return int(host_lib().b200_dense_kernel_launches(self._h))

    def _pop(self, n):
        out = []
        for _ in range(n):
            r = FrameResult()
            assert host_lib().b200_dense_pop(self._h, C.byref(r)) == 0
            out.append(result_to_dict(r))
        return out

    def close(self):
        if self._h:
            host_lib().b200_dense_destroy(self._h)
            self._h = None


# ---- the reference's own SegmentationWriter / SegmentationReader / StripToEssentials (segmentation_io.cpp) ----

def _io_lib():
    L = host_lib()
    if not getattr(L, "_io_ready", False):
        L.ref_io_write.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.c_char_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.ref_io_read.argtypes = [C.c_char_p, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_size_t, C.c_void_p,
                                  C.c_void_p, C.c_int]
        L.ref_io_strip.argtypes = [C.POINTER(FrameResult), C.c_int, C.c_void_p, C.c_size_t]
        L.ref_io_strip.restype = C.c_longlong
        L._io_ready = True
    return L


def ref_io_write(filename, header_entries, payloads, pts, chunk_every=0):
    ent = np.asarray(header_entries, np.int32)
    sizes = np.asarray([len(p) for p in payloads], np.int64)
    ts = np.asarray(pts, np.int64)
    blob = b"".join(payloads)
    rc = _io_lib().ref_io_write(filename.encode(), ent.ctypes.data, len(ent), blob, sizes.ctypes.data, ts.ctypes.data,
                                len(payloads), chunk_every)
    assert rc == 0


def ref_io_read(filename, max_frames=4096, max_bytes=1 << 26):
    flags = np.zeros(64, np.int32)
    nf = C.c_int()
    blob = np.zeros(max_bytes, np.uint8)
    sizes = np.zeros(max_frames, np.int64)
    ts = np.zeros(max_frames, np.int64)
    n = _io_lib().ref_io_read(filename.encode(), flags.ctypes.data, 64, C.byref(nf), blob.ctypes.data, max_bytes,
                              sizes.ctypes.data, ts.ctypes.data, max_frames)
    assert n >= 0
    out, pos = [], 0
    for k in range(n):
        out.append(blob[pos:pos + sizes[k]].tobytes())
        pos += int(sizes[k])
    return flags[:nf.value].tolist(), out, ts[:n].tolist()


def result_struct(d: dict):
    """A FrameResult structure over the arrays of a result dict (keeps them alive through ._keep)."""
    r = FrameResult()
    for k in ("width", "height", "chunk_id", "chunk_size", "overlap_start", "hierarchy_frame_idx", "connectedness"):
        setattr(r, k, int(d[k]))
    keep = {k: np.ascontiguousarray(d[k], np.float32 if k == "shape_moments" else np.int32)
            for k in ("region_id", "interval_offset", "intervals", "shape_moments", "compound", "neighbor_offset", "neighbor_id")}
    r.n_regions = len(keep["region_id"])
    r.n_compound = len(keep["compound"])
    for k, a in keep.items():
        setattr(r, k, a.ctypes.data_as(C.POINTER(C.c_float if k == "shape_moments" else C.c_int32)))
    r.pts = int(d.get("pts", 0))
    r._keep = keep
    return r


def ref_io_strip(d: dict, save_shape_moments: bool) -> bytes:
    r = result_struct(d)
    n = _io_lib().ref_io_strip(C.byref(r), int(save_shape_moments), None, 0)
    buf = C.create_string_buffer(max(1, n))
    _io_lib().ref_io_strip(C.byref(r), int(save_shape_moments), buf, n)
    return buf.raw[:n]


def host_shape_check(seed: int, n_cases: int):
    """(#mismatches, first mismatch) of csrc/shape_math.hpp + csrc/region_raster.hpp against the reference's segmentation_util.cpp functions on
    random rasters (tests/host_shape_check.cpp)."""
    L = host_lib()
    L.host_shape_check.argtypes = [C.c_uint, C.c_int, C.c_char_p, C.c_int]
    msg = C.create_string_buffer(256)
    bad = L.host_shape_check(seed, n_cases, msg, 256)
    return bad, msg.value.decode()
