"""Host-side mirror of the reference's plug point for this path: ``DenseSegmentationUnit``
(reference segmentation/segmentation_unit.h:63-124, segmentation_unit.cpp:48-178) and
``DenseSegmentationOptions`` (segmentation/dense_segmentation.h:42-95), over the C ABI.

Same names, argument meaning and error behaviour as the reference:
  * construction + ``open_streams(width, height, pixel_format, flow=False)`` <-> OpenStreams():
    returns False (and logs) on a bad stream set-up instead of raising;
  * ``process_frame(bgr, flow=None, pts=...)`` <-> ProcessFrame(): buffers the frame and returns
    the list of frame results that became available (chunk latency, input order);
  * ``post_process()`` <-> PostProcess(): flushes the remaining frames.
Every result is a dict with exactly the SegmentationDesc fields (segmentation.proto:55-172).
There is no CPU fallback: without the CUDA library / a B200 the unit cannot be opened.
"""
from __future__ import annotations

import ctypes as C
import logging
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from ._lib import DenseOpts, FrameResult, check, lib

log = logging.getLogger("video_segment_b200")

PIXEL_FORMAT_BGR24 = "BGR24"

PRESMOOTH_NONE, PRESMOOTH_GAUSSIAN, PRESMOOTH_BILATERAL = 0, 1, 2
COLOR_DISTANCE_L1, COLOR_DISTANCE_L2 = 0, 1


@dataclass
class DenseSegmentationOptions:
    """Field-for-field DenseSegmentationOptions (dense_segmentation.h:42-95)."""
    presmoothing: int = PRESMOOTH_BILATERAL
    frac_min_region_size: float = 0.01
    chunk_size: int = 20
    chunk_overlap_ratio: float = 0.2
    two_stage_oversegment: bool = False
    num_constraint_frames: int = 1
    thin_structure_suppression: bool = False
    enforce_n4_connectivity: bool = True
    enforce_spatial_connectedness: bool = True
    color_distance: int = COLOR_DISTANCE_L2
    compute_vectorization: bool = False


@dataclass
class DenseSegmentationUnitOptions:
    """segmentation_unit.h:52-59."""
    video_stream_name: str = "VideoStream"
    flow_stream_name: str = "BackwardFlowStream"
    segment_stream_name: str = "SegmentationStream"


def _result_to_dict(r: FrameResult) -> dict:
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


class DenseSegmentationUnit:
    def __init__(self, options: Optional[DenseSegmentationUnitOptions] = None,
                 dense_seg_options: Optional[DenseSegmentationOptions] = None,
                 device: int = 0, want_id_maps: bool = False, want_proto: bool = False):
        self.options = options or DenseSegmentationUnitOptions()
        self.dense_seg_options = dense_seg_options or DenseSegmentationOptions()
        self.device = device
        self.want_id_maps = want_id_maps
        self._imported = False
        self.want_proto = want_proto
        # optional segio.SegmentationWriter: every output frame is appended to its current chunk as it is popped, which
        # is what a SegmentationWriterUnit placed behind this unit does (segmentation_unit.cpp:373-410)
        self.segmentation_writer = None
        self._h = C.c_void_p()
        self.frame_width = self.frame_height = 0
        self.input_frames = self.output_frames = 0
        self._use_flow = False

    # --- OpenStreams (segmentation_unit.cpp:58-105) ---
    def open_streams(self, frame_width: int, frame_height: int, pixel_format: str = PIXEL_FORMAT_BGR24,
                     flow_stream_present: bool = False) -> bool:
        if pixel_format != PIXEL_FORMAT_BGR24:
            log.error("Expecting video format to be BGR24.")
            return False
        o = DenseOpts()
        lib().vsb200_dense_default_opts(C.byref(o))
        d = self.dense_seg_options
        o.presmoothing = int(d.presmoothing)
        o.frac_min_region_size = float(d.frac_min_region_size)
        o.chunk_size = int(d.chunk_size)
        o.chunk_overlap_ratio = float(d.chunk_overlap_ratio)
        o.num_constraint_frames = int(d.num_constraint_frames)
        o.two_stage_oversegment = int(d.two_stage_oversegment)
        o.thin_structure_suppression = int(d.thin_structure_suppression)
        o.enforce_n4_connectivity = int(d.enforce_n4_connectivity)
        o.enforce_spatial_connectedness = int(d.enforce_spatial_connectedness)
        o.color_distance = int(d.color_distance)
        o.compute_vectorization = int(d.compute_vectorization)
        o.device = int(self.device)
        o.want_id_maps = int(self.want_id_maps)
        rc = lib().vsb200_dense_create(C.byref(o), frame_width, frame_height, int(flow_stream_present), C.byref(self._h))
        if rc != 0:
            log.error("Could not open dense segmentation: %s", lib().vsb200_last_error().decode(errors="replace"))
            self._h = C.c_void_p()
            return False
        self.frame_width, self.frame_height = frame_width, frame_height
        self._use_flow = flow_stream_present
        return True

    def _collect(self, n: int) -> List[dict]:
        out = []
        for _ in range(n):
            r = FrameResult()
            check(lib().vsb200_dense_pop(self._h, C.byref(r)), "vsb200_dense_pop")
            d = _result_to_dict(r)
            if self.want_id_maps:
                p = lib().vsb200_dense_last_id_map(self._h)
                d["id_map"] = np.ctypeslib.as_array(p, shape=(r.height, r.width)).copy()
            if self.want_proto:
                nb = lib().vsb200_dense_last_proto(self._h, None, 0)
                buf = (C.c_uint8 * nb)()
                lib().vsb200_dense_last_proto(self._h, buf, nb)
                d["proto"] = bytes(buf)
            if self.segmentation_writer is not None:
                self.segmentation_writer.add_segmentation_to_chunk(self, r.pts)
            out.append(d)
            self.output_frames += 1
        if n:
            log.info("__STREAMING_SIZE__: %d", self.output_frames)     # segmentation_unit.cpp:177
        return out

    # --- ProcessFrame (segmentation_unit.cpp:118-142) ---
    def process_frame(self, bgr: np.ndarray, flow: Optional[np.ndarray] = None, pts: Optional[int] = None,
                      width_step: Optional[int] = None) -> List[dict]:
        if not self._h:
            raise RuntimeError("open_streams() was not called or failed")
        if bgr.dtype != np.uint8 or bgr.ndim != 3 or bgr.shape[2] != 3 or bgr.shape[0] != self.frame_height \
                or bgr.shape[1] != self.frame_width:
            raise ValueError("frame must be uint8 (H, W, 3) BGR of the stream's size")
        if not bgr.flags["C_CONTIGUOUS"] and width_step is None:
            bgr = np.ascontiguousarray(bgr)
        stride = width_step if width_step is not None else bgr.strides[0]
        fl_ptr, fl_stride = None, 0
        if self._use_flow and (self.input_frames > 0 or self._imported):
            if flow is None:
                raise ValueError("Flow always has to be passed or be absent.")      # dense_segmentation.cpp:139
            flow = np.ascontiguousarray(flow, np.float32)
            fl_ptr, fl_stride = flow.ctypes.data, flow.strides[0]
        n = C.c_int()
        check(lib().vsb200_dense_push(self._h, bgr.ctypes.data, stride, fl_ptr, fl_stride,
                                      self.input_frames if pts is None else pts, C.byref(n)), "vsb200_dense_push")
        self.input_frames += 1
        return self._collect(n.value)

    def process_device_frame(self, dev_ptr: int, row_stride_bytes: int, pts: Optional[int] = None) -> List[dict]:
        """Same as process_frame for a BGR24 frame already resident in device memory."""
        n = C.c_int()
        check(lib().vsb200_dense_push_device(self._h, C.c_void_p(dev_ptr), row_stride_bytes,
                                             self.input_frames if pts is None else pts, C.byref(n)),
              "vsb200_dense_push_device")
        self.input_frames += 1
        return self._collect(n.value)

    def export_halo(self, dev_prev_ptr: int, dev_last_ptr: int) -> List[int]:
        """Copies the two overlap frames' region-id maps into device buffers (int32 [h*w] each); returns the
        chain state [max region id, id of the chunk the maps constrain, frames output so far]."""
        st = (C.c_int32 * 3)()
        check(lib().vsb200_dense_export_halo(self._h, C.c_void_p(dev_prev_ptr), C.c_void_p(dev_last_ptr), st),
              "vsb200_dense_export_halo")
        return [int(v) for v in st]

    def import_halo(self, dev_prev_ptr: int, dev_last_ptr: int, chain_state) -> None:
        """Successor side of the seam: must precede the first frame; the first frame pushed afterwards has to be
        the predecessor's last pushed frame (the frame of the second id map)."""
        st = (C.c_int32 * 3)(*[int(v) for v in chain_state])
        check(lib().vsb200_dense_import_halo(self._h, C.c_void_p(dev_prev_ptr), C.c_void_p(dev_last_ptr), st),
              "vsb200_dense_import_halo")
        self._imported = True      # the group's first frame carries a flow field like any later frame

    def set_profiling(self, time_edge_kernel: bool = True) -> None:
        lib().vsb200_dense_set_profiling(self._h, int(time_edge_kernel))

    def io_stats(self) -> dict:
        a = (C.c_double * 4)()
        lib().vsb200_dense_io_stats(self._h, a)
        return dict(h2d_bytes=a[0], d2h_bytes=a[1], edge_ms=a[2], edge_launches=a[3])

    # --- PostProcess (segmentation_unit.cpp:154-161) ---
    def post_process(self) -> List[dict]:
        n = C.c_int()
        check(lib().vsb200_dense_flush(self._h, C.byref(n)), "vsb200_dense_flush")
        return self._collect(n.value)

    def stats(self) -> dict:
        a = (C.c_double * 9)()
        lib().vsb200_dense_stats(self._h, a)
        keys = ["h2d_preprocess_edges_ms", "unused", "sort_ms", "merge_ms", "labels_n4_rle_ms", "host_shape_ms",
                "neighbors_ms", "kernel_launches", "merge_rounds"]
        return dict(zip(keys, list(a)))

    def close(self):
        if self._h:
            lib().vsb200_dense_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def id_map_from_result(d: dict) -> np.ndarray:
    """SegmentationDescToIdImage (segment_util/segmentation_util.cpp:741-770) on a result dict."""
    img = np.full((d["height"], d["width"]), -1, np.int32)
    off = d["interval_offset"]
    for k, rid in enumerate(d["region_id"]):
        for y, lx, rx in d["intervals"][off[k]:off[k + 1]]:
            img[y, lx:rx + 1] = rid
    return img


# ---------------------------------------------------------------------------------------------
# RegionSegmentationUnit (reference segmentation/segmentation_unit.h:126-190, segmentation_unit.cpp:180-331)
# ---------------------------------------------------------------------------------------------
@dataclass
class RegionSegmentationOptions:
    """Field-for-field RegionSegmentationOptions (segmentation/region_segmentation.h:41-82)."""
    min_region_num: int = 10
    max_region_num: int = 10000
    level_cutoff_fraction: float = 0.8
    small_region_penalizer: float = 0.25
    luminance_bins: int = 10
    color_bins: int = 20
    flow_bins: int = 16
    chunk_set_size: int = 6
    chunk_set_overlap: int = 2
    constraint_chunks: int = 1
    save_descriptors: bool = False
    use_appearance: bool = True
    use_flow: bool = True
    use_size_penalizer: bool = True
    compute_vectorization: bool = False      # the reference defaults to True (cv::approxPolyDP, SURVEY row N3: not built)


def _dict_to_result(d: dict):
    """vsb200_frame_result over the arrays of an over-segmentation result dict (kept alive by the caller)."""
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


def parse_region_record(rec: np.ndarray) -> dict:
    """Flat record of vsb200_region_pop -> SegmentationDesc fields: over-segmentation of the frame plus, on the first
    frame of a chunk set, the hierarchy levels (lists of dicts id / size / parent_id / start_frame / end_frame /
    neighbors / children)."""
    d = dict(zip(("width", "height", "chunk_id", "chunk_size", "overlap_start", "hierarchy_frame_idx"), map(int, rec[:6])))
    n_regions, n_levels = int(rec[6]), int(rec[7])
    pos = 8
    ids, offs, ivs, moms = [], [0], [], []
    for _ in range(n_regions):
        ids.append(int(rec[pos]))
        n = int(rec[pos + 1])
        ivs.append(rec[pos + 2:pos + 2 + 3 * n].reshape(-1, 3))
        offs.append(offs[-1] + n)
        moms.append(rec[pos + 2 + 3 * n:pos + 8 + 3 * n].view(np.float32))
        pos += 8 + 3 * n
    d["region_id"] = np.asarray(ids, np.int32)
    d["interval_offset"] = np.asarray(offs, np.int32)
    d["intervals"] = np.concatenate(ivs).astype(np.int32) if ivs else np.zeros((0, 3), np.int32)
    d["shape_moments"] = np.stack(moms) if moms else np.zeros((0, 6), np.float32)
    levels = []
    for _ in range(n_levels):
        nc = int(rec[pos])
        pos += 1
        comps = []
        for _ in range(nc):
            cid, size, parent, start, end, nn, nch = map(int, rec[pos:pos + 7])
            pos += 7
            comps.append(dict(id=cid, size=size, parent_id=parent, start_frame=start, end_frame=end,
                              neighbors=rec[pos:pos + nn].tolist(), children=rec[pos + nn:pos + nn + nch].tolist()))
            pos += nn + nch
        levels.append(comps)
    d["levels"] = levels
    if pos != len(rec):
        raise ValueError("malformed region record")
    return d


class RegionSegmentationUnit:
    """Mirror of RegionSegmentationUnit: consumes the over-segmentation stream (the dicts a DenseSegmentationUnit
    returns) together with the video frames (and the flow stream when present), returns hierarchical results."""

    def __init__(self, region_options: Optional[RegionSegmentationOptions] = None, raw_records: bool = False):
        self.region_options = region_options or RegionSegmentationOptions()
        self._h = C.c_void_p()
        self._raw = raw_records
        self.frame_width = self.frame_height = 0
        self.input_frames = 0

    def open_streams(self, frame_width: int, frame_height: int, pixel_format: str = PIXEL_FORMAT_BGR24,
                     flow_stream_present: bool = False) -> bool:
        if pixel_format != PIXEL_FORMAT_BGR24:
            log.error("Expecting video format to be BGR24")                    # segmentation_unit.cpp:215-218
            return False
        from ._lib import RegionOpts
        o = RegionOpts()
        lib().vsb200_region_default_opts(C.byref(o))
        ro = self.region_options
        for k in ("min_region_num", "max_region_num", "luminance_bins", "color_bins", "flow_bins", "chunk_set_size",
                  "chunk_set_overlap", "constraint_chunks"):
            setattr(o, k, int(getattr(ro, k)))
        o.level_cutoff_fraction = float(ro.level_cutoff_fraction)
        o.small_region_penalizer = float(ro.small_region_penalizer)
        o.save_descriptors = int(ro.save_descriptors)
        o.use_appearance = int(ro.use_appearance)
        o.use_flow = int(ro.use_flow and flow_stream_present)                # CreateRegionSegmentation, :303-308
        o.use_size_penalizer = int(ro.use_size_penalizer)
        o.compute_vectorization = int(ro.compute_vectorization)
        rc = lib().vsb200_region_create(C.byref(o), frame_width, frame_height, C.byref(self._h))
        if rc != 0:
            log.error("RegionSegmentationUnit: %s", lib().vsb200_last_error().decode(errors="replace"))
            self._h = C.c_void_p()
            return False
        self.frame_width, self.frame_height = frame_width, frame_height
        self._use_flow = bool(o.use_flow)
        return True

    def _collect(self, n: int) -> List:
        out = []
        for _ in range(n):
            p = C.POINTER(C.c_int32)()
            nw = lib().vsb200_region_pop(self._h, C.byref(p))
            rec = np.ctypeslib.as_array(p, shape=(nw,)).copy()
            out.append(rec if self._raw else parse_region_record(rec))
        return out

    def process_frame(self, overseg: dict, bgr: np.ndarray, flow: Optional[np.ndarray] = None) -> List:
        if not self._h:
            raise RuntimeError("open_streams() was not called or failed")
        r, keep = _dict_to_result(overseg)
        bgr = np.ascontiguousarray(bgr)
        fl_ptr, fl_stride = None, 0
        if self._use_flow and self.input_frames > 0:
            if flow is None:
                raise ValueError("Flow always has to be passed or be absent.")
            flow = np.ascontiguousarray(flow, np.float32)
            fl_ptr, fl_stride = flow.ctypes.data, flow.strides[0]
        n = C.c_int()
        check(lib().vsb200_region_push(self._h, C.byref(r), bgr.ctypes.data, bgr.strides[0], fl_ptr, fl_stride, C.byref(n)),
              "vsb200_region_push")
        del keep
        self.input_frames += 1
        return self._collect(n.value)

    def post_process(self) -> List:
        n = C.c_int()
        check(lib().vsb200_region_flush(self._h, C.byref(n)), "vsb200_region_flush")
        return self._collect(n.value)

    def stats(self) -> dict:
        a = (C.c_double * 2)()
        lib().vsb200_region_stats(self._h, a)
        return dict(kernel_launches=a[0], chunk_sets=a[1])

    def close(self):
        if self._h:
            lib().vsb200_region_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
