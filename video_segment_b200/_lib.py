"""ctypes loader of the in-tree CUDA library (libvsb200.so).  Fails loudly when the
library is missing or cannot be loaded -- there is no CPU or PyTorch fallback."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VSB200_LIB") or os.path.join(HERE, "libvsb200.so")   # VSB200_LIB: development builds (tools/gpu_merge_sweep_constants.py)


class DenseOpts(C.Structure):
    """vsb200_dense_opts (include/vsb200.h) == DenseSegmentationOptions
    (reference segmentation/dense_segmentation.h:42-95)."""
    _fields_ = [
        ("presmoothing", C.c_int32), ("frac_min_region_size", C.c_float),
        ("chunk_size", C.c_int32), ("chunk_overlap_ratio", C.c_float),
        ("num_constraint_frames", C.c_int32), ("two_stage_oversegment", C.c_int32),
        ("thin_structure_suppression", C.c_int32), ("enforce_n4_connectivity", C.c_int32),
        ("enforce_spatial_connectedness", C.c_int32), ("color_distance", C.c_int32),
        ("compute_vectorization", C.c_int32), ("device", C.c_int32), ("want_id_maps", C.c_int32),
    ]


class RegionOpts(C.Structure):
    """vsb200_region_opts (include/vsb200.h) == RegionSegmentationOptions
    (reference segmentation/region_segmentation.h:41-82)."""
    _fields_ = [
        ("min_region_num", C.c_int32), ("max_region_num", C.c_int32),
        ("level_cutoff_fraction", C.c_float), ("small_region_penalizer", C.c_float),
        ("luminance_bins", C.c_int32), ("color_bins", C.c_int32), ("flow_bins", C.c_int32),
        ("chunk_set_size", C.c_int32), ("chunk_set_overlap", C.c_int32), ("constraint_chunks", C.c_int32),
        ("save_descriptors", C.c_int32), ("use_appearance", C.c_int32), ("use_flow", C.c_int32),
        ("use_size_penalizer", C.c_int32), ("compute_vectorization", C.c_int32), ("device", C.c_int32),
    ]


class FrameResult(C.Structure):
    """vsb200_frame_result (include/vsb200.h)."""
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


# every symbol include/vsb200.h declares
EXPORTED_SYMBOLS = [
    "vsb200_dense_default_opts", "vsb200_last_error", "vsb200_device_count",
    "vsb200_dense_create", "vsb200_dense_push", "vsb200_dense_push_device", "vsb200_dense_flush", "vsb200_dense_pop",
    "vsb200_dense_set_profiling", "vsb200_dense_io_stats",
    "vsb200_dense_last_id_map", "vsb200_dense_last_proto", "vsb200_dense_stats",
    "vsb200_dense_destroy", "vsb200_dense_export_halo", "vsb200_dense_import_halo",
    "vsb200_preprocess_scratch_bytes", "vsb200_preprocess", "vsb200_edge_build",
    "vsb200_bucket_index", "vsb200_sort_edges", "vsb200_sort_scratch_bytes", "vsb200_segment_chunk", "vsb200_label_components",
    "vsb200_bgr2lab", "vsb200_region_hist_scratch_bytes", "vsb200_region_hist_reset", "vsb200_region_hist_add",
    "vsb200_region_hist_finish", "vsb200_hist_chisquare",
    "vsb200_seg_writer_open", "vsb200_seg_writer_add", "vsb200_seg_writer_add_last_frame", "vsb200_seg_writer_write_chunk",
    "vsb200_seg_writer_close", "vsb200_seg_reader_open", "vsb200_seg_reader_num_frames", "vsb200_seg_reader_num_header_flags",
    "vsb200_seg_reader_header_flags", "vsb200_seg_reader_time_stamps", "vsb200_seg_reader_read", "vsb200_seg_reader_read_frame", "vsb200_seg_reader_close",
    "vsb200_strip_to_essentials", "vsb200_encode_frame_proto",
    "vsb200_region_default_opts", "vsb200_region_create", "vsb200_region_push", "vsb200_region_flush", "vsb200_region_pop",
    "vsb200_region_stats", "vsb200_region_destroy",
    "vsb200_shard_unique_id", "vsb200_shard_create", "vsb200_shard_exchange", "vsb200_shard_pred_maps", "vsb200_shard_vote",
    "vsb200_shard_relabel", "vsb200_shard_stats", "vsb200_shard_destroy",
]

_lib = None


def lib() -> C.CDLL:
    """Loads libvsb200.so; raises (never falls back) if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m video_segment_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU / PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    sigs = {
        "vsb200_last_error": (None, C.c_char_p),
        "vsb200_device_count": (None, C.c_int),
        "vsb200_dense_default_opts": ([C.POINTER(DenseOpts)], None),
        "vsb200_preprocess_scratch_bytes": (None, C.c_size_t),
        "vsb200_preprocess": ([vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, vp], C.c_int),
        "vsb200_edge_build": ([vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp, vp], C.c_int),
        "vsb200_bucket_index": ([C.c_float], C.c_int),
        "vsb200_sort_scratch_bytes": ([C.c_int, C.c_int, C.c_int], C.c_size_t),
        "vsb200_sort_edges": ([C.POINTER(vp), C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_size_t, vp], C.c_int),
        "vsb200_segment_chunk": ([vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_double), vp], C.c_int),
        "vsb200_label_components": ([vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.POINTER(C.c_int), vp], C.c_int),
        "vsb200_dense_create": ([C.POINTER(DenseOpts), C.c_int, C.c_int, C.c_int, C.POINTER(vp)], C.c_int),
        "vsb200_dense_push": ([vp, vp, C.c_int, vp, C.c_int, C.c_int64, C.POINTER(C.c_int)], C.c_int),
        "vsb200_dense_push_device": ([vp, vp, C.c_int, C.c_int64, C.POINTER(C.c_int)], C.c_int),
        "vsb200_dense_set_profiling": ([vp, C.c_int], None),
        "vsb200_dense_io_stats": ([vp, C.POINTER(C.c_double)], None),
        "vsb200_dense_flush": ([vp, C.POINTER(C.c_int)], C.c_int),
        "vsb200_dense_pop": ([vp, C.POINTER(FrameResult)], C.c_int),
        "vsb200_dense_last_id_map": ([vp], C.POINTER(C.c_int32)),
        "vsb200_dense_last_proto": ([vp, vp, C.c_size_t], C.c_size_t),
        "vsb200_dense_stats": ([vp, C.POINTER(C.c_double)], None),
        "vsb200_dense_destroy": ([vp], None),
        "vsb200_dense_export_halo": ([vp, vp, vp, C.POINTER(C.c_int32)], C.c_int),
        "vsb200_dense_import_halo": ([vp, vp, vp, C.POINTER(C.c_int32)], C.c_int),
        "vsb200_bgr2lab": ([vp, C.c_int, C.c_int, C.c_int, vp, vp], C.c_int),
        "vsb200_region_hist_scratch_bytes": ([C.c_int, C.c_int, C.c_int], C.c_size_t),
        "vsb200_region_hist_reset": ([vp, C.c_int, C.c_int, C.c_int, vp], C.c_int),
        "vsb200_region_hist_add": ([vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp], C.c_int),
        "vsb200_region_hist_finish": ([vp, C.c_int, C.c_int, C.c_int, vp, vp, vp], C.c_int),
        "vsb200_hist_chisquare": ([vp, C.c_int, vp, C.c_int, vp, vp], C.c_int),
        "vsb200_seg_writer_open": ([C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(vp)], C.c_int),
        "vsb200_seg_writer_add": ([vp, C.c_char_p, C.c_size_t, C.c_int64], C.c_int),
        "vsb200_seg_writer_add_last_frame": ([vp, vp, C.c_int64], C.c_int),
        "vsb200_seg_writer_write_chunk": ([vp], C.c_int),
        "vsb200_seg_writer_close": ([vp], C.c_int),
        "vsb200_seg_reader_open": ([C.c_char_p, C.POINTER(vp)], C.c_int),
        "vsb200_seg_reader_num_frames": ([vp], C.c_int),
        "vsb200_seg_reader_num_header_flags": ([vp], C.c_int),
        "vsb200_seg_reader_header_flags": ([vp], C.POINTER(C.c_int32)),
        "vsb200_seg_reader_time_stamps": ([vp], C.POINTER(C.c_int64)),
        "vsb200_seg_reader_read": ([vp, C.c_int, vp, C.c_size_t], C.c_size_t),
        "vsb200_seg_reader_read_frame": ([vp, C.c_int, vp, C.c_size_t, C.POINTER(C.c_size_t)], C.c_int),
        "vsb200_seg_reader_close": ([vp], None),
        "vsb200_strip_to_essentials": ([vp, C.c_int, vp, C.c_size_t], C.c_size_t),
        "vsb200_encode_frame_proto": ([vp, vp, C.c_size_t], C.c_size_t),
        "vsb200_region_default_opts": ([C.POINTER(RegionOpts)], None),
        "vsb200_region_create": ([C.POINTER(RegionOpts), C.c_int, C.c_int, C.POINTER(vp)], C.c_int),
        "vsb200_region_push": ([vp, C.POINTER(FrameResult), vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)], C.c_int),
        "vsb200_region_flush": ([vp, C.POINTER(C.c_int)], C.c_int),
        "vsb200_region_pop": ([vp, C.POINTER(C.POINTER(C.c_int32))], C.c_longlong),
        "vsb200_region_stats": ([vp, C.POINTER(C.c_double)], None),
        "vsb200_region_destroy": ([vp], None),
        "vsb200_shard_unique_id": ([vp], C.c_int),
        "vsb200_shard_create": ([vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(vp)], C.c_int),
        "vsb200_shard_exchange": ([vp, vp, C.POINTER(C.c_int64), C.POINTER(C.c_int)], C.c_int),
        "vsb200_shard_pred_maps": ([vp], vp),
        "vsb200_shard_vote": ([vp, vp, C.c_int, C.c_int64, vp], C.c_int),
        "vsb200_shard_relabel": ([vp, vp, C.c_size_t, vp, C.c_int], C.c_int),
        "vsb200_shard_stats": ([vp, C.POINTER(C.c_double)], None),
        "vsb200_shard_destroy": ([vp], None),
    }
    missing = []
    for name, (argtypes, restype) in sigs.items():
        try:
            fn = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        if argtypes is not None:
            fn.argtypes = argtypes
        fn.restype = restype
    if missing and not os.environ.get("VSB200_ALLOW_PARTIAL_LIB"):
        raise RuntimeError(f"{LIB_PATH} does not export {missing}: stale build, rebuild it")
    _lib = L
    return L


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().vsb200_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")
