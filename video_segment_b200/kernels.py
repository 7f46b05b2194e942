"""Kernel-level entry points of the C ABI on torch CUDA tensors (torch is only the
device-memory / stream plumbing here).  Mirrors the reference call sites named in
include/vsb200.h; raises if the CUDA library or a B200 is missing."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import check, lib


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def preprocess(bgr_u8: torch.Tensor, presmoothing: int = 2) -> torch.Tensor:
    """DenseSegmentation::PreprocessFeatures (dense_segmentation.cpp:164-198):
    uint8 (H, W, 3) BGR on the GPU -> float32 (H, W, 3)."""
    assert bgr_u8.is_cuda and bgr_u8.dtype == torch.uint8 and bgr_u8.dim() == 3 and bgr_u8.is_contiguous()
    h, w, _ = bgr_u8.shape
    out = torch.empty((h, w, 3), dtype=torch.float32, device=bgr_u8.device)
    scratch = torch.empty(lib().vsb200_preprocess_scratch_bytes(), dtype=torch.uint8, device=bgr_u8.device)
    check(lib().vsb200_preprocess(_ptr(bgr_u8), w * 3, w, h, presmoothing, _ptr(out), _ptr(scratch), _stream()),
          "vsb200_preprocess")
    return out


def edge_build(curr: torch.Tensor, prev: torch.Tensor | None = None, flow: torch.Tensor | None = None,
               l1: bool = False, spatial_out=None, temporal_out=None):
    """AddSpatialEdgesImpl + AddTemporal[Flow]EdgesImpl (dense_segmentation_graph.h:956-1142).
    Returns (spatial (H, W, 4), temporal (H, W, 9) or None); missing edges hold -1."""
    assert curr.is_cuda and curr.dtype == torch.float32 and curr.is_contiguous()
    h, w, _ = curr.shape
    spatial = spatial_out if spatial_out is not None else torch.empty((h, w, 4), dtype=torch.float32, device=curr.device)
    temporal = None
    if prev is not None:
        temporal = temporal_out if temporal_out is not None else torch.empty((h, w, 9), dtype=torch.float32, device=curr.device)
    check(lib().vsb200_edge_build(_ptr(curr), _ptr(prev), _ptr(flow), w, h, int(l1), _ptr(spatial), _ptr(temporal),
                                  _stream()), "vsb200_edge_build")
    return spatial, temporal


def bucket_index(weight: float) -> int:
    return lib().vsb200_bucket_index(float(weight))


def sort_edges(lists, width: int, height: int):
    """Stable bucket sort of the chunk graph's edge lists (segmentation_graph.h:158-162,367-374).
    `lists[q]` = weight tensor of bucket list q or None.  Returns (codes uint32 [E], bucket_start int64 [2049])."""
    dev = next(t for t in lists if t is not None).device
    n = width * height
    total = sum(n * (9 if q & 1 else 4) for q, t in enumerate(lists) if t is not None)
    codes = torch.empty(total, dtype=torch.int32, device=dev)
    bstart = torch.empty(2049, dtype=torch.int64, device=dev)
    sb = lib().vsb200_sort_scratch_bytes(len(lists), width, height)
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    arr = (C.c_void_p * len(lists))(*[(t.data_ptr() if t is not None else None) for t in lists])
    check(lib().vsb200_sort_edges(arr, len(lists), width, height, _ptr(codes), _ptr(bstart), _ptr(scratch), sb,
                                  _stream()), "vsb200_sort_edges")
    return codes, bstart


def segment_chunk(frames: torch.Tensor, min_region_size: int, l1: bool = False):
    """One unconstrained chunk: smoothed frames (T, H, W, 3) float32 -> node labels (T, H, W) int32
    (graph build + sort + FastSegmentationGraph::SegmentGraph + flatten).  Also returns
    [edge ms, sort ms, merge ms, merge rounds]."""
    assert frames.is_cuda and frames.dtype == torch.float32 and frames.is_contiguous()
    t, h, w, _ = frames.shape
    labels = torch.empty((t, h, w), dtype=torch.int32, device=frames.device)
    stats = (C.c_double * 4)()
    check(lib().vsb200_segment_chunk(_ptr(frames), w, h, t, int(l1), int(min_region_size), _ptr(labels), stats,
                                     _stream()), "vsb200_segment_chunk")
    return labels, list(stats)
