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


def bgr2lab(bgr: torch.Tensor) -> torch.Tensor:
    """cv::cvtColor(CV_BGR2Lab) on 8-bit data (region_descriptor.cpp:73): (H, W, 3) uint8 -> (H, W, 3) uint8."""
    assert bgr.is_cuda and bgr.dtype == torch.uint8 and bgr.is_contiguous()
    h, w, _ = bgr.shape
    out = torch.empty_like(bgr)
    check(lib().vsb200_bgr2lab(_ptr(bgr), w * 3, w, h, _ptr(out), _stream()), "vsb200_bgr2lab")
    return out


def region_hist(bgr_frames, id_maps, n_regions: int, lum_bins: int = 10, color_bins: int = 20):
    """AppearanceDescriptor3D of every region over the frames of a chunk set (region_descriptor.cpp:97-111,
    histograms.cpp:140-211,340-360).  bgr_frames: (H, W, 3) uint8 tensors, id_maps: (H, W) int32 region ids.
    Returns (normalised histograms float32 [n_regions, lum * col * col], weight sums float32 [n_regions])."""
    dev = bgr_frames[0].device
    total = lum_bins * color_bins * color_bins
    sb = lib().vsb200_region_hist_scratch_bytes(n_regions, lum_bins, color_bins)
    if sb == 0:
        raise ValueError("region_hist: bad histogram geometry")
    scratch = torch.empty(sb, dtype=torch.uint8, device=dev)
    check(lib().vsb200_region_hist_reset(_ptr(scratch), n_regions, lum_bins, color_bins, _stream()), "vsb200_region_hist_reset")
    for bgr, ids in zip(bgr_frames, id_maps):
        assert bgr.is_cuda and bgr.dtype == torch.uint8 and bgr.is_contiguous()
        assert ids.is_cuda and ids.dtype == torch.int32 and ids.is_contiguous()
        h, w, _ = bgr.shape
        check(lib().vsb200_region_hist_add(_ptr(bgr), w * 3, _ptr(ids), w, h, n_regions, lum_bins, color_bins, _ptr(scratch),
                                           _stream()), "vsb200_region_hist_add")
    hist = torch.empty((n_regions, total), dtype=torch.float32, device=dev)
    wsum = torch.empty(n_regions, dtype=torch.float32, device=dev)
    check(lib().vsb200_region_hist_finish(_ptr(scratch), n_regions, lum_bins, color_bins, _ptr(hist), _ptr(wsum), _stream()),
          "vsb200_region_hist_finish")
    return hist, wsum


def hist_chisquare(hist: torch.Tensor, pairs: torch.Tensor) -> torch.Tensor:
    """ColorHistogram::ChiSquareDist (histograms.cpp:391-407) for region pairs (n, 2) int32."""
    assert hist.is_cuda and hist.dtype == torch.float32 and hist.is_contiguous()
    assert pairs.is_cuda and pairs.dtype == torch.int32 and pairs.is_contiguous()
    out = torch.empty(pairs.shape[0], dtype=torch.float32, device=hist.device)
    check(lib().vsb200_hist_chisquare(_ptr(hist), hist.shape[1], _ptr(pairs), pairs.shape[0], _ptr(out), _stream()),
          "vsb200_hist_chisquare")
    return out


def label_components(labels: torch.Tensor):
    """K11 + K10 (csrc/shape.cu): N4 connected components of every label in every frame of an int32 label volume
    (S, H, W) on the GPU and their shape moments -- ConnectedComponents(raster, N4_CONNECT) and
    ShapeMomentsFromRasterization (segment_util/segmentation_util.cpp:1007-1101, 652-693) for all regions at once.
    Returns (component map (S, H, W) int32 on the GPU, records): components are numbered in the order of their first
    scan interval; records is a dict of numpy arrays first / count / label / slice / area (int32) and moments
    (float32 [n, 5]: mean_x, mean_y, xx, xy, yy)."""
    assert labels.is_cuda and labels.dtype == torch.int32 and labels.dim() == 3 and labels.is_contiguous()
    s, h, w = labels.shape
    comp = torch.full_like(labels, -1)
    cap = s * h * w
    rec = np.zeros((cap, 10), np.int32)
    n = C.c_int(0)
    check(lib().vsb200_label_components(_ptr(labels), w, h, s, _ptr(comp), rec.ctypes.data_as(C.c_void_p), cap, C.byref(n), _stream()),
          "vsb200_label_components")
    rec = rec[:n.value]
    return comp, dict(first=rec[:, 0], count=rec[:, 1], label=rec[:, 2], slice=rec[:, 3], area=rec[:, 4],
                      moments=np.ascontiguousarray(rec[:, 5:10]).view(np.float32))
