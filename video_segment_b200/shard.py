"""Multi-GPU sharding of one video over the ranks of a box.

The reference chain is sequential across chunks (dense_segmentation.cpp:281-432), so the unit of
sharding is a contiguous *frame group* per rank: rank g segments frames
[g * L, (g + 1) * L] (one read-overlap frame) with its own chunk chain.  Two exchanges tie the
groups together (both are tiny next to NVLink bandwidth, the point is ordering):
  C1  the region-id maps of the last two overlap frames of group g go to group g + 1
      (overlap_segmentations_, dense_segmentation.cpp:300-315)      -> send / recv
  C2  globally unique region ids = exclusive prefix of the groups' region-id counts
      (max_region_id_, dense_segmentation.cpp:360-365)              -> all-gather
The functions work on any torch.distributed backend (nccl on the GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def group_range(rank: int, world: int, frames_per_group: int) -> Tuple[int, int]:
    """First frame and frame count of rank's group: groups overlap by one frame (the successor
    re-reads the predecessor's last frame, which is where the seam hand-over applies)."""
    if not (0 <= rank < world) or frames_per_group < 2:
        raise ValueError("bad group geometry")
    start = rank * (frames_per_group - 1)
    return start, frames_per_group


def id_offsets(counts: List[int]) -> List[int]:
    """C2: exclusive prefix of the per-group region-id counts."""
    out, acc = [], 0
    for c in counts:
        out.append(acc)
        acc += int(c)
    return out


def seam_exchange(halo_out: torch.Tensor, halo_in: torch.Tensor, max_region_id: int, rank: int, world: int
                  ) -> Tuple[Optional[torch.Tensor], List[int]]:
    """C1 + C2.  halo_out / halo_in: int32 [2, H, W] on the backend's device.  Returns (the
    predecessor's overlap id maps or None on rank 0, id offsets of all groups)."""
    if world == 1:
        return None, [0]
    ops = []
    if rank + 1 < world:
        ops.append(dist.P2POp(dist.isend, halo_out, rank + 1))
    if rank > 0:
        ops.append(dist.P2POp(dist.irecv, halo_in, rank - 1))
    reqs = dist.batch_isend_irecv(ops) if ops else []
    counts = [torch.zeros(1, dtype=torch.int64, device=halo_out.device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([max_region_id], dtype=torch.int64, device=halo_out.device))
    for r in reqs:
        r.wait()
    return (halo_in if rank > 0 else None), id_offsets([int(c.item()) for c in counts])


def seam_vote(pred_map: torch.Tensor, succ_map: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """Parallel seam (second strategy): the successor group segmented the shared frame
    on its own; every region id it uses there is mapped to the predecessor id that covers most of its pixels.
    pred_map / succ_map: int32 [H, W] id maps of the SAME frame (the predecessor's last overlap map and the
    successor's first output frame), on any device.  Returns (succ_ids, pred_ids, overlap_px), succ_ids sorted
    ascending; ties go to the smaller predecessor id.  Runs on the tensors' device (a handful of sort / unique
    passes over H*W keys -- the exchange itself is the point, not this bookkeeping)."""
    if pred_map.shape != succ_map.shape:
        raise ValueError("seam_vote: the two maps must show the same frame")
    s = succ_map.reshape(-1).to(torch.int64)
    p = pred_map.reshape(-1).to(torch.int64)
    ok = (s >= 0) & (p >= 0)
    s, p = s[ok], p[ok]
    if s.numel() == 0:
        e = torch.empty(0, dtype=torch.int64, device=succ_map.device)
        return e, e.clone(), e.clone()
    pairs, counts = torch.unique((s << 32) | p, return_counts=True)          # sorted by (succ, pred)
    ps, pp = pairs >> 32, pairs & 0xFFFFFFFF
    # best predecessor per successor id: sort by (succ asc, count desc, pred asc) and keep the first of each run
    order = torch.argsort(pp, stable=True)
    order = order[torch.argsort(-counts[order], stable=True)]
    order = order[torch.argsort(ps[order], stable=True)]
    ps, pp, counts = ps[order], pp[order], counts[order]
    first = torch.ones_like(ps, dtype=torch.bool)
    first[1:] = ps[1:] != ps[:-1]
    return ps[first], pp[first], counts[first]


def relabel_table(succ_ids: torch.Tensor, pred_ids: torch.Tensor, num_succ_ids: int, id_offset: int) -> torch.Tensor:
    """Successor-local region id -> id in the predecessor's numbering (regions visible in the shared frame)
    or a fresh global id `id_offset + local id` (regions born later).  C2's exclusive prefix is id_offset."""
    table = torch.arange(num_succ_ids, dtype=torch.int64, device=succ_ids.device) + int(id_offset)
    table[succ_ids] = pred_ids
    return table


# ---------------------------------------------------------------------------------------------
# The product's seam hand-over: NCCL in the C++ host layer (csrc/shard.cu, vsb200_shard_*), vote and relabel on the
# device.  torch is not involved; the functions above remain as the backend-agnostic statement of C1 / C2 that the CPU
# tests exercise on gloo.
# ---------------------------------------------------------------------------------------------
import ctypes as _C

import numpy as _np


def nccl_unique_id() -> bytes:
    """Rank 0: the 128-byte NCCL id the other ranks need for SeamLink (send it over any side channel)."""
    from ._lib import check, lib
    buf = (_C.c_uint8 * 128)()
    check(lib().vsb200_shard_unique_id(buf), "vsb200_shard_unique_id")
    return bytes(buf)


class SeamLink:
    """One rank's end of the group seams: vsb200_shard_* over a DenseSegmentationUnit."""

    def __init__(self, unique_id: Optional[bytes], rank: int, world: int, device: int, width: int, height: int):
        from ._lib import check, lib
        self._h = _C.c_void_p()
        self.rank, self.world, self.w, self.h = rank, world, width, height
        idbuf = (_C.c_uint8 * 128)(*unique_id) if unique_id else None
        check(lib().vsb200_shard_create(idbuf, rank, world, device, width, height, _C.byref(self._h)), "vsb200_shard_create")
        self._table = None
        self._keep = None

    def exchange(self, unit) -> List[int]:
        """C1 + C2 at a group boundary of `unit` (right after a chunk boundary).  Returns the exclusive prefix of the
        groups' region-id counts, world + 1 entries (offsets[rank + 1] - offsets[rank] = this group's count)."""
        from ._lib import check, lib
        offs = (_C.c_int64 * (self.world + 1))()
        have = _C.c_int()
        check(lib().vsb200_shard_exchange(self._h, unit._h, offs, _C.byref(have)), "vsb200_shard_exchange")
        self.have_pred = bool(have.value)
        return [int(v) for v in offs]

    def relabel_table(self, own_first_map_dev_ptr: int, n_ids: int, id_offset: int) -> "_np.ndarray":
        """Vote on the device; returns the table (host copy, int32 [n_ids]); a device copy stays for relabel()."""
        import torch
        from ._lib import check, lib
        table = torch.empty(n_ids, dtype=torch.int32, device="cuda")
        check(lib().vsb200_shard_vote(self._h, _C.c_void_p(own_first_map_dev_ptr), n_ids, id_offset, _C.c_void_p(table.data_ptr())),
              "vsb200_shard_vote")
        self._table = table
        return table.cpu().numpy()

    def relabel_device(self, ids_dev_ptr: int, n: int) -> None:
        from ._lib import check, lib
        check(lib().vsb200_shard_relabel(self._h, _C.c_void_p(ids_dev_ptr), n, _C.c_void_p(self._table.data_ptr()), int(self._table.numel())),
              "vsb200_shard_relabel")

    def stats(self) -> dict:
        from ._lib import lib
        a = (_C.c_double * 3)()
        lib().vsb200_shard_stats(self._h, a)
        return dict(exchange_ms=a[0], exchanges=a[1], kernel_launches=a[2])

    def close(self):
        if self._h:
            from ._lib import lib
            lib().vsb200_shard_destroy(self._h)
            self._h = _C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def relabel_results(results: List[dict], table: "_np.ndarray") -> None:
    """Applies a relabel table to the per-region / per-compound id arrays of frame results, in place (ids beyond the
    table are left alone); regions and compounds are re-sorted by id where the reference's constrained output is."""
    n = len(table)

    def m(a):
        a = _np.asarray(a)
        out = a.copy()
        ok = (a >= 0) & (a < n)
        out[ok] = table[a[ok]]
        return out
    for r in results:
        r["region_id"] = m(r["region_id"]).astype(_np.int32)
        if "id_map" in r and r["id_map"] is not None:
            r["id_map"] = m(r["id_map"]).astype(_np.int32)
        if len(r.get("compound", [])):
            r["compound"] = r["compound"].copy()
            r["compound"][:, 0] = m(r["compound"][:, 0])
            r["neighbor_id"] = m(r["neighbor_id"]).astype(_np.int32)
