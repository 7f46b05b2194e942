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
