"""CPU tests of the multi-GPU host logic: group geometry, global id offsets and
the seam exchange (C1 send/recv + C2 all-gather) on world_size 2 and 3 with the gloo backend."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from video_segment_b200.shard import seam_vote, relabel_table, group_range, id_offsets, seam_exchange


def test_group_ranges_cover_the_video_with_one_overlap_frame():
    world, per = 4, 58
    spans = [group_range(r, world, per) for r in range(world)]
    for r in range(1, world):
        prev_start, prev_n = spans[r - 1]
        assert spans[r][0] == prev_start + prev_n - 1          # successor re-reads the predecessor's last frame
    assert spans[0][0] == 0
    with pytest.raises(ValueError):
        group_range(4, 4, 58)


def test_id_offsets_exclusive_prefix():
    assert id_offsets([5, 0, 7]) == [0, 5, 5]
    assert id_offsets([]) == []


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    halo_out = torch.full((2, h, w), rank + 1, dtype=torch.int32)
    halo_out[1] += 100
    halo_in = torch.zeros((2, h, w), dtype=torch.int32)
    got, offs = seam_exchange(halo_out, halo_in, 10 * (rank + 1), rank, world)
    out[rank] = (None if got is None else got.clone().numpy(), offs)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_seam_exchange_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 6, 8, out), nprocs=world, join=True)
    for r in range(world):
        got, offs = out[r]
        assert offs == [sum(10 * (k + 1) for k in range(j)) for j in range(world)]
        if r == 0:
            assert got is None
        else:
            assert np.all(got[0] == r) and np.all(got[1] == r + 100)       # the predecessor's two overlap maps


def test_seam_vote_max_overlap_and_ties():
    pred = torch.tensor([[7, 7, 7, 9], [7, 8, 8, 9], [3, 3, 8, 9]], dtype=torch.int32)
    succ = torch.tensor([[0, 0, 0, 1], [0, 2, 2, 1], [4, 4, 2, -1]], dtype=torch.int32)
    s, p, c = seam_vote(pred, succ)
    assert s.tolist() == [0, 1, 2, 4] and p.tolist() == [7, 9, 8, 3] and c.tolist() == [4, 2, 3, 2]
    # a successor region split evenly between two predecessor regions goes to the smaller id
    s, p, c = seam_vote(torch.tensor([[5, 5, 2, 2]], dtype=torch.int32), torch.tensor([[1, 1, 1, 1]], dtype=torch.int32))
    assert s.tolist() == [1] and p.tolist() == [2] and c.tolist() == [2]
    tab = relabel_table(torch.tensor([0, 1, 2, 4]), torch.tensor([7, 9, 8, 3]), 6, 100)
    assert tab.tolist() == [7, 9, 8, 103, 3, 105]
    # identical partitions under a permutation are recovered exactly
    g = torch.Generator().manual_seed(1)
    a = torch.randint(0, 50, (40, 60), generator=g, dtype=torch.int32)
    perm = torch.randperm(50, generator=g).to(torch.int32)
    s, p, c = seam_vote(a, perm[a.long()])
    assert torch.equal(perm[p.long()].long(), s)
    e = seam_vote(torch.full((2, 2), -1, dtype=torch.int32), torch.zeros((2, 2), dtype=torch.int32))
    assert all(t.numel() == 0 for t in e)
