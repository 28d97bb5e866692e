"""world_size-2 gloo test of the N>1 host path (tile sharding + all-gather reassembly order), on CPU."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_items, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from instarevive_b200.pipeline import all_gather_items, shard_range
        s, e = shard_range(n_items, rank, world)
        # item i is a (2, 3) block filled with i: every rank contributes only the items it owns
        local = torch.stack([torch.full((2, 3), float(i)) for i in range(s, e)]) if e > s else torch.empty(0, 2, 3)
        full = all_gather_items(local, n_items)
        ok = full.shape == (n_items, 2, 3) and all(float(full[i, 0, 0]) == float(i) for i in range(n_items))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_items", [(2, 9), (2, 2), (3, 25), (2, 1)])
def test_all_gather_items_restores_list_order(world, n_items):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_items, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_sharded_blend_equals_single_rank_blend():
    """Bit-exactness argument of SURVEY hard part 7 on the CPU oracle: shard the tile list, all-gather (emulated by
    concatenation in list order), blend in list order == reference loop."""
    from instarevive_b200.pipeline import shard_range
    from oracle.tiles_oracle import sliding_windows
    g = torch.Generator().manual_seed(0)
    h = w = 128
    wins = sliding_windows(h, w, 64, 56)
    tiles = [torch.randn(1, 4, 64, 64, generator=g) for _ in wins]

    def blend(ts):
        buf = torch.zeros(1, 4, h, w)
        cnt = torch.zeros(1, 4, h, w)
        for t, (hi, he, wi, we) in zip(ts, wins):
            buf[:, :, hi:he, wi:we] += t
            cnt[:, :, hi:he, wi:we] += 1
        return buf / cnt

    ref = blend(tiles)
    for world in (2, 4, 8):
        gathered = []
        for r in range(world):
            s, e = shard_range(len(wins), r, world)
            gathered += tiles[s:e]
        assert torch.equal(blend(gathered), ref)
