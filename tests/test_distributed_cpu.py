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


def _plan_worker(rank, world, port, hw, ret):
    """The communication pattern of the three-phase tiled restore (pipeline.TilePlan) with stand-in payloads: tile t's
    "latent" is t + 0.5, its "decoded tile" 1000 + t. Checks on every rank: all-gather A gives list order, all-gather B
    delivers the late latents, early tiles never look at a late tile, all-gather C restores list order."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from instarevive_b200.pipeline import TilePlan, _all_gather_equal, _sliding_windows, gather_decoded_tiles
        windows = _sliding_windows(hw, hw, 64, 56)
        nt = len(windows)
        plan = TilePlan(windows, world)
        ok = plan.three_phase
        first, late = plan.dit_tiles(rank)
        lat = torch.tensor([t + 0.5 for t in first]).view(-1, 1)
        a = _all_gather_equal(lat).view(-1, 1)
        n_a = plan.base * world
        ok &= a.flatten().tolist() == [t + 0.5 for t in range(n_a)]
        early, rest = plan.decode_tiles(rank)
        if early is not None:   # an early tile only needs tiles that all-gather A delivered
            need = [t for t in range(nt) if windows[t][0] < windows[early][1] and windows[early][0] < windows[t][1]
                    and windows[t][2] < windows[early][3] and windows[early][2] < windows[t][3]]
            ok &= all(t < n_a for t in need)
        late_lat = torch.tensor([[late + 0.5]]) if late is not None else torch.zeros(1, 1)
        b = _all_gather_equal(late_lat)[: plan.rem, 0]
        full = torch.cat([a, b]).flatten().tolist()
        ok &= full == [t + 0.5 for t in range(nt)]
        mine = ([early] if early is not None else []) + rest
        px = torch.tensor([1000.0 + t for t in mine]).view(-1, 1, 1)
        out = gather_decoded_tiles(plan, px)
        ok &= out.flatten().tolist() == [1000.0 + t for t in range(nt)]
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,hw", [(2, 256), (3, 256), (2, 128), (4, 128)])
def test_three_phase_tile_plan_communication(world, hw):
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_plan_worker, args=(world, port, hw, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_tile_plan_covers_every_tile_once_and_bounds_the_critical_path():
    from instarevive_b200.pipeline import TilePlan, _sliding_windows
    for hw in (128, 184, 256):
        windows = _sliding_windows(hw, hw, 64, 56)
        nt = len(windows)
        for world in (1, 2, 3, 4, 5, 8, 16, 32):
            plan = TilePlan(windows, world)
            dit, dec = [], []
            for r in range(world):
                first, late = plan.dit_tiles(r)
                early, rest = plan.decode_tiles(r)
                dit += first + ([late] if late is not None else [])
                dec += ([early] if early is not None else []) + rest
                if plan.three_phase:   # critical path: base DiT tiles + one mixed phase + the even phase-3 split
                    assert len(first) == plan.base and len(rest) <= -(-(nt - len(plan.early)) // world)
                    assert not (late is not None and early is not None)
            assert sorted(dit) == list(range(nt)) and sorted(dec) == list(range(nt)), (hw, world)
    plan = TilePlan(_sliding_windows(256, 256, 64, 56), 8)   # BASELINE configs[3]: 25 tiles on 8 GPUs
    assert plan.three_phase and plan.late == [24] and len(plan.early) == 7
    assert max(len(plan.decode_tiles(r)[1]) for r in range(8)) == 3   # 3 d + max(d, v) + 3 v instead of 4 d + 4 v
