"""Multi-process test of the tile-sharded restore on ONE GPU (`pytest -m gpu`): world_size 2 / 4 processes share cuda:0
and talk over gloo (CUDA tensors staged through the host; NCCL refuses several ranks on one device), which exercises the
default two-phase split and the whole three-phase schedule (tile_plan="overlap") of pipeline.restore_latents -- all-gather A, late DiT tile / early decodes, all-gather B,
phase-3 decodes, all-gather C -- with the real kernels. Every rank checks that the sharded image and latents are
bit-identical to its own single-rank restore."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import instarevive_b200 as ir
        from instarevive_b200 import pipeline, weights
        dev = torch.device("cuda:0")
        torch.cuda.set_device(dev)
        net = ir.ControlPixArtMSHalf(ir.PixArtMS(depth=2, input_size=64, micro_condition=True, init_weights=False), 1).eval()
        net.load_state_dict(weights.make_dit_state_dict(depth=2, copy_blocks=1, seed=21), strict=True)
        net = net.to(dev)
        vae = ir.AutoencoderKLDecoder(weights.make_vae_decoder_state_dict(seed=2), device=dev)
        _, _, y, mask, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))
        y, mask = y.to(dev), mask.to(dev)
        H = W = 1024   # 9 tiles of 512 px: 2 ranks -> 4 + 4 + one late tile, 4 ranks -> 2 each + one late tile
        control = torch.from_numpy(weights.synthetic_degraded_image(H, W, seed=5)).to(dev).float().div(255).permute(2, 0, 1)[None]
        init = (weights.SyntheticVAE(None).encode(control * 2 - 1).latent_dist.mode() * 0.18215).contiguous()
        plan = pipeline.TilePlan(pipeline._sliding_windows(128, 128, 64, 56), world)
        img_d, lat_d = pipeline.restore_latents(net, vae, control, init, y, mask, tiled=True, return_latents=True, use_control=True,
                                                tile_plan="overlap")
        img_c = pipeline.restore_latents(net, vae, control, init, y, mask, tiled=True, use_control=True)   # default two-phase split
        img_l, lat_l = pipeline.restore_latents(net, vae, control, init, y, mask, tiled=True, return_latents=True, use_control=True,
                                                distributed=False)
        torch.cuda.synchronize()
        ret[rank] = (bool(plan.three_phase), bool(torch.equal(lat_d, lat_l)), bool(torch.equal(img_d, img_l)),
                     bool(torch.equal(img_c, img_l)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_three_phase_tiled_restore_is_bit_identical_to_single_rank(world):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    port = _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    for r in range(world):
        assert ret.get(r) == (True, True, True, True), (r, ret.get(r))
