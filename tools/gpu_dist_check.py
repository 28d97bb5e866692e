"""2-GPU diagnostic (torchrun): tile-sharded restore vs the same restore done locally on every rank."""
import os, sys
from pathlib import Path
import torch, torch.distributed as dist
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import instarevive_b200 as ir
from instarevive_b200 import pipeline, weights

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
net = ir.ControlPixArtMSHalf(ir.PixArtMS(depth=2, input_size=64, micro_condition=True, init_weights=False), 1).eval()
net.load_state_dict(weights.make_dit_state_dict(depth=2, copy_blocks=1, seed=21), strict=True)
net = net.to(dev)
vae = ir.AutoencoderKLDecoder(weights.make_vae_decoder_state_dict(seed=2), device=dev)
_, _, y, mask, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))
y, mask = y.to(dev), mask.to(dev)
H = W = 2048
control = torch.from_numpy(weights.synthetic_degraded_image(H, W, seed=0)).to(dev).float().div(255).permute(2, 0, 1)[None].contiguous()
init = (weights.SyntheticVAE(None).encode(control * 2 - 1).latent_dist.mode() * 0.18215).contiguous()
# are the inputs identical on both ranks?
g = [torch.empty_like(init) for _ in range(dist.get_world_size())]
dist.all_gather(g, init)
print(f"rank {rank}: init identical across ranks: {all(torch.equal(g[0], t) for t in g)} max diff {max((g[0]-t).abs().max().item() for t in g)}", flush=True)
img_d, lat_d = pipeline.restore_latents(net, vae, control, init, y, mask, tiled=True, return_latents=True, use_control=True)
img_l, lat_l = pipeline.restore_latents(net, vae, control, init, y, mask, tiled=True, return_latents=True, distributed=False, use_control=True)
torch.cuda.synchronize()
print(f"rank {rank}: latents identical {torch.equal(lat_d, lat_l)} ({(lat_d-lat_l).abs().max().item():.3g}); "
      f"pixels identical {torch.equal(img_d, img_l)} ({(img_d-img_l).abs().max().item():.3g})", flush=True)
dist.destroy_process_group()
