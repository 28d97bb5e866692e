"""Bisect batch-dependence: per-tile DiT output for a 25-tile batch vs 13/12-tile batches, and VAE decode."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import instarevive_b200 as ir
from instarevive_b200 import pipeline, weights

dev = torch.device("cuda:0")
depth, cb = int(sys.argv[1]) if len(sys.argv) > 1 else 2, int(sys.argv[2]) if len(sys.argv) > 2 else 1
net = ir.ControlPixArtMSHalf(ir.PixArtMS(depth=depth, input_size=64, micro_condition=True, init_weights=False), cb).eval()
net.load_state_dict(weights.make_dit_state_dict(depth=depth, copy_blocks=cb, seed=21), strict=True)
net = net.to(dev)
vae = ir.AutoencoderKLDecoder(weights.make_vae_decoder_state_dict(seed=2), device=dev)
_, _, y, mask, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))
y, mask = y.to(dev), mask.to(dev)
H = W = 2048
control = torch.from_numpy(weights.synthetic_degraded_image(H, W, seed=0)).to(dev).float().div(255).permute(2, 0, 1)[None].contiguous()
init = (weights.SyntheticVAE(None).encode(control * 2 - 1).latent_dist.mode() * 0.18215).contiguous()
windows = pipeline._sliding_windows(256, 256, 64, 56)
coords = torch.tensor([(c[0], c[2]) for c in windows], dtype=torch.int32, device=dev)
sched = ir.DDPMSchedulerLite()
tin = pipeline.tile_gather(init, coords, 64, 64, 1).view(-1, 4, 64, 64)
full = ir.generate_sample_1step(net, sched, tin, 400, y, mask, use_control=True)
again = ir.generate_sample_1step(net, sched, tin, 400, y, mask, use_control=True)
print("run-to-run identical (batch 25):", torch.equal(full, again))
a = ir.generate_sample_1step(net, sched, tin[:13].contiguous(), 400, y, mask, use_control=True)
b = ir.generate_sample_1step(net, sched, tin[13:].contiguous(), 400, y, mask, use_control=True)
parts = torch.cat([a, b])
print("25 vs 13+12 identical:", torch.equal(full, parts), "max diff", (full - parts).abs().max().item())
for i in range(25):
    if not torch.equal(full[i], parts[i]):
        print("  tile", i, "differs, max", (full[i] - parts[i]).abs().max().item())
for nb in (1, 5, 8, 12):
    c = ir.generate_sample_1step(net, sched, tin[:nb].contiguous(), 400, y, mask, use_control=True)
    print(f"batch {nb} prefix identical to batch-25 prefix:", torch.equal(c, full[:nb]))
# VAE decode batch dependence on real latents
lat = pipeline.tile_blend(full.view(25, 1, 4, 64, 64), coords, 256, 256, 1)
zt = pipeline.tile_gather(lat, coords, 64, 64, 1).view(-1, 4, 64, 64)
d8 = torch.cat([vae.decode_tensor(zt[i:i + 8].contiguous()) for i in range(0, 25, 8)])
d5 = torch.cat([vae.decode_tensor(zt[i:i + 5].contiguous()) for i in range(0, 25, 5)])
d1 = torch.cat([vae.decode_tensor(zt[i:i + 1].contiguous()) for i in range(0, 25, 1)])
print("decode 8 vs 5:", torch.equal(d8, d5), "8 vs 1:", torch.equal(d8, d1), (d8 - d1).abs().max().item())
d8b = torch.cat([vae.decode_tensor(zt[i:i + 8].contiguous()) for i in range(0, 25, 8)])
print("decode run-to-run:", torch.equal(d8, d8b))
