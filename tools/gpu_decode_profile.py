#!/usr/bin/env python
"""Per-launch view of the VAE decoder at a 128x128 latent (1024^2 image): whole-call time, then (under ncu) every launch.
usage: python tools/gpu_decode_profile.py [latent_side]"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import instarevive_b200 as ir
from instarevive_b200 import weights
dev = torch.device("cuda:0")
side = int(sys.argv[1]) if len(sys.argv) > 1 else 128
vae = ir.AutoencoderKLDecoder(weights.make_vae_decoder_state_dict(seed=2), device=dev)
z = torch.randn(1, 4, side, side, device=dev)
for _ in range(3):
    vae.decode_tensor(z, in_scale=1.0 / 0.18215, out_scale=0.5, out_shift=0.5)
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(10):
    vae.decode_tensor(z, in_scale=1.0 / 0.18215, out_scale=0.5, out_shift=0.5)
b.record()
torch.cuda.synchronize()
print(f"decode {8 * side}^2: {a.elapsed_time(b) / 10:.3f} ms per call")
