#!/usr/bin/env python
"""VAE decode / encode whole-call time with CUDA-graph replay on and off, at several sizes (batch 1)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import instarevive_b200 as ir
from instarevive_b200 import weights
dev = torch.device("cuda:0")
vae = ir.AutoencoderKL(weights.make_vae_state_dict(dec_seed=2, enc_seed=5), device=dev)


def timed(fn, n=20):
    for _ in range(4):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for side in (32, 64, 128):
    z = torch.randn(1, 4, side, side, device=dev)
    x = torch.rand(1, 3, 8 * side, 8 * side, device=dev) * 2 - 1
    for rep in range(2):
        for on in (True, False):
            vae.set_cuda_graphs(on)
            d = timed(lambda: vae.decode_tensor(z, in_scale=1.0 / 0.18215, out_scale=0.5, out_shift=0.5))
            e = timed(lambda: vae.encode_moments(x))
            print(f"{8 * side:5d}^2 graphs {'on ' if on else 'off'}: decode {d:7.3f} ms  encode {e:7.3f} ms", flush=True)
