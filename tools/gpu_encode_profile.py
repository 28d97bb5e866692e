#!/usr/bin/env python
"""Per-launch view of AutoencoderKL.encode at 1024^2 (ir_profile_records: class, shape, CUDA-event time per tensor-core launch)
next to the whole-call time; usage: python tools/gpu_encode_profile.py"""
import ctypes as C
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import instarevive_b200 as ir
from instarevive_b200 import _lib, weights
L = _lib.lib()
dev = torch.device("cuda:0")
vae = ir.AutoencoderKL(weights.make_vae_state_dict(dec_seed=2, enc_seed=5), device=dev)
x = torch.rand(1, 3, 1024, 1024, device=dev) * 2 - 1
for _ in range(3):
    vae.encode_moments(x)
torch.cuda.synchronize()
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record()
for _ in range(10):
    vae.encode_moments(x)
b.record()
torch.cuda.synchronize()
print(f"encode 1024^2: {a.elapsed_time(b) / 10:.3f} ms per call")
L.ir_profile_begin()
vae.encode_moments(x)
torch.cuda.synchronize()
n = int(L.ir_profile_records(None, None, None, None, None, 0))
kl, M, N, K, ms = (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)(), (C.c_int * n)(), (C.c_float * n)()
L.ir_profile_records(kl, M, N, K, ms, n)
L.ir_profile_end((C.c_double * 8)(), (C.c_double * 8)(), (C.c_longlong * 8)())
tot = 0.0
for i in range(n):
    fl = 2.0 * M[i] * N[i] * K[i]
    tot += ms[i]
    print(f"  #{i:2d} class {kl[i]} M{M[i]:8d} N{N[i]:4d} K{K[i]:5d}: {ms[i] * 1e3:7.1f} us {fl / (ms[i] * 1e-3) / 1e12 if ms[i] > 0 else 0:6.0f} TFLOP/s")
print(f"tensor-core launches: {tot:.3f} ms of the call")
