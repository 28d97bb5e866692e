#!/usr/bin/env python
"""One profiled launch of a DiT GEMM shape for `ncu --set full --import-source on --profile-from-start off`:
usage: python tools/gpu_gemm_ncu_case.py {fc1|proj|qlin|fc2|qkv}"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from instarevive_b200 import _lib
L = _lib.lib(); P = _lib.ptr; S = _lib.stream_ptr; dev = "cuda"
CASES = {"fc1": (4096, 4608, 1152, 1), "proj": (4096, 1152, 1152, 2), "qlin": (4096, 1152, 1152, 0), "fc2": (4096, 1152, 4608, 2),
         "qkv": (4096, 3456, 1152, 0)}
M, N, K, epi = CASES[sys.argv[1] if len(sys.argv) > 1 else "fc1"]
T = 4096
A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); b = torch.randn(N, device=dev)
o = torch.empty(M, N, device=dev, dtype=torch.bfloat16); x = torch.randn(M, N, device=dev); gate = torch.randn(M // T, 6 * N, device=dev)
def run():
    _lib.check(L.ir_gemm_bf16(P(A), P(W), P(b), M, N, K, 1, 0, 0, 0, epi, 1.0, P(o), P(x) if epi == 2 else None, P(x) if epi == 2 else None,
                              gate.data_ptr() + 2 * N * 4 if epi == 2 else None, 6 * N, T, 0, S()))
for _ in range(3):
    run()
torch.cuda.synchronize()
torch.cuda.profiler.start()
run()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
