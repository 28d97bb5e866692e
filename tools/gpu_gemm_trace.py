#!/usr/bin/env python
"""Where does a small GEMM launch spend its time? Needs a library built with IR_DEBUG=1 (python -m instarevive_b200.csrc.build
--force with IR_DEBUG=1 in the environment). Runs back-to-back launches of one DiT GEMM shape and prints, for CTA 0 of the last
launch, %globaltimer deltas between its role stamps, next to the event-timed launch-to-launch period."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from instarevive_b200 import _lib
L = _lib.lib(); P = _lib.ptr; S = _lib.stream_ptr; dev = "cuda"
buf = torch.zeros(16, dtype=torch.int64, device=dev)
n = L.ir_debug_gemm_trace(buf.data_ptr())
print("trace slots:", n)
NAMES = ["entry", "prologue done", "pdl_wait done", "first TMA issued", "producer done", "first operands landed", "tile0 issued",
         "tile1 issued", "tile2 issued", "tile3+ issued", "acc tile0 complete", "acc last complete", "epi tile0 done", "epi last done",
         "roles joined", "exit"]
for (M, N, K, epi, cfg, label) in ((4096, 1152, 1152, 2, 0, "proj (EPI_F32 resid+gate+bf16 copy)"), (4096, 1152, 1152, 0, 0, "q_linear (EPI_BF16)"),
                                   (4096, 4608, 1152, 1, 0, "fc1 (GELU)"), (4096, 1152, 4608, 2, 0, "fc2"), (4096, 3456, 1152, 0, 0, "qkv-like (EPI_BF16)"),
                                   (4096, 1152, 1152, 2, 1128, "proj forced 1x128"), (4096, 1152, 1152, 2, 1064, "proj forced 1x64")):
    T = 4096
    A = torch.randn(M, K, device=dev).bfloat16(); W = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); b = torch.randn(N, device=dev)
    o = torch.empty(M, N, device=dev, dtype=torch.bfloat16); x = torch.randn(M, N, device=dev); gate = torch.randn(M // T, 6 * N, device=dev)
    def run():
        _lib.check(L.ir_gemm_bf16(P(A), P(W), P(b), M, N, K, 1, 0, 0, 0, epi, 1.0, P(o), P(x) if epi == 2 else None, P(x) if epi == 2 else None,
                                  gate.data_ptr() + 2 * N * 4 if epi == 2 else None, 6 * N, T, cfg, S()))
    for _ in range(5):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(20):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    base = t[0]
    print(f"\n{label}: M{M} N{N} K{K}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch (back to back); CTA 0 of the last launch (us from entry):")
    print("   " + "  ".join(f"{NAMES[i]} {(t[i] - base) / 1e3:.2f}" for i in range(16) if t[i] >= base and t[i] > 0))
