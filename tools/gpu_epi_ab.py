"""Isolated timing of the DiT linear GEMMs with their own epilogues (heuristic tile configuration), for A/B runs of the
epilogue variants of a debug build: IR_GEMM_DIRECT=0 python tools/gpu_epi_ab.py  vs  IR_GEMM_DIRECT=15 ...
Back-to-back launches (warm L2, boost clocks): compare variants against each other, not against in-step numbers."""
import sys; sys.path.insert(0, '.')
import torch
from instarevive_b200 import _lib
L = _lib.lib(); P = _lib.ptr; S = _lib.stream_ptr; dev = 'cuda'


def t(fn, it=30):
    _lib.check(fn(), 'gemm')
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / it * 1e3


D, H, hd = 1152, 16, 72
# clock / allocator warm-up: the first timed case of a fresh process otherwise reads 2-4x high
_w = torch.randn(4096, 4096, device=dev).bfloat16()
for _ in range(200): _w @ _w
torch.cuda.synchronize(); del _w
for M, T in ((4096, 4096), (1024, 1024), (25600, 1024)):
    Tp = (T + 7) // 8 * 8
    A = torch.randn(M, D, device=dev).bfloat16(); A4 = torch.randn(M, 4 * D, device=dev).bfloat16()
    W = (torch.randn(4 * D, 4 * D, device=dev) * 0.02).bfloat16(); b = torch.randn(4 * D, device=dev)
    o = torch.empty(M, 4 * D, device=dev, dtype=torch.bfloat16); x = torch.randn(M, D, device=dev); gate = torch.randn(M // T, 6 * D, device=dev)
    q = torch.empty(M * D, device=dev, dtype=torch.bfloat16); k = torch.empty_like(q); vt = torch.empty((M // T) * D * Tp, device=dev, dtype=torch.bfloat16)
    gp = gate.data_ptr() + 2 * D * 4
    cases = {
        "qkv   (N3456 K1152, scatter)": (lambda: L.ir_gemm_qkv_heads(P(A), P(W), P(b), M, D, T, Tp, H, hd, P(q), P(k), P(vt), 0, S()), 3 * D, D),
        "proj  (N1152 K1152, f32+gate+copy)": (lambda: L.ir_gemm_bf16(P(A), P(W), P(b), M, D, D, 1, 0, 0, 0, 2, 1.0, P(o), P(x), P(x), gp, 6 * D, T, 0, S()), D, D),
        "q_lin (N1152 K1152, bf16)": (lambda: L.ir_gemm_bf16(P(A), P(W), P(b), M, D, D, 1, 0, 0, 0, 0, 1.0, P(o), None, None, None, 0, 1, 0, S()), D, D),
        "xproj (N1152 K1152, f32 in place)": (lambda: L.ir_gemm_bf16(P(A), P(W), P(b), M, D, D, 1, 0, 0, 0, 2, 1.0, None, P(x), P(x), None, 0, 1, 0, S()), D, D),
        "fc1   (N4608 K1152, GELU)": (lambda: L.ir_gemm_bf16(P(A), P(W), P(b), M, 4 * D, D, 1, 0, 0, 0, 1, 1.0, P(o), None, None, None, 0, 1, 0, S()), 4 * D, D),
        "fc2   (N1152 K4608, f32+gate)": (lambda: L.ir_gemm_bf16(P(A4), P(W), P(b), M, D, 4 * D, 1, 0, 0, 0, 2, 1.0, None, P(x), P(x), gp, 6 * D, T, 0, S()), D, 4 * D),
    }
    tot = 0.0
    for name, (fn, N, K) in cases.items():
        us = t(fn); tot += us
        print(f"M{M:6d} {name:36s} {us:8.1f} us {2.0 * M * N * K / us / 1e6:7.0f} TF")
    print(f"M{M:6d} one block's linear GEMMs: {tot:8.1f} us")
