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
n = L.ir_debug_gemm_trace(buf.data_ptr(), 1)
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

# ---- implicit-GEMM convs of the VAE decoder (full-resolution N = 128 layers, a 256-channel layer, an upsample phase conv)
for (n, H, W, C, Cout, label) in ((1, 1024, 1024, 128, 128, "conv 1024^2 C128->128 (K=1152)"), (1, 1024, 1024, 256, 128, "conv 1024^2 C256->128 (K=2304)"),
                                  (1, 512, 512, 256, 256, "conv 512^2 C256->256 (K=2304)"), (1, 256, 256, 512, 512, "conv 256^2 C512->512 (K=4608)")):
    act = torch.randn(n, H, W, C, device=dev).bfloat16()
    wt = (torch.randn(Cout, 3, 3, C, device=dev) * 0.02).bfloat16()
    bias = torch.randn(Cout, device=dev)
    out = torch.empty(n, H, W, Cout, device=dev, dtype=torch.bfloat16)
    def run():
        _lib.check(L.ir_conv3x3_bf16(P(act), P(wt), P(bias), n, H, W, C, Cout, P(out), None, None, None, 0, S()))
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        run()
    e1.record()
    torch.cuda.synchronize()
    t = buf.cpu().tolist()
    base = t[0]
    us = e0.elapsed_time(e1) / 10 * 1e3
    print(f"\n{label}: {us:.1f} us per launch ({2.0 * n * H * W * Cout * 9 * C / us / 1e6:.0f} TFLOP/s); CTA 0 of the last launch (us from entry):")
    print("   " + "  ".join(f"{NAMES[i]} {(t[i] - base) / 1e3:.2f}" for i in range(16) if t[i] >= base and t[i] > 0))


# ---- the decoder's convs in situ (fused GroupNorm statistics, residual adds): one record per launch of a 1024^2 decode
import instarevive_b200 as ir
from instarevive_b200 import weights
vae = ir.AutoencoderKLDecoder(weights.make_vae_decoder_state_dict(seed=2), device=torch.device(dev))
z = torch.randn(1, 4, 128, 128, device=dev)
for _ in range(2):
    vae.decode_tensor(z)
torch.cuda.synchronize()
ring = torch.zeros(128, 16, dtype=torch.int64, device=dev)
L.ir_debug_gemm_trace(ring.data_ptr(), 128)
L.ir_profile_begin()
vae.decode_tensor(z)
torch.cuda.synchronize()
import ctypes as C
nrec = int(L.ir_profile_records(None, None, None, None, None, 0))
kl, Mv, Nv, Kv, ms = (C.c_int * nrec)(), (C.c_int * nrec)(), (C.c_int * nrec)(), (C.c_int * nrec)(), (C.c_float * nrec)()
L.ir_profile_records(kl, Mv, Nv, Kv, ms, nrec)
L.ir_profile_end((C.c_double * 8)(), (C.c_double * 8)(), (C.c_longlong * 8)())
L.ir_debug_gemm_trace(None, 1)
rows = ring.cpu().tolist()
print("\nVAE decode 1024^2, per GEMM / conv launch: event us | CTA 0: start->first landed, mainloop tile0, tile1, epilogue tile0, last epilogue exposed (us)")
for i in range(min(nrec, 128)):
    t = rows[i]
    if t[0] == 0:
        continue
    f = lambda a, b: (t[a] - t[b]) / 1e3 if t[a] > 0 and t[b] > 0 else float("nan")
    print(f"  #{i:3d} class {kl[i]} M{Mv[i]:8d} N{Nv[i]:4d} K{Kv[i]:5d}: {ms[i] * 1e3:7.1f} | fill {f(5, 2):5.2f} main0 {f(6, 5):6.2f} main1 {f(7, 6):6.2f} epi0 {f(12, 10):6.2f} tail {f(13, 9 if t[9] else (8 if t[8] else (7 if t[7] else 6))):6.2f} total(cta0) {f(15, 2):7.1f}")
