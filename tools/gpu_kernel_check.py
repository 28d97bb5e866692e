"""Bring-up check of the individual sm_100a kernels against plain torch references (run on a B200 via gpurun).

Not part of the test-suite (tests/test_gpu_*.py are); this prints max errors and first timings so that kernel
bugs can be localised from one GPU trip.
"""
from __future__ import annotations

import math
import sys
import time
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from instarevive_b200 import _lib  # noqa: E402

# bring-up convenience: tolerate a library that does not yet export every declared symbol
import ctypes as _C  # noqa: E402
_probe = _C.CDLL(str(_lib.LIB_PATH))
_lib.PROTOTYPES = {k: v for k, v in _lib.PROTOTYPES.items() if hasattr(_probe, k)}
L = _lib.lib()
dev = torch.device("cuda:0")
P = _lib.ptr
S = _lib.stream_ptr
ok_all = True


def report(name, got, ref, tol):
    global ok_all
    err = (got.float() - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    good = err <= tol * max(scale, 1.0) and math.isfinite(err)
    ok_all &= good
    print(f"[{'ok' if good else 'FAIL'}] {name}: max_abs_err={err:.4g} ref_max={scale:.4g} tol={tol * max(scale, 1.0):.4g}", flush=True)
    return good


def time_ms(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def gemm_case(M, N, K, epi, bn, batch=1, shared_a=False):
    g = torch.Generator(device="cpu").manual_seed(M * 7 + N * 3 + K + epi + bn)
    ab = 1 if shared_a else batch
    A = (torch.randn(ab, M, K, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(batch, N, K, generator=g) * 0.05).to(dev).bfloat16()
    bias = torch.randn(N, generator=g).to(dev)
    ref = torch.einsum("bmk,bnk->bmn", A.float().expand(batch, M, K), W.float()) + bias
    strideA = 0 if (shared_a or batch == 1) else M * K
    if epi == 2:
        T = 128 if M % 128 == 0 else M
        nb = M // T
        gate = torch.randn(nb, N, generator=g).to(dev)
        resid = torch.randn(batch, M, N, generator=g).to(dev)
        out = torch.empty(batch, M, N, device=dev)
        outb = torch.empty(batch, M, N, device=dev, dtype=torch.bfloat16)
        ref = resid + gate.repeat_interleave(T, 0)[None] * ref
        _lib.check(L.ir_gemm_bf16(P(A), P(W), P(bias), M, N, K, batch, strideA, N * K, M * N, 2, 1.0, P(outb), P(out),
                                  P(resid), P(gate), N, T, bn, S()), "gemm")
        torch.cuda.synchronize()
        report(f"gemm f32 M{M} N{N} K{K} bn{bn} b{batch}", out, ref, 2e-3)
        report(f"gemm f32 bf16-copy M{M} N{N} K{K} bn{bn}", outb, ref, 1e-2)
    else:
        if epi == 1:
            ref = F.gelu(ref, approximate="tanh")
        out = torch.empty(batch, M, N, device=dev, dtype=torch.bfloat16)
        _lib.check(L.ir_gemm_bf16(P(A), P(W), P(bias), M, N, K, batch, strideA, N * K, M * N, epi, 1.0, P(out), None,
                                  None, None, 0, 1, bn, S()), "gemm")
        torch.cuda.synchronize()
        report(f"gemm bf16 epi{epi} M{M} N{N} K{K} bn{bn} b{batch} sharedA={shared_a}", out, ref, 1e-2)


def gemm_perf(M, N, K, bn):
    A = torch.randn(M, K, device=dev).bfloat16()
    W = torch.randn(N, K, device=dev).bfloat16()
    bias = torch.randn(N, device=dev)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    ms = time_ms(lambda: L.ir_gemm_bf16(P(A), P(W), P(bias), M, N, K, 1, 0, 0, 0, 0, 1.0, P(out), None, None, None, 0, 1,
                                        bn, S()))
    ms_t = time_ms(lambda: torch.matmul(A, W.t()))
    print(f"[perf] gemm M{M} N{N} K{K} bn{bn}: {ms * 1e3:.1f} us  {2.0 * M * N * K / ms / 1e9:.1f} TFLOP/s "
          f"(torch.matmul {ms_t * 1e3:.1f} us {2.0 * M * N * K / ms_t / 1e9:.1f} TFLOP/s)", flush=True)


def conv_case(n, H, W, C, Cout, bn, f32=False):
    g = torch.Generator(device="cpu").manual_seed(H + W + C + Cout)
    x = (torch.randn(n, C, H, W, generator=g)).to(dev)
    w = (torch.randn(Cout, C, 3, 3, generator=g) * 0.03).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    xb = x.bfloat16()
    wb = w.bfloat16()
    ref = F.conv2d(xb.float(), wb.float(), b, padding=1)
    act = xb.permute(0, 2, 3, 1).contiguous()
    wk = wb.permute(0, 2, 3, 1).contiguous().view(Cout, 9 * C)  # (Cout, ky, kx, C) -> tap-major K
    if f32:
        resid = torch.randn(n, H, W, Cout, generator=g).to(dev)
        out = torch.empty(n, H, W, Cout, device=dev)
        _lib.check(L.ir_conv3x3_bf16(P(act), P(wk), P(b), n, H, W, C, Cout, None, P(out), None, P(resid), bn, S()), "conv")
        torch.cuda.synchronize()
        report(f"conv3x3 f32 n{n} {H}x{W} C{C}->{Cout} bn{bn}", out.permute(0, 3, 1, 2), ref + resid.permute(0, 3, 1, 2), 2e-3)
    else:
        out = torch.empty(n, H, W, Cout, device=dev, dtype=torch.bfloat16)
        _lib.check(L.ir_conv3x3_bf16(P(act), P(wk), P(b), n, H, W, C, Cout, P(out), None, None, None, bn, S()), "conv")
        torch.cuda.synchronize()
        report(f"conv3x3 bf16 n{n} {H}x{W} C{C}->{Cout} bn{bn}", out.permute(0, 3, 1, 2), ref, 1e-2)
    return act, wk, b, out


def conv_perf(n, H, W, C, Cout, bn):
    act = torch.randn(n, H, W, C, device=dev).bfloat16()
    wk = (torch.randn(Cout, 9 * C, device=dev) * 0.03).bfloat16()
    b = torch.randn(Cout, device=dev)
    out = torch.empty(n, H, W, Cout, device=dev, dtype=torch.bfloat16)
    ms = time_ms(lambda: L.ir_conv3x3_bf16(P(act), P(wk), P(b), n, H, W, C, Cout, P(out), None, None, None, bn, S()), 10)
    fl = 2.0 * n * H * W * Cout * 9 * C
    print(f"[perf] conv3x3 n{n} {H}x{W} C{C}->{Cout} bn{bn}: {ms * 1e3:.1f} us {fl / ms / 1e9:.1f} TFLOP/s", flush=True)


def attn_case(B, T, heads=16, hd=72, cross_lens=None):
    g = torch.Generator(device="cpu").manual_seed(B * 11 + T)
    D = heads * hd
    if cross_lens is None:
        qkv = torch.randn(B * T, 3 * D, generator=g).to(dev).bfloat16()
        out = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
        _lib.check(L.ir_attention_bf16(P(qkv), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, P(out), 3 * D, 3 * D, 3 * D,
                                       D, B, heads, hd, T, T, None, None, hd ** -0.5, S()), "attention")
        torch.cuda.synchronize()
        q, k, v = qkv.float().view(B, T, 3, heads, hd).permute(2, 0, 3, 1, 4)
        ref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(B * T, D)
        report(f"self-attn B{B} T{T}", out, ref, 1e-2)
        ms = time_ms(lambda: L.ir_attention_bf16(P(qkv), qkv.data_ptr() + 2 * D, qkv.data_ptr() + 4 * D, P(out), 3 * D,
                                                 3 * D, 3 * D, D, B, heads, hd, T, T, None, None, hd ** -0.5, S()))
        print(f"[perf] self-attn B{B} T{T}: {ms * 1e3:.1f} us {4.0 * B * heads * T * T * hd / ms / 1e9:.1f} TFLOP/s", flush=True)
    else:
        sumL = sum(cross_lens)
        qm = torch.randn(B * T, D, generator=g).to(dev).bfloat16()
        kv = torch.randn(sumL, 2 * D, generator=g).to(dev).bfloat16()
        off = torch.tensor([sum(cross_lens[:i]) for i in range(B)], dtype=torch.int32, device=dev)
        ln = torch.tensor(cross_lens, dtype=torch.int32, device=dev)
        out = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
        _lib.check(L.ir_attention_bf16(P(qm), P(kv), kv.data_ptr() + 2 * D, P(out), D, 2 * D, 2 * D, D, B, heads, hd, T, 0,
                                       P(off), P(ln), hd ** -0.5, S()), "attention")
        torch.cuda.synchronize()
        refs = []
        for b in range(B):
            q = qm[b * T:(b + 1) * T].float().view(T, heads, hd).permute(1, 0, 2)
            kk = kv[off[b]:off[b] + ln[b]].float().view(-1, 2, heads, hd)
            k, v = kk[:, 0].permute(1, 0, 2), kk[:, 1].permute(1, 0, 2)
            refs.append(F.scaled_dot_product_attention(q, k, v).permute(1, 0, 2).reshape(T, D))
        report(f"cross-attn B{B} T{T} lens{cross_lens}", out, torch.cat(refs), 1e-2)


def attn_tc_case(B, T, heads=16, hd=72):
    """EPI_QKV GEMM (head-major scatter) + tcgen05 attention vs torch SDPA on the same bf16 q/k/v."""
    g = torch.Generator(device="cpu").manual_seed(B * 13 + T)
    D = heads * hd
    M = B * T
    Tp = (T + 7) // 8 * 8
    A = (torch.randn(M, D, generator=g) * 0.5).to(dev).bfloat16()
    W = (torch.randn(3 * D, D, generator=g) * 0.03).to(dev).bfloat16()
    bias = (torch.randn(3 * D, generator=g) * 0.1).to(dev)
    qh = torch.zeros(B, heads, T, hd, device=dev, dtype=torch.bfloat16)
    kh = torch.zeros_like(qh)
    vt = torch.zeros(B, heads, hd, Tp, device=dev, dtype=torch.bfloat16)
    _lib.check(L.ir_gemm_qkv_heads(P(A), P(W), P(bias), M, D, T, Tp, heads, hd, P(qh), P(kh), P(vt), 0, S()), "qkv heads")
    torch.cuda.synchronize()
    ref = (A.float() @ W.float().t() + bias).view(B, T, 3, heads, hd)
    report(f"qkv-heads q B{B} T{T}", qh, ref[:, :, 0].permute(0, 2, 1, 3), 1e-2)
    report(f"qkv-heads k B{B} T{T}", kh, ref[:, :, 1].permute(0, 2, 1, 3), 1e-2)
    report(f"qkv-heads vt B{B} T{T}", vt[..., :T], ref[:, :, 2].permute(0, 2, 3, 1), 1e-2)
    out = torch.zeros(M, D, device=dev, dtype=torch.bfloat16)
    _lib.check(L.ir_attention_tc_bf16(P(qh), P(kh), P(vt), P(out), D, B, heads, hd, T, Tp, hd ** -0.5, S()), "attention_tc")
    torch.cuda.synchronize()
    q, k, v = qh.float(), kh.float(), vt[..., :T].float().transpose(2, 3)
    sref = F.scaled_dot_product_attention(q, k, v).permute(0, 2, 1, 3).reshape(M, D)
    report(f"attn-tc B{B} T{T}", out, sref, 1e-2)
    # d truncated to 64: what the result would be if the 16-column remainder k-step were ignored (bisecting aid)
    s64 = F.scaled_dot_product_attention(q[..., :64], k[..., :64], v, scale=hd ** -0.5).permute(0, 2, 1, 3).reshape(M, D)
    print(f"      (vs d-truncated-to-64 reference: {(out.float() - s64).abs().max().item():.4g})", flush=True)
    ms = time_ms(lambda: L.ir_attention_tc_bf16(P(qh), P(kh), P(vt), P(out), D, B, heads, hd, T, Tp, hd ** -0.5, S()))
    print(f"[perf] attn-tc B{B} T{T}: {ms * 1e3:.1f} us {4.0 * B * heads * T * T * hd / ms / 1e9:.1f} TFLOP/s", flush=True)
    ms = time_ms(lambda: L.ir_gemm_qkv_heads(P(A), P(W), P(bias), M, D, T, Tp, heads, hd, P(qh), P(kh), P(vt), 0, S()))
    print(f"[perf] qkv-heads gemm M{M}: {ms * 1e3:.1f} us {2.0 * M * 3 * D * D / ms / 1e9:.1f} TFLOP/s", flush=True)


def ln_case(B, T, D=1152):
    g = torch.Generator(device="cpu").manual_seed(5)
    x = (torch.randn(B * T, D, generator=g) * 3 + 1).to(dev)
    shift = torch.randn(B, 6, D, generator=g).to(dev)
    out = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
    _lib.check(L.ir_ln_modulate(P(x), P(out), P(shift), shift.data_ptr() + 4 * D, 6 * D, B * T, T, D, S()), "ln")
    torch.cuda.synchronize()
    ref = F.layer_norm(x, (D,), eps=1e-6).view(B, T, D) * (1 + shift[:, 1:2]) + shift[:, 0:1]
    report(f"ln_modulate B{B} T{T}", out, ref.view(B * T, D), 1e-2)
    ms = time_ms(lambda: L.ir_ln_modulate(P(x), P(out), P(shift), shift.data_ptr() + 4 * D, 6 * D, B * T, T, D, S()))
    print(f"[perf] ln_modulate rows {B * T}: {ms * 1e3:.1f} us {6.0 * B * T * D / ms / 1e6:.1f} GB/s", flush=True)


def main():
    print(torch.cuda.get_device_name(0), L.ir_version().decode(), flush=True)
    which = sys.argv[1:] or ["gemm", "conv", "attn", "attn_tc", "ln", "perf"]
    if "gemm" in which:
        for bn in (64, 128, 256, 2128, 2256):
            gemm_case(256, 256, 128, 0, bn)
        for bn in (64, 128, 256, 2128, 2256):
            gemm_case(1024, 1152, 1152, 0, bn)
        for bn in (2128, 2256):
            gemm_case(1000, 3456, 1152, 0, bn)
            gemm_case(1152, 1152, 4608, 2, bn)
            gemm_case(384, 4608, 1152, 1, bn)
            gemm_case(120, 2304, 1152, 0, bn, batch=5, shared_a=True)
            gemm_case(640, 512, 512, 2, bn, batch=3)
        gemm_case(1000, 3456, 1152, 0, 0)
        gemm_case(1024, 4608, 1152, 1, 0)
        gemm_case(1024, 1152, 4608, 2, 0)
        gemm_case(120, 2304, 1152, 0, 0, batch=5, shared_a=True)
        gemm_case(512, 512, 512, 2, 128, batch=3)
        gemm_case(77, 1152, 4096, 1, 0)
    if "conv" in which:
        for bn in (64, 128, 256, 2128, 2256):
            conv_case(1, 64, 64, 128, 256, bn)
        conv_case(2, 24, 40, 64, 128, 2128)
        conv_case(1, 40, 72, 128, 128, 2128)
        conv_case(1, 64, 64, 512, 512, 2256, f32=True)
        conv_case(2, 24, 40, 64, 128, 0)
        conv_case(1, 64, 64, 512, 512, 0, f32=True)
    if "attn" in which:
        attn_case(1, 1024)
        attn_case(2, 1000)
        attn_case(1, 4096)
        attn_case(2, 1024, cross_lens=[120, 77])
        attn_case(3, 600, cross_lens=[1, 300, 64])
    if "attn_tc" in which:
        attn_tc_case(1, 256)
        attn_tc_case(1, 1024)
        attn_tc_case(2, 1000)
        attn_tc_case(1, 4096)
        attn_tc_case(3, 1296)
    if "ln" in which:
        ln_case(2, 1024)
    if "perf" in which:
        for (M, N, K) in [(1024, 1152, 1152), (1024, 3456, 1152), (1024, 4608, 1152), (1024, 1152, 4608),
                          (4096, 1152, 1152), (4096, 4608, 1152), (4096, 1152, 4608), (8192, 8192, 8192)]:
            for bn in (64, 128, 256, 2128, 2256):
                gemm_perf(M, N, K, bn)
        for (H, C, Co) in [(64, 512, 512), (128, 512, 512), (256, 256, 256), (512, 128, 128), (1024, 128, 128)]:
            for bn in (128, 256, 2128, 2256):
                if bn % 1000 <= Co:
                    conv_perf(1, H, H, C, Co, bn)
    if "calib" in which:
        # tile-configuration calibration sweep for pick_cfg (gemm.cu): the shapes of the 1024^2 and tiled-2048^2 workloads
        for M in (4096, 9216, 25600):
            for (N, K) in [(1152, 1152), (3456, 1152), (4608, 1152), (1152, 4608)]:
                for bn in (128, 256, 2128, 2256):
                    gemm_perf(M, N, K, bn)
        for (n, H, C, Co) in [(1, 128, 512, 512), (1, 256, 512, 512), (1, 256, 512, 256), (1, 256, 256, 256),
                              (1, 512, 256, 256), (1, 512, 256, 128), (1, 512, 128, 128), (1, 1024, 128, 128),
                              (4, 64, 512, 512), (4, 128, 512, 512), (4, 256, 256, 256), (4, 512, 128, 128)]:
            for bn in (128, 2128, 2256):
                if bn % 1000 <= Co:
                    conv_perf(n, H, H, C, Co, bn)
    print("ALL OK" if ok_all else "SOME FAILED", flush=True)
    return 0 if ok_all else 1


if __name__ == "__main__":
    sys.exit(main())
