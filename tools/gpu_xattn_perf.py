#!/usr/bin/env python
"""Isolated timing of the cross-attention kernels (CUDA events, 50 launches after 5 warm-ups): tcgen05 (xattention_tc.cu)
vs the mma.sync flash_attn_kernel on the DiT shapes. usage: python tools/gpu_xattn_perf.py"""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from instarevive_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
heads, hd = 16, 72
D = heads * hd
for (B, T, lens) in ((1, 4096, [77]), (1, 1024, [77]), (25, 1024, [77] * 25), (8, 4096, [120] * 8), (4, 1024, [300, 120, 77, 5])):
    qm = torch.randn(B * T, D, device=dev).bfloat16()
    kv = torch.randn(sum(lens), 2 * D, device=dev).bfloat16()
    off = torch.tensor([sum(lens[:i]) for i in range(B)], dtype=torch.int32, device=dev)
    ln = torch.tensor(lens, dtype=torch.int32, device=dev)
    win = max((sum(lens[:i]) % 8) + lens[i] for i in range(B))   # key window: the TMA box starts at a multiple of 8 rows
    out = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
    vt = torch.empty(L.ir_cross_attention_vt_bytes(heads, sum(lens)), dtype=torch.uint8, device=dev)
    s = _lib.stream_ptr()

    def tc():
        _lib.check(L.ir_cross_attention_tc_bf16(qm.data_ptr(), kv.data_ptr(), vt.data_ptr(), out.data_ptr(), D, 2 * D, D, B, heads, hd, T, sum(lens),
                                                off.data_ptr(), ln.data_ptr(), win, hd ** -0.5, s))

    def legacy():
        _lib.check(L.ir_attention_bf16(qm.data_ptr(), kv.data_ptr(), kv.data_ptr() + 2 * D, out.data_ptr(), D, 2 * D, 2 * D, D, B, heads, hd,
                                       T, 0, off.data_ptr(), ln.data_ptr(), hd ** -0.5, s))

    res = {}
    for name, fn in (("tcgen05", tc), ("mma.sync", legacy)):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(50):
            fn()
        b.record()
        torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 50 * 1e3
    gb = 2 * B * T * D * 2 / 1e9
    print(f"B{B} T{T} L{max(lens)}: tcgen05 {res['tcgen05']:.1f} us ({gb / res['tcgen05'] * 1e6:.0f} GB/s q+out), mma.sync {res['mma.sync']:.1f} us")
