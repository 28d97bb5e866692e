"""ORACLE (test infrastructure, not product code): fp32 CPU restatement of the reference's stage-1 SwinIR forward
(diffusion/model/swinir.py:845-905 `forward` / `forward_features`, RSTB :430-493, SwinTransformerBlock :175-290,
WindowAttention :76-156, window_partition / window_reverse :44-73) for configs/swinir.yaml (embed 180, 8 x 6 blocks, 6 heads,
window 8, mlp_ratio 2, PixelUnshuffle(8), 'nearest+conv' upsampler with upscale 8, '1conv' residual connection).

Parity status: PINNED against outputs of the reference SwinIR class itself (imported unmodified from /root/reference by
oracle/make_goldens_swinir.py; fixtures tests/golden/swinir_*.npz). Only tests/ may import this module."""
from __future__ import annotations

import torch
import torch.nn.functional as F

RGB_MEAN = (0.4488, 0.4371, 0.4040)   # swinir.py:692-693


def relative_position_index(ws: int) -> torch.Tensor:
    """swinir.py:103-114."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_mask(H: int, W: int, ws: int, shift: int) -> torch.Tensor:
    """calculate_mask, swinir.py:227-248: (nW, ws*ws, ws*ws) of 0 / -100."""
    img = torch.zeros(1, H, W, 1)
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = img.view(1, H // ws, ws, W // ws, ws, 1).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws)
    m = mw.unsqueeze(1) - mw.unsqueeze(2)
    return m.masked_fill(m != 0, -100.0).masked_fill(m == 0, 0.0)


def _block(sd, p, x, H, W, ws, heads, shift):
    """SwinTransformerBlock.forward, swinir.py:250-290 (drop_path is the identity at inference)."""
    B, L, C = x.shape
    h = F.layer_norm(x, (C,), sd[f"{p}.norm1.weight"], sd[f"{p}.norm1.bias"]).view(B, H, W, C)
    if shift > 0:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    win = h.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)   # window_partition
    N = ws * ws
    qkv = F.linear(win, sd[f"{p}.attn.qkv.weight"], sd[f"{p}.attn.qkv.bias"]).view(-1, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (C // heads) ** -0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    bias = sd[f"{p}.attn.relative_position_bias_table"][relative_position_index(ws).view(-1)].view(N, N, -1).permute(2, 0, 1)
    attn = attn + bias.unsqueeze(0)
    if shift > 0:
        m = shift_mask(H, W, ws, shift)
        nW = m.shape[0]
        attn = (attn.view(-1, nW, heads, N, N) + m.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = attn.softmax(dim=-1)
    o = (attn @ v).transpose(1, 2).reshape(-1, N, C)
    o = F.linear(o, sd[f"{p}.attn.proj.weight"], sd[f"{p}.attn.proj.bias"])
    o = o.view(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, C)            # window_reverse
    if shift > 0:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    x = x + o.view(B, L, C)
    h2 = F.layer_norm(x, (C,), sd[f"{p}.norm2.weight"], sd[f"{p}.norm2.bias"])
    h2 = F.linear(F.gelu(F.linear(h2, sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"])), sd[f"{p}.mlp.fc2.weight"],
                  sd[f"{p}.mlp.fc2.bias"])                                                             # Mlp, nn.GELU (erf)
    return x + h2


@torch.no_grad()
def swinir_forward(sd, x, depths=(6,) * 8, heads: int = 6, ws: int = 8, sf: int = 8):
    """x: (B, 3, H, W) in [0, 1], H and W multiples of sf*ws -> (B, 3, H, W)."""
    sd = {k: v.float() for k, v in sd.items()}
    Hin, Win = x.shape[2:]
    mean = torch.tensor(RGB_MEAN).view(1, 3, 1, 1)
    x = x.float() - mean                                                                  # swinir.py:871-872 (img_range 1)
    x = F.conv2d(F.pixel_unshuffle(x, sf), sd["conv_first.1.weight"], sd["conv_first.1.bias"], padding=1)   # :883
    B, C, H, W = x.shape
    first = x
    t = x.flatten(2).transpose(1, 2)                                                      # PatchEmbed :535-539
    t = F.layer_norm(t, (C,), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"])
    for li, depth in enumerate(depths):                                                   # RSTB :492-493
        r = t
        for bi in range(depth):                                                           # BasicLayer :408-416
            r = _block(sd, f"layers.{li}.residual_group.blocks.{bi}", r, H, W, ws, heads, 0 if bi % 2 == 0 else ws // 2)
        r = r.transpose(1, 2).view(B, C, H, W)
        r = F.conv2d(r, sd[f"layers.{li}.conv.weight"], sd[f"layers.{li}.conv.bias"], padding=1)
        t = r.flatten(2).transpose(1, 2) + t
    t = F.layer_norm(t, (C,), sd["norm.weight"], sd["norm.bias"])                       # :864-865
    x = t.transpose(1, 2).view(B, C, H, W)
    x = F.conv2d(x, sd["conv_after_body.weight"], sd["conv_after_body.bias"], padding=1) + first           # :884
    x = F.leaky_relu(F.conv2d(x, sd["conv_before_upsample.0.weight"], sd["conv_before_upsample.0.bias"], padding=1), 0.01)
    for n in ("conv_up1", "conv_up2", "conv_up3"):                                       # :886-891
        x = F.leaky_relu(F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), sd[f"{n}.weight"], sd[f"{n}.bias"],
                                  padding=1), 0.2)
    x = F.leaky_relu(F.conv2d(x, sd["conv_hr.weight"], sd["conv_hr.bias"], padding=1), 0.2)
    x = F.conv2d(x, sd["conv_last.weight"], sd["conv_last.bias"], padding=1)             # :892
    x = x + mean                                                                          # :899
    return x[:, :, :Hin, :Win]
