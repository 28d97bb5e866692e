"""ORACLE (test infrastructure, not product code): fp32 CPU restatement of the reference VAE decode path (and, for
SURVEY 8f row 1, of the encode path: AutoencoderKL.encode, ldm/models/autoencoder.py:82-86 -> Encoder.forward,
ldm/modules/diffusionmodules/model.py:521-546, Downsample model.py:70-89),
AutoencoderKL.decode (ldm/models/autoencoder.py:88-91) -> Decoder.forward (ldm/modules/diffusionmodules/model.py:622-655)
with the ddconfig of configs/cldm.yaml:69-84 (ch 128, ch_mult (1,2,4,4), 2 res blocks, z 4, no attn_resolutions).

Parity status: PINNED against outputs of the reference Decoder itself (imported unmodified from /root/reference by
oracle/make_goldens.py; fixtures under tests/golden/). Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this module.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def _gn(sd, name, x):
    """Normalize = GroupNorm(32, C, eps=1e-6, affine), model.py:48-49."""
    return F.group_norm(x, 32, sd[f"{name}.weight"], sd[f"{name}.bias"], eps=1e-6)


def _swish(x):
    """nonlinearity, model.py:43-45."""
    return x * torch.sigmoid(x)


def _conv(sd, name, x, padding):
    return F.conv2d(x, sd[f"{name}.weight"], sd[f"{name}.bias"], padding=padding)


def resnet_block(sd, p, x):
    """ResnetBlock.forward with temb None and dropout 0, model.py:131-151."""
    h = _conv(sd, f"{p}.conv1", _swish(_gn(sd, f"{p}.norm1", x)), 1)
    h = _conv(sd, f"{p}.conv2", _swish(_gn(sd, f"{p}.norm2", h)), 1)
    if f"{p}.nin_shortcut.weight" in sd:
        x = _conv(sd, f"{p}.nin_shortcut", x, 0)
    return x + h


def attn_block(sd, p, x):
    """AttnBlock.forward, model.py:181-205: single-head attention over the h*w positions, scale C^-1/2."""
    h_ = _gn(sd, f"{p}.norm", x)
    q, k, v = (_conv(sd, f"{p}.{n}", h_, 0) for n in ("q", "k", "v"))
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.softmax(torch.bmm(q, k) * (int(c) ** -0.5), dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, f"{p}.proj_out", h_, 0)


@torch.no_grad()
def vae_decode(sd, z, num_resolutions: int = 4, num_res_blocks: int = 2):
    """post_quant_conv then Decoder.forward; z: (B, 4, h, w) -> (B, 3, 8h, 8w), fp32."""
    sd = {k: v.float() for k, v in sd.items()}
    d = "decoder"
    z = _conv(sd, "post_quant_conv", z.float(), 0)                       # autoencoder.py:89
    h = _conv(sd, f"{d}.conv_in", z, 1)                                  # model.py:630
    h = resnet_block(sd, f"{d}.mid.block_1", h)                          # model.py:633-635
    h = attn_block(sd, f"{d}.mid.attn_1", h)
    h = resnet_block(sd, f"{d}.mid.block_2", h)
    for i_level in reversed(range(num_resolutions)):                     # model.py:638-644
        for i_block in range(num_res_blocks + 1):
            h = resnet_block(sd, f"{d}.up.{i_level}.block.{i_block}", h)
        if i_level != 0:                                                 # Upsample.forward, model.py:63-67
            h = F.interpolate(h, scale_factor=2.0, mode="nearest")
            h = _conv(sd, f"{d}.up.{i_level}.upsample.conv", h, 1)
    h = _swish(_gn(sd, f"{d}.norm_out", h))                              # model.py:650-652
    return _conv(sd, f"{d}.conv_out", h, 1)


@torch.no_grad()
def vae_encode_moments(sd, x, num_resolutions: int = 4, num_res_blocks: int = 2):
    """Encoder.forward then quant_conv; x: (B, 3, H, W) in [-1, 1] -> moments (B, 8, H/8, W/8), fp32.
    DiagonalGaussianDistribution.mode() (distributions.py:61-62) is moments[:, :4]."""
    sd = {k: v.float() for k, v in sd.items()}
    e = "encoder"
    h = _conv(sd, f"{e}.conv_in", x.float(), 1)                          # model.py:526
    for i_level in range(num_resolutions):                               # model.py:527-535
        for i_block in range(num_res_blocks):
            h = resnet_block(sd, f"{e}.down.{i_level}.block.{i_block}", h)
        if i_level != num_resolutions - 1:                               # Downsample.forward, model.py:82-89
            h = F.pad(h, (0, 1, 0, 1), mode="constant", value=0)
            h = F.conv2d(h, sd[f"{e}.down.{i_level}.downsample.conv.weight"], sd[f"{e}.down.{i_level}.downsample.conv.bias"],
                         stride=2, padding=0)
    h = resnet_block(sd, f"{e}.mid.block_1", h)                          # model.py:538-541
    h = attn_block(sd, f"{e}.mid.attn_1", h)
    h = resnet_block(sd, f"{e}.mid.block_2", h)
    h = _swish(_gn(sd, f"{e}.norm_out", h))                              # model.py:544-546
    h = _conv(sd, f"{e}.conv_out", h, 1)
    return _conv(sd, "quant_conv", h, 0)                                 # autoencoder.py:84
