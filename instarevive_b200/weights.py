"""Deterministic, seeded weight factories with the reference's state_dict keys and shapes.

There is no network (no released checkpoint), so parity is measured on random-init weights. The reference's own
initialisation (PixArtMS.initialize, diffusion/model/nets/PixArtMS.py:250-285; ControlT2IDitBlockHalf.__init__,
pixart_controlnet.py:30-36) leaves every bias, every cross_attn.proj, final_layer.linear, before_proj and after_proj
at ZERO, which would make the output identically zero and the control branch inert. The factory therefore follows
the reference's distributions where they are non-degenerate (Xavier-uniform Linear weights, N(0, 0.02) embedder MLPs,
randn/sqrt(D) adaLN tables) and draws every tensor the reference zero-initialises from N(0, 0.02).

The same function runs in the build container (to load the UNMODIFIED reference modules and mint tests/golden/) and
on the GPU box (to load the CUDA path), so both sides see bit-identical fp32 weights: generation is on the CPU with a
torch.Generator and a fixed parameter order.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import torch


def _xavier(gen, out_f, in_f):
    bound = math.sqrt(6.0 / (in_f + out_f))
    return (torch.rand(out_f, in_f, generator=gen) * 2 - 1) * bound


def _normal(gen, *shape, std=0.02):
    return torch.randn(*shape, generator=gen) * std


def _block(sd, gen, prefix, D, Dm):
    sd[f"{prefix}.scale_shift_table"] = torch.randn(6, D, generator=gen) / D ** 0.5
    sd[f"{prefix}.attn.qkv.weight"] = _xavier(gen, 3 * D, D)
    sd[f"{prefix}.attn.qkv.bias"] = _normal(gen, 3 * D)
    sd[f"{prefix}.attn.proj.weight"] = _xavier(gen, D, D)
    sd[f"{prefix}.attn.proj.bias"] = _normal(gen, D)
    sd[f"{prefix}.cross_attn.q_linear.weight"] = _xavier(gen, D, D)
    sd[f"{prefix}.cross_attn.q_linear.bias"] = _normal(gen, D)
    sd[f"{prefix}.cross_attn.kv_linear.weight"] = _xavier(gen, 2 * D, D)
    sd[f"{prefix}.cross_attn.kv_linear.bias"] = _normal(gen, 2 * D)
    sd[f"{prefix}.cross_attn.proj.weight"] = _normal(gen, D, D)  # zero in the reference init
    sd[f"{prefix}.cross_attn.proj.bias"] = _normal(gen, D)
    sd[f"{prefix}.mlp.fc1.weight"] = _xavier(gen, Dm, D)
    sd[f"{prefix}.mlp.fc1.bias"] = _normal(gen, Dm)
    sd[f"{prefix}.mlp.fc2.weight"] = _xavier(gen, D, Dm)
    sd[f"{prefix}.mlp.fc2.bias"] = _normal(gen, D)


def make_dit_state_dict(depth: int = 28, copy_blocks: int = 13, seed: int = 1, hidden: int = 1152,
                        caption_channels: int = 4096, model_max_length: int = 120, in_channels: int = 4,
                        mlp_ratio: int = 4, input_size: int = 64) -> "OrderedDict[str, torch.Tensor]":
    """fp32 state_dict of ControlPixArtMSHalf(PixArtMS(depth, hidden, patch 2, 16 heads, micro_condition=True)).

    Keys follow the weight contract of SURVEY 8b (668 tensors for depth 28 / 13 copied blocks, including the two
    buffers pos_embed and y_embedder.y_embedding that forward() never reads).
    """
    gen = torch.Generator(device="cpu").manual_seed(seed)
    D, Dm, Dz = hidden, hidden * mlp_ratio, hidden // 3
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    bm = "base_model"
    n_patches = (input_size // 2) ** 2
    sd[f"{bm}.pos_embed"] = torch.zeros(1, n_patches, D)  # buffer; unused by ControlPixArtMSHalf.forward
    sd[f"{bm}.x_embedder.proj.weight"] = _xavier(gen, D, in_channels * 4).view(D, in_channels, 2, 2)
    sd[f"{bm}.x_embedder.proj.bias"] = _normal(gen, D)
    sd[f"{bm}.t_embedder.mlp.0.weight"] = _normal(gen, D, 256)
    sd[f"{bm}.t_embedder.mlp.0.bias"] = _normal(gen, D)
    sd[f"{bm}.t_embedder.mlp.2.weight"] = _normal(gen, D, D)
    sd[f"{bm}.t_embedder.mlp.2.bias"] = _normal(gen, D)
    sd[f"{bm}.t_block.1.weight"] = _normal(gen, 6 * D, D)
    sd[f"{bm}.t_block.1.bias"] = _normal(gen, 6 * D)
    sd[f"{bm}.y_embedder.y_embedding"] = torch.randn(model_max_length, caption_channels, generator=gen) / caption_channels ** 0.5
    sd[f"{bm}.y_embedder.y_proj.fc1.weight"] = _normal(gen, D, caption_channels)
    sd[f"{bm}.y_embedder.y_proj.fc1.bias"] = _normal(gen, D)
    sd[f"{bm}.y_embedder.y_proj.fc2.weight"] = _normal(gen, D, D)
    sd[f"{bm}.y_embedder.y_proj.fc2.bias"] = _normal(gen, D)
    for i in range(depth):
        _block(sd, gen, f"{bm}.blocks.{i}", D, Dm)
    sd[f"{bm}.final_layer.scale_shift_table"] = torch.randn(2, D, generator=gen) / D ** 0.5
    sd[f"{bm}.final_layer.linear.weight"] = _normal(gen, 4 * 2 * in_channels, D)
    sd[f"{bm}.final_layer.linear.bias"] = _normal(gen, 4 * 2 * in_channels)
    for name in ("csize_embedder", "ar_embedder"):
        sd[f"{bm}.{name}.mlp.0.weight"] = _normal(gen, Dz, 256)
        sd[f"{bm}.{name}.mlp.0.bias"] = _normal(gen, Dz)
        sd[f"{bm}.{name}.mlp.2.weight"] = _normal(gen, Dz, Dz)
        sd[f"{bm}.{name}.mlp.2.bias"] = _normal(gen, Dz)
    for j in range(copy_blocks):
        _block(sd, gen, f"controlnet.{j}.copied_block", D, Dm)
        if j == 0:
            sd["controlnet.0.before_proj.weight"] = _normal(gen, D, D)
            sd["controlnet.0.before_proj.bias"] = _normal(gen, D)
        sd[f"controlnet.{j}.after_proj.weight"] = _normal(gen, D, D)
        sd[f"controlnet.{j}.after_proj.bias"] = _normal(gen, D)
    return sd


# ---------------------------------------------------------------------------------------------- VAE decoder
def _conv(sd, gen, name, cout, cin, k, gain=1.0):
    fan_in = cin * k * k
    bound = gain * math.sqrt(3.0 / fan_in)  # unit-gain uniform: keeps activation scale through the 30-layer stack
    sd[f"{name}.weight"] = (torch.rand(cout, cin, k, k, generator=gen) * 2 - 1) * bound
    sd[f"{name}.bias"] = _normal(gen, cout, std=0.05)


def _norm(sd, gen, name, c):
    sd[f"{name}.weight"] = 1.0 + _normal(gen, c, std=0.1)
    sd[f"{name}.bias"] = _normal(gen, c, std=0.1)


def _resblock(sd, gen, name, cin, cout):
    _norm(sd, gen, f"{name}.norm1", cin)
    _conv(sd, gen, f"{name}.conv1", cout, cin, 3)
    _norm(sd, gen, f"{name}.norm2", cout)
    _conv(sd, gen, f"{name}.conv2", cout, cout, 3, gain=0.5)
    if cin != cout:
        _conv(sd, gen, f"{name}.nin_shortcut", cout, cin, 1)


def make_vae_decoder_state_dict(seed: int = 2, ch: int = 128, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                                z_channels: int = 4, out_ch: int = 3) -> "OrderedDict[str, torch.Tensor]":
    """fp32 weights of `post_quant_conv` + ldm Decoder (ldm/modules/diffusionmodules/model.py:549-655) with the
    ddconfig of configs/cldm.yaml:69-84 (ch 128, ch_mult (1,2,4,4), 2 res blocks, z 4, no attn_resolutions).
    Keys: post_quant_conv.*, decoder.* (SURVEY 8b "VAE surface")."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    _conv(sd, gen, "post_quant_conv", z_channels, z_channels, 1)
    d = "decoder"
    block_in = ch * ch_mult[-1]
    _conv(sd, gen, f"{d}.conv_in", block_in, z_channels, 3)
    _resblock(sd, gen, f"{d}.mid.block_1", block_in, block_in)
    _norm(sd, gen, f"{d}.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(sd, gen, f"{d}.mid.attn_1.{n}", block_in, block_in, 1, gain=1.0 if n != "proj_out" else 0.5)
    _resblock(sd, gen, f"{d}.mid.block_2", block_in, block_in)
    for i_level in reversed(range(len(ch_mult))):
        block_out = ch * ch_mult[i_level]
        for i_block in range(num_res_blocks + 1):
            _resblock(sd, gen, f"{d}.up.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != 0:
            _conv(sd, gen, f"{d}.up.{i_level}.upsample.conv", block_in, block_in, 3)
    _norm(sd, gen, f"{d}.norm_out", block_in)
    _conv(sd, gen, f"{d}.conv_out", out_ch, block_in, 3, gain=0.35)
    return sd


def make_vae_encoder_state_dict(seed: int = 5, ch: int = 128, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2,
                                z_channels: int = 4, in_channels: int = 3) -> "OrderedDict[str, torch.Tensor]":
    """fp32 weights of ldm Encoder (ldm/modules/diffusionmodules/model.py:440-546) + `quant_conv`
    (ldm/models/autoencoder.py:84) with the ddconfig of configs/cldm.yaml:69-84 (double_z, no attn_resolutions).
    Keys: encoder.*, quant_conv.* (SURVEY 8f row 1)."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    e = "encoder"
    _conv(sd, gen, f"{e}.conv_in", ch, in_channels, 3)
    block_in = ch
    for i_level in range(len(ch_mult)):
        block_out = ch * ch_mult[i_level]
        for i_block in range(num_res_blocks):
            _resblock(sd, gen, f"{e}.down.{i_level}.block.{i_block}", block_in, block_out)
            block_in = block_out
        if i_level != len(ch_mult) - 1:
            _conv(sd, gen, f"{e}.down.{i_level}.downsample.conv", block_in, block_in, 3)
    _resblock(sd, gen, f"{e}.mid.block_1", block_in, block_in)
    _norm(sd, gen, f"{e}.mid.attn_1.norm", block_in)
    for n in ("q", "k", "v", "proj_out"):
        _conv(sd, gen, f"{e}.mid.attn_1.{n}", block_in, block_in, 1, gain=1.0 if n != "proj_out" else 0.5)
    _resblock(sd, gen, f"{e}.mid.block_2", block_in, block_in)
    _norm(sd, gen, f"{e}.norm_out", block_in)
    _conv(sd, gen, f"{e}.conv_out", 2 * z_channels, block_in, 3)
    _conv(sd, gen, "quant_conv", 2 * z_channels, 2 * z_channels, 1)
    return sd


def make_vae_state_dict(dec_seed: int = 2, enc_seed: int = 5) -> "OrderedDict[str, torch.Tensor]":
    """Decoder + encoder weights under the reference's AutoencoderKL key names."""
    sd = make_vae_decoder_state_dict(seed=dec_seed)
    sd.update(make_vae_encoder_state_dict(seed=enc_seed))
    return sd


# ---------------------------------------------------------------------------------------------- SwinIR stage 1
def make_swinir_state_dict(seed: int = 7, embed_dim: int = 180, depths=(6,) * 8, num_heads: int = 6, window: int = 8,
                           mlp_ratio: int = 2, in_chans: int = 3, unshuffle: int = 8, num_feat: int = 64
                           ) -> "OrderedDict[str, torch.Tensor]":
    """fp32 parameters of the stage-1 SwinIR (diffusion/model/swinir.py:629-825 with configs/swinir.yaml: embed 180,
    8 RSTB x 6 blocks, 6 heads, window 8, mlp_ratio 2, pixel-unshuffle 8, 'nearest+conv' upsampler, '1conv'). Keys are the
    reference's parameter names; the buffers (relative_position_index, attn_mask) are functions of the window size.
    Scales keep the 48-block residual stream O(1) so that every branch matters in the parity tests."""
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    C, hid = embed_dim, embed_dim * mlp_ratio

    def lin(name, cout, cin, gain=1.0):
        sd[f"{name}.weight"] = torch.randn(cout, cin, generator=gen) * (gain / math.sqrt(cin))
        sd[f"{name}.bias"] = _normal(gen, cout, std=0.02)

    _conv(sd, gen, "conv_first.1", C, in_chans * unshuffle * unshuffle, 3)
    _norm(sd, gen, "patch_embed.norm", C)
    for li, depth in enumerate(depths):
        for bi in range(depth):
            p = f"layers.{li}.residual_group.blocks.{bi}"
            _norm(sd, gen, f"{p}.norm1", C)
            sd[f"{p}.attn.relative_position_bias_table"] = torch.randn((2 * window - 1) ** 2, num_heads, generator=gen) * 0.5
            lin(f"{p}.attn.qkv", 3 * C, C)
            lin(f"{p}.attn.proj", C, C, gain=0.4)
            _norm(sd, gen, f"{p}.norm2", C)
            lin(f"{p}.mlp.fc1", hid, C)
            lin(f"{p}.mlp.fc2", C, hid, gain=0.4)
        _conv(sd, gen, f"layers.{li}.conv", C, C, 3, gain=0.4)
    _norm(sd, gen, "norm", C)
    _conv(sd, gen, "conv_after_body", C, C, 3, gain=0.5)
    _conv(sd, gen, "conv_before_upsample.0", num_feat, C, 3)
    for n in ("conv_up1", "conv_up2", "conv_up3", "conv_hr"):
        _conv(sd, gen, n, num_feat, num_feat, 3, gain=1.3)
    _conv(sd, gen, "conv_last", in_chans, num_feat, 3, gain=0.5)
    return sd


# ---------------------------------------------------------------------------------------------- synthetic inputs
def make_inputs(B: int, h: int, w: int, seed: int = 0, lmax: int = 120, lens=(77,), caption_channels: int = 4096):
    """Synthetic DiT inputs of SURVEY 8d: x = c = unit-scale latents, 120-token caption with `lens` valid tokens."""
    gen = torch.Generator(device="cpu").manual_seed(1000 + seed)
    x = torch.randn(B, 4, h, w, generator=gen)
    y = torch.randn(B, 1, lmax, caption_channels, generator=gen)
    mask = torch.zeros(B, 1, 1, lmax, dtype=torch.int64)
    for b in range(B):
        mask[b, :, :, : lens[b % len(lens)]] = 1
    timestep = torch.full((B,), 400.0)
    data_info = {
        "img_hw": torch.tensor([[float(h * 8), float(w * 8)]] * B),
        "aspect_ratio": torch.tensor([[float(h) / float(w)]] * B),
    }
    return x, timestep, y, mask, data_info


# ---------------------------------------------------------------------------------------------- synthetic images / encoder
def synthetic_degraded_image(height: int, width: int, seed: int = 0):
    """uint8 RGB "degraded image" of SURVEY 8d: uniform noise, 3x box blur, sigma=10 gaussian noise, clamp.
    Pure single-threaded numpy in float64 with a fixed summation order, so the image is bit-identical on every host
    and for every OpenMP thread count (torchrun sets OMP_NUM_THREADS=1)."""
    import numpy as np
    rs = np.random.RandomState(seed)
    img = rs.randint(0, 256, size=(height, width, 3)).astype(np.float64)
    for _ in range(3):
        p = np.pad(img, ((2, 2), (2, 2), (0, 0)), mode="reflect")
        acc = np.zeros_like(img)
        for dy in range(5):
            for dx in range(5):
                acc += p[dy:dy + height, dx:dx + width]
        img = acc / 25.0
    img = (img - img.mean()) * 6.0 + 128.0  # restore contrast lost to the blur
    img = img + rs.randn(height, width, 3) * 10.0
    return np.clip(img, 0, 255).astype(np.uint8)


class SyntheticVAE:
    """Stand-in for the diffusers AutoencoderKL object that process() receives (test_scripts/inference.py:104-117,142):
    `.config.scaling_factor`, `.encode(x).latent_dist.mode()`, `.decode(z).sample`.

    The VAE *encoder* is outside the hot path ("next" row 1), so encode() is a fixed seeded 8x8 stride-8
    projection that produces latents of the right shape and scale; decode() is supplied by the caller (the CUDA
    decoder in the product, the reference / oracle Decoder when minting goldens)."""

    class _Cfg:
        scaling_factor = 0.18215  # sd-vae-ft-ema, configs/PixArt_xl2_internal.py:49

    class _Out:
        def __init__(self, t):
            self.sample = t

    class _Dist:
        def __init__(self, t):
            self._t = t

        def mode(self):
            return self._t

    class _Enc:
        def __init__(self, t):
            self.latent_dist = SyntheticVAE._Dist(t)

    def __init__(self, decode_fn, seed: int = 3):
        gen = torch.Generator(device="cpu").manual_seed(seed)
        self.enc_w = torch.randn(4, 3, 8, 8, generator=gen) * (11.0 / 192 ** 0.5)
        self.decode_fn = decode_fn
        self.config = SyntheticVAE._Cfg()

    def encode(self, x):
        w = self.enc_w.to(device=x.device, dtype=x.dtype)
        return SyntheticVAE._Enc(torch.nn.functional.conv2d(x, w, stride=8))

    def decode(self, z):
        return SyntheticVAE._Out(self.decode_fn(z))
