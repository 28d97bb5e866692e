"""ORACLE (test infrastructure, not product code): fp32 CPU restatement of the reference's one-step DiT path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module, and
only as the checker / CPU baseline. The product path (instarevive_b200/) never imports it.

Parity status: PINNED against outputs of the reference itself. The UNMODIFIED reference modules
(/root/reference/diffusion/model/nets/{pixart_controlnet,PixArtMS,PixArt_blocks,PixArt}.py) were executed in the
build container by oracle/make_goldens.py on the seeded weights of instarevive_b200/weights.py and their outputs are
committed under tests/golden/ (the reference ships no golden vectors or unit tests of its own);
tests/test_oracle.py checks this restatement against those files.

Everything is written as plain functions over a state_dict (reference key names), torch fp32 on the CPU.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ position table
def pos_embed_2d(embed_dim: int, gh: int, gw: int, pe_interpolation: float = 1.0, base_size: int = 16) -> np.ndarray:
    """get_2d_sincos_pos_embed, diffusion/model/nets/PixArt.py:258-307: float32 grid, float64 angles, the first half
    of the channels encodes grid[0] (the w-coordinate: `np.meshgrid(grid_w, grid_h)`, "w goes first")."""
    grid_h = np.arange(gh, dtype=np.float32) / (gh / base_size) / pe_interpolation      # PixArt.py:266
    grid_w = np.arange(gw, dtype=np.float32) / (gw / base_size) / pe_interpolation      # PixArt.py:267
    gx, gy = np.meshgrid(grid_w, grid_h)                                                # PixArt.py:268

    def one_d(dim, pos):                                                                # PixArt.py:287-307
        omega = np.arange(dim // 2, dtype=np.float64)
        omega /= dim / 2.0
        omega = 1.0 / 10000 ** omega
        out = np.einsum("m,d->md", pos.reshape(-1), omega)
        return np.concatenate([np.sin(out), np.cos(out)], axis=1)

    return np.concatenate([one_d(embed_dim // 2, gx), one_d(embed_dim // 2, gy)], axis=1)  # PixArt.py:282-285


# ------------------------------------------------------------------------------------------------ embedders
def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: int = 10000) -> torch.Tensor:
    """TimestepEmbedder.timestep_embedding, PixArt_blocks.py:336-353 (cos first, fp32)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _mlp2(sd, prefix, x):
    """Linear -> SiLU -> Linear, PixArt_blocks.py:329-333."""
    h = F.silu(F.linear(x, sd[f"{prefix}.mlp.0.weight"], sd[f"{prefix}.mlp.0.bias"]))
    return F.linear(h, sd[f"{prefix}.mlp.2.weight"], sd[f"{prefix}.mlp.2.bias"])


def size_embed(sd, prefix, s: torch.Tensor, bs: int) -> torch.Tensor:
    """SizeEmbedder.forward, PixArt_blocks.py:381-393: every scalar of s (b, dims) is embedded separately and the
    results are concatenated per sample."""
    if s.ndim == 1:
        s = s[:, None]
    if s.shape[0] != bs:
        s = s.repeat(bs // s.shape[0], 1)
    b, dims = s.shape
    emb = _mlp2(sd, prefix, timestep_embedding(s.reshape(-1)))
    return emb.reshape(b, dims * emb.shape[-1])


# ------------------------------------------------------------------------------------------------ attention
def _mha(q, k, v, heads):
    """softmax(q k^T / sqrt(hd)) v over (L, D) tensors: xformers.ops.memory_efficient_attention semantics
    (default scale = head_dim^-1/2), call sites PixArt_blocks.py:52-53,153."""
    Lq, D = q.shape
    hd = D // heads
    qh = q.view(Lq, heads, hd).transpose(0, 1)
    kh = k.view(-1, heads, hd).transpose(0, 1)
    vh = v.view(-1, heads, hd).transpose(0, 1)
    w = torch.softmax(qh @ kh.transpose(1, 2) * hd ** -0.5, dim=-1)
    return (w @ vh).transpose(0, 1).reshape(Lq, D)


def block_forward(sd, p, x, y_packed, y_lens, t0, heads=16):
    """PixArtMSBlock.forward, PixArtMS.py:71-79. x: (B, T, D); y_packed: (sumL, D); t0: (B, 6D)."""
    B, T, D = x.shape
    mod = sd[f"{p}.scale_shift_table"][None] + t0.reshape(B, 6, D)                      # PixArtMS.py:74
    shift_msa, scale_msa, gate_msa, shift_mlp, scale_mlp, gate_mlp = mod.chunk(6, dim=1)
    # self-attention, AttentionKVCompress.forward with sr_ratio 1 (PixArt_blocks.py:123-158)
    h = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_msa) + shift_msa                   # PixArt_blocks.py:24-25
    qkv = F.linear(h, sd[f"{p}.attn.qkv.weight"], sd[f"{p}.attn.qkv.bias"]).reshape(B, T, 3, D)
    a = torch.stack([_mha(qkv[b, :, 0], qkv[b, :, 1], qkv[b, :, 2], heads) for b in range(B)])
    a = F.linear(a, sd[f"{p}.attn.proj.weight"], sd[f"{p}.attn.proj.bias"])
    x = x + gate_msa * a                                                                # PixArtMS.py:75
    # cross-attention, MultiHeadCrossAttention.forward (PixArt_blocks.py:43-58), block-diagonal over samples
    q = F.linear(x, sd[f"{p}.cross_attn.q_linear.weight"], sd[f"{p}.cross_attn.q_linear.bias"])
    kv = F.linear(y_packed, sd[f"{p}.cross_attn.kv_linear.weight"], sd[f"{p}.cross_attn.kv_linear.bias"])
    kv = kv.view(-1, 2, D)
    outs, s = [], 0
    for b in range(B):
        L = y_lens[b]
        outs.append(_mha(q[b], kv[s:s + L, 0], kv[s:s + L, 1], heads))
        s += L
    c = F.linear(torch.stack(outs), sd[f"{p}.cross_attn.proj.weight"], sd[f"{p}.cross_attn.proj.bias"])
    x = x + c                                                                           # PixArtMS.py:76
    # MLP (timm Mlp: fc1 -> GELU(tanh) -> fc2), PixArtMS.py:66-67,77
    h = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale_mlp) + shift_mlp
    h = F.gelu(F.linear(h, sd[f"{p}.mlp.fc1.weight"], sd[f"{p}.mlp.fc1.bias"]), approximate="tanh")
    h = F.linear(h, sd[f"{p}.mlp.fc2.weight"], sd[f"{p}.mlp.fc2.bias"])
    return x + gate_mlp * h


# ------------------------------------------------------------------------------------------------ whole forward
@torch.no_grad()
def control_pixart_forward(sd, x, timestep, y, mask=None, data_info=None, c=None, *, depth=28, copy_blocks=13,
                           heads=16, base_size=32, pe_interpolation=1.0):
    """ControlPixArtMSHalf.forward, diffusion/model/nets/pixart_controlnet.py:191-251 (eval mode, fp32)."""
    sd = {k: v.float() for k, v in sd.items()}
    bm = "base_model"
    B, _, H, W = x.shape
    D = sd[f"{bm}.x_embedder.proj.bias"].shape[0]
    gh, gw = H // 2, W // 2
    pos = torch.from_numpy(pos_embed_2d(D, gh, gw, pe_interpolation, base_size)).unsqueeze(0).float()  # :209-214

    def embed(z):                                                                       # PixArtMS.py:42-44
        t = F.conv2d(z.float(), sd[f"{bm}.x_embedder.proj.weight"], sd[f"{bm}.x_embedder.proj.bias"], stride=2)
        return t.flatten(2).transpose(1, 2) + pos

    if c is not None:
        c = embed(c)                                                                    # :199-201, 78-87
    x = embed(x)                                                                        # :215
    t = _mlp2(sd, f"{bm}.t_embedder", timestep_embedding(timestep.float()))             # :216
    csize = size_embed(sd, f"{bm}.csize_embedder", data_info["img_hw"].float(), B)      # :217
    ar = size_embed(sd, f"{bm}.ar_embedder", data_info["aspect_ratio"].float(), B)      # :218
    t = t + torch.cat([csize, ar], dim=1)                                               # :219
    t0 = F.linear(F.silu(t), sd[f"{bm}.t_block.1.weight"], sd[f"{bm}.t_block.1.bias"])  # :220
    # CaptionEmbedder.forward in eval mode = y_proj (timm Mlp, GELU tanh), PixArt_blocks.py:455-463
    yy = y.float()
    yy = F.linear(F.gelu(F.linear(yy, sd[f"{bm}.y_embedder.y_proj.fc1.weight"], sd[f"{bm}.y_embedder.y_proj.fc1.bias"]),
                         approximate="tanh"),
                  sd[f"{bm}.y_embedder.y_proj.fc2.weight"], sd[f"{bm}.y_embedder.y_proj.fc2.bias"])
    if mask is not None:                                                                # :222-228
        if mask.shape[0] != yy.shape[0]:
            mask = mask.repeat(yy.shape[0] // mask.shape[0], 1, 1, 1)
        m = mask.squeeze(1).squeeze(1)
        y_packed = yy.squeeze(1).masked_select(m.unsqueeze(-1) != 0).view(-1, D)
        y_lens = [int(v) for v in m.sum(dim=1).tolist()]
    else:                                                                               # :229-231
        y_lens = [yy.shape[2]] * yy.shape[0]
        y_packed = yy.squeeze(1).reshape(-1, D)

    x = block_forward(sd, f"{bm}.blocks.0", x, y_packed, y_lens, t0, heads)             # :234
    if c is not None:
        for i in range(1, copy_blocks + 1):                                             # :238-240
            p = f"controlnet.{i - 1}"
            if i == 1:                                                                  # pixart_controlnet.py:40-44
                c = F.linear(c, sd[f"{p}.before_proj.weight"], sd[f"{p}.before_proj.bias"])
                c = block_forward(sd, f"{p}.copied_block", x + c, y_packed, y_lens, t0, heads)
            else:                                                                       # pixart_controlnet.py:45-48
                c = block_forward(sd, f"{p}.copied_block", c, y_packed, y_lens, t0, heads)
            c_skip = F.linear(c, sd[f"{p}.after_proj.weight"], sd[f"{p}.after_proj.bias"])
            x = block_forward(sd, f"{bm}.blocks.{i}", x + c_skip, y_packed, y_lens, t0, heads)
        rest = range(copy_blocks + 1, depth)                                            # :243-244
    else:
        rest = range(1, depth)                                                          # :245-247
    for i in rest:
        x = block_forward(sd, f"{bm}.blocks.{i}", x, y_packed, y_lens, t0, heads)

    # T2IFinalLayer.forward (PixArt_blocks.py:271-275): modulation from t, NOT t0
    shift, scale = (sd[f"{bm}.final_layer.scale_shift_table"][None] + t[:, None]).chunk(2, dim=1)
    x = F.layer_norm(x, (D,), eps=1e-6) * (1 + scale) + shift
    x = F.linear(x, sd[f"{bm}.final_layer.linear.weight"], sd[f"{bm}.final_layer.linear.bias"])
    # unpatchify, pixart_controlnet.py:165-177
    cout = x.shape[-1] // 4
    x = x.reshape(B, gh, gw, 2, 2, cout)
    x = torch.einsum("nhwpqc->nchpwq", x)
    return x.reshape(B, cout, gh * 2, gw * 2)


# ------------------------------------------------------------------------------------------------ one-step math
def alphas_cumprod(num_steps: int = 1000, beta_start: float = 1e-4, beta_end: float = 2e-2) -> np.ndarray:
    """Linear beta schedule (get_named_beta_schedule("linear"), diffusion/model/gaussian_diffusion.py:99-116, which
    equals diffusers DDPMScheduler(beta_schedule="linear") used at test_scripts/inference.py:36)."""
    betas = np.linspace(beta_start, beta_end, num_steps, dtype=np.float64)
    return np.cumprod(1.0 - betas, axis=0)


def eps_to_mu(model_output, sample, t: int = 400):
    """eps_to_mu, scripts/DMD/transformer_train/generate.py:44-51 (alphas_cumprod cast to the sample dtype)."""
    a = torch.tensor(alphas_cumprod()[t], dtype=torch.float64).to(sample.dtype)
    return (sample - (1 - a) ** 0.5 * model_output) / a ** 0.5


def generate_sample_1step(sd, latents, y, mask, t: int = 400, **cfg):
    """generate_sample_1step + forward_model on operator surface (A): x = c = latents, timestep 400, learned-sigma half
    dropped (generate.py:22-42,84-85; the authors feed the degraded latent as control, test_controlnet.py:137-139)."""
    B, _, H, W = latents.shape
    data_info = {"img_hw": torch.tensor([[H * 8.0, W * 8.0]] * B), "aspect_ratio": torch.tensor([[H / W]] * B)}
    ts = torch.full((B,), float(t))
    out = control_pixart_forward(sd, latents, ts, y, mask, data_info, c=latents, **cfg)
    eps = out.chunk(2, dim=1)[0]
    return eps_to_mu(eps, latents, t)
