"""Host-side mirror of the reference's operator surface for the one-step restoration forward.

`PixArtMS` / `PixArtMS_XL_2` / `ControlPixArtMSHalf` keep the reference's constructor arguments, `state_dict` keys
(668 tensors for XL/2 + 13 copied blocks), `forward(x, timestep, y, mask, data_info, c)` signature,
`forward_with_dpmsolver`, `forward_c`, `.dtype` and the attribute fall-through to `base_model`
(diffusion/model/nets/pixart_controlnet.py:54-251, PixArtMS.py:86-293). The modules are *parameter containers*:
all arithmetic happens in libinstarevive_b200.so (hand-written sm_100a kernels) through the C ABI of
include/instarevive_b200.h. There is no PyTorch / CPU fallback: calling forward without the CUDA library or on a
CPU tensor raises.
"""
from __future__ import annotations

import ctypes as C
import re
from collections import OrderedDict
from typing import Any, Mapping, Optional

import torch
import torch.nn as nn

from . import _lib


# ------------------------------------------------------------------------------------------------ containers
class _Mlp(nn.Module):  # timm Mlp parameter layout: fc1, fc2
    def __init__(self, i, h, o):
        super().__init__()
        self.fc1 = nn.Linear(i, h)
        self.fc2 = nn.Linear(h, o)


class _SelfAttn(nn.Module):  # AttentionKVCompress parameter layout (PixArt_blocks.py:61-96): qkv, proj
    def __init__(self, d):
        super().__init__()
        self.qkv = nn.Linear(d, 3 * d)
        self.proj = nn.Linear(d, d)


class _CrossAttn(nn.Module):  # MultiHeadCrossAttention (PixArt_blocks.py:28-41)
    def __init__(self, d):
        super().__init__()
        self.q_linear = nn.Linear(d, d)
        self.kv_linear = nn.Linear(d, 2 * d)
        self.proj = nn.Linear(d, d)


class PixArtMSBlock(nn.Module):
    """Parameter layout of PixArtMSBlock (PixArtMS.py:49-69); executed inside ir_dit_forward."""

    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0, **_):
        super().__init__()
        self.hidden_size = hidden_size
        self.attn = _SelfAttn(hidden_size)
        self.cross_attn = _CrossAttn(hidden_size)
        self.mlp = _Mlp(hidden_size, int(hidden_size * mlp_ratio), hidden_size)
        self.scale_shift_table = nn.Parameter(torch.randn(6, hidden_size) / hidden_size ** 0.5)


class _Embedder(nn.Module):  # TimestepEmbedder / SizeEmbedder: mlp = Sequential(Linear, SiLU, Linear)
    def __init__(self, hidden, freq=256):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(freq, hidden), nn.SiLU(), nn.Linear(hidden, hidden))


class _PatchEmbed(nn.Module):
    def __init__(self, patch, cin, d):
        super().__init__()
        self.patch_size = (patch, patch)
        self.proj = nn.Conv2d(cin, d, kernel_size=patch, stride=patch)


class _Caption(nn.Module):
    def __init__(self, cin, d, token_num):
        super().__init__()
        self.y_proj = _Mlp(cin, d, d)
        self.register_buffer("y_embedding", torch.randn(token_num, cin) / cin ** 0.5)


class _Final(nn.Module):
    def __init__(self, d, patch, cout):
        super().__init__()
        self.linear = nn.Linear(d, patch * patch * cout)
        self.scale_shift_table = nn.Parameter(torch.randn(2, d) / d ** 0.5)
        self.out_channels = cout


class PixArtMS(nn.Module):
    """Constructor surface of PixArtMS (PixArtMS.py:86-164). Only the XL/2 geometry is supported by the kernels."""

    def __init__(self, input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, learn_sigma=True, pred_sigma=True, drop_path: float = 0.0,
                 caption_channels=4096, pe_interpolation=1.0, config=None, model_max_length=120,
                 micro_condition=False, qk_norm=False, kv_compress_config=None, init_weights: bool = True, **kwargs):
        super().__init__()
        if hidden_size != 1152 or num_heads != 16 or patch_size != 2 or in_channels != 4 or int(mlp_ratio) != 4:
            raise ValueError("instarevive_b200 kernels are specialised for PixArt XL/2 (hidden 1152, 16 heads, patch 2)")
        if qk_norm or (kv_compress_config and kv_compress_config.get("kv_compress_layer")):
            raise ValueError("qk_norm / KV compression are outside the restoration hot path (SURVEY section 5)")
        if not pred_sigma:
            raise ValueError("pred_sigma=False (4 output channels) is not supported")
        self.pred_sigma = pred_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels * 2
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.pe_interpolation = pe_interpolation
        self.depth = depth
        self.hidden_size = hidden_size
        self.caption_channels = caption_channels
        self.base_size = input_size // patch_size
        self.h = self.w = 0
        self.micro_conditioning = micro_condition
        self.register_buffer("pos_embed", torch.zeros(1, (input_size // patch_size) ** 2, hidden_size))
        self.x_embedder = _PatchEmbed(patch_size, in_channels, hidden_size)
        self.t_embedder = _Embedder(hidden_size)
        self.t_block = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size))
        self.y_embedder = _Caption(caption_channels, hidden_size, model_max_length)
        self.blocks = nn.ModuleList([PixArtMSBlock(hidden_size, num_heads, mlp_ratio) for _ in range(depth)])
        self.final_layer = _Final(hidden_size, patch_size, self.out_channels)
        if micro_condition:
            self.csize_embedder = _Embedder(hidden_size // 3)
            self.ar_embedder = _Embedder(hidden_size // 3)
        if init_weights:
            self.initialize()

    def initialize(self):
        """PixArtMS.initialize (PixArtMS.py:250-285)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        embs = [self.t_embedder] + ([self.csize_embedder, self.ar_embedder] if self.micro_conditioning else [])
        for e in embs:
            nn.init.normal_(e.mlp[0].weight, std=0.02)
            nn.init.normal_(e.mlp[2].weight, std=0.02)
        nn.init.normal_(self.t_block[1].weight, std=0.02)
        nn.init.normal_(self.y_embedder.y_proj.fc1.weight, std=0.02)
        nn.init.normal_(self.y_embedder.y_proj.fc2.weight, std=0.02)
        for b in self.blocks:
            nn.init.constant_(b.cross_attn.proj.weight, 0)
            nn.init.constant_(b.cross_attn.proj.bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def forward(self, *a, **k):
        raise RuntimeError("PixArtMS is a parameter container here; wrap it in ControlPixArtMSHalf (c=None runs the "
                           "plain 28-block path)")


def PixArtMS_XL_2(**kwargs):
    """PixArtMS.py:291-293."""
    return PixArtMS(depth=28, hidden_size=1152, patch_size=2, num_heads=16, **kwargs)


class ControlT2IDitBlockHalf(nn.Module):
    """Parameter layout of ControlT2IDitBlockHalf (pixart_controlnet.py:17-36): copied block + zero linears."""

    def __init__(self, base_block: PixArtMSBlock, block_index: int = 0):
        super().__init__()
        d = base_block.hidden_size
        self.copied_block = PixArtMSBlock(d, 16)
        self.copied_block.load_state_dict(base_block.state_dict())
        self.block_index = block_index
        self.hidden_size = d
        if block_index == 0:
            self.before_proj = nn.Linear(d, d)
            nn.init.zeros_(self.before_proj.weight)
            nn.init.zeros_(self.before_proj.bias)
        self.after_proj = nn.Linear(d, d)
        nn.init.zeros_(self.after_proj.weight)
        nn.init.zeros_(self.after_proj.bias)


# ------------------------------------------------------------------------------------------------ the operator
class ControlPixArtMSHalf(nn.Module):
    """Drop-in for the reference's ControlPixArtMSHalf (pixart_controlnet.py:186-251) backed by sm_100a kernels."""

    def __init__(self, base_model: PixArtMS, copy_blocks_num: int = 13) -> None:
        super().__init__()
        if not getattr(base_model, "micro_conditioning", False):
            # the reference forward dereferences csize_embedder / ar_embedder unconditionally (:217-218)
            raise AttributeError("ControlPixArtMSHalf requires base_model built with micro_condition=True")
        self.base_model = base_model.eval()
        self.copy_blocks_num = copy_blocks_num
        self.total_blocks_num = len(base_model.blocks)
        for p in self.base_model.parameters():
            p.requires_grad_(False)
        self.controlnet = nn.ModuleList([ControlT2IDitBlockHalf(base_model.blocks[i], i) for i in range(copy_blocks_num)])
        self._handle: Optional[int] = None
        self._packed_version = None
        self._ws: Optional[torch.Tensor] = None
        self._cap_key = None
        self._cap_tensors = None

    # attribute fall-through of the reference (:70-76)
    def __getattr__(self, name: str):
        try:
            return super().__getattr__(name)
        except AttributeError:
            if name in ("base_model", "controlnet"):
                raise
            return getattr(super().__getattr__("base_model"), name)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):
        """Accepts full keys (base_model.* / controlnet.*) or bare PixArt keys (pixart_controlnet.py:151-163), and the
        diffusers Transformer2DModel layout of the released checkpoints (converted by instarevive_b200.convert; the
        PixArt buffers y_embedding / pos_embed that diffusers checkpoints do not carry keep their current values)."""
        self._packed_version = None
        from . import convert
        if convert.is_diffusers_layout(state_dict):
            state_dict = convert.diffusers_to_pixart(state_dict)
            own = self.state_dict() if any(k.startswith("base_model") for k in state_dict) else self.base_model.state_dict()
            for k, v in own.items():
                if k.endswith("y_embedder.y_embedding") or k.endswith("pos_embed"):
                    state_dict.setdefault(k, v)
        if all((k.startswith("base_model") or k.startswith("controlnet")) for k in state_dict.keys()):
            return super().load_state_dict(state_dict, strict)
        return self.base_model.load_state_dict(state_dict, strict)

    # ------------------------------------------------------------------ packing into the C library
    def _apply(self, fn, *a, **k):  # .to() / .cuda() / .float(): the packed copy is stale afterwards
        self._packed_version = None
        return super()._apply(fn, *a, **k)

    def pack(self, force: bool = False) -> None:
        """Upload (fp32 -> packed bf16 / fp32) every parameter into the library handle. Called lazily by forward
        after construction, load_state_dict() or .to(); call pack(force=True) after modifying parameters in place."""
        if self._handle is not None and not force and self._packed_version is not None:
            return
        dev = self.device
        if dev.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: move the module to a CUDA device first")
        ver = 1
        L = _lib.lib()
        with torch.cuda.device(dev):
            if self._handle is None:
                cfg = _lib.DitConfig(self.depth, self.copy_blocks_num, self.hidden_size, self.num_heads, self.patch_size,
                                     self.in_channels, self.out_channels, self.caption_channels, 4, self.base_size,
                                     float(self.pe_interpolation))
                h = C.c_void_p()
                _lib.check(L.ir_dit_create(C.byref(cfg), C.byref(h)), "ir_dit_create")
                self._handle = h.value
            sd = OrderedDict(self.named_parameters())
            n = L.ir_dit_num_params(self._handle)
            name = C.create_string_buffer(256)
            numel = C.c_longlong()
            for i in range(n):
                _lib.check(L.ir_dit_param_info(self._handle, i, name, 256, C.byref(numel), None, None), "param_info")
                key = name.value.decode()
                if key not in sd:
                    raise KeyError(f"instarevive_b200: module has no parameter '{key}' required by the CUDA library")
                t = sd[key].detach().to(device=dev, dtype=torch.float32).contiguous()
                if t.numel() != numel.value:
                    raise ValueError(f"parameter {key}: {t.numel()} elements, library expects {numel.value}")
                _lib.check(L.ir_dit_load_param(self._handle, key.encode(), t.data_ptr(), t.numel(), _lib.stream_ptr()),
                           f"load {key}")
                del t
            torch.cuda.current_stream().synchronize()
        self._packed_version = ver
        self._cap_key = None

    def set_cuda_graphs(self, enable: bool) -> None:
        """CUDA-graph replay of the forward (default on: from the third call with a given batch / latent size / caption
        layout the ~450 launches are replayed as one graph). Off = plain stream-ordered launches."""
        self.pack()
        _lib.check(_lib.lib().ir_dit_set_graphs(self._handle, 1 if enable else 0), "ir_dit_set_graphs")

    def set_dual_chain(self, enable: bool) -> None:
        """Run the ControlNet chain on a second stream ahead of the base chain (default on); off = one stream."""
        self.pack()
        _lib.check(_lib.lib().ir_dit_set_dual_chain(self._handle, 1 if enable else 0), "ir_dit_set_dual_chain")

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ir_dit_destroy(self._handle)
        except Exception:
            pass

    # ------------------------------------------------------------------ caption bookkeeping
    def _caption_tables(self, y: torch.Tensor, mask: Optional[torch.Tensor], bs: int):
        """Packing of the valid caption tokens (pixart_controlnet.py:221-231): returns device int32 tables
        (y_index, kv_off, kv_len), sum_l and whether the cached caption K/V can be reused. One host sync when the
        caption changes (the reference pays `mask.sum().tolist()` on every call)."""
        # The cache entry HOLDS the caption and mask tensors it was computed from and is hit only by the very same tensor
        # objects at an unchanged in-place version. (Keying on data_ptr alone is unsafe: the caching allocator hands a
        # freed caption's address to the next tensor of the same shape, whose _version starts at 0 again.)
        ref = self._cap_key
        if (ref is not None and ref[0] is y and ref[1] == y._version and ref[2] is mask
                and (mask is None or ref[3] == mask._version) and ref[4] == bs):
            return self._cap_tensors + (True,)
        key = (y, y._version, mask, None if mask is None else mask._version, bs)
        ny, lmax = y.shape[0], y.shape[2]
        if mask is not None:
            m = mask
            if m.shape[0] != ny:
                m = m.repeat(ny // m.shape[0], 1, 1, 1)
            m = (m.reshape(ny, lmax) != 0).cpu()
        else:
            m = torch.ones(ny, lmax, dtype=torch.bool)
        rows = [torch.nonzero(m[b]).flatten() + b * lmax for b in range(ny)]
        lens = [int(r.numel()) for r in rows]
        if min(lens) == 0:
            raise ValueError("every sample needs at least one valid caption token")
        offs = [sum(lens[:b]) for b in range(ny)]
        if ny == bs:
            kv_off, kv_len = offs, lens
        elif ny == 1:  # one caption shared by the whole batch (tiles of one image): K/V computed once
            kv_off, kv_len = [0] * bs, lens * bs
        else:
            raise ValueError(f"caption batch {ny} does not match latent batch {bs}")
        dev = y.device
        if max(lens) > 384:
            raise ValueError(f"captions of more than 384 valid tokens are not supported (got {max(lens)})")
        tens = (torch.cat(rows).to(torch.int32).to(dev), torch.tensor(kv_off, dtype=torch.int32, device=dev),
                torch.tensor(kv_len, dtype=torch.int32, device=dev), int(sum(lens)),
                # key-window size of the cross-attention kernel: its TMA window starts at the packed row rounded down to 8
                int(max((o % 8) + l for o, l in zip(kv_off, kv_len))), int(sum(kv_len)))
        self._cap_key, self._cap_tensors = key, tens
        return tens + (False,)

    def _workspace(self, nbytes: int, dev) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != dev:
            self._ws = None
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        return self._ws

    # ------------------------------------------------------------------ forward
    @torch.no_grad()
    def forward(self, x, timestep, y, mask=None, data_info=None, c=None, **kwargs):
        """x, c: (N,4,H,W); timestep: (N,); y: (N,1,L,4096) (or (1,1,L,4096) shared by the batch); mask: (N,1,1,L);
        data_info: {'img_hw': (N,2), 'aspect_ratio': (N,1)}  ->  (N,8,H,W) fp32 (eps, learned sigma)."""
        if self.training:
            raise RuntimeError("call .eval(): the restoration forward is inference-only (caption dropout is train-time)")
        if data_info is None:
            raise TypeError("data_info with 'img_hw' and 'aspect_ratio' is required (pixart_controlnet.py:206)")
        if x.device.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: inputs must be CUDA tensors")
        self.pack()
        L = _lib.lib()
        dev = x.device
        bs, _, H, W = x.shape
        self.h, self.w = H // self.patch_size, W // self.patch_size
        f32 = dict(device=dev, dtype=torch.float32)
        xx = x.to(**f32).contiguous()
        cc = None if c is None else c.to(**f32).contiguous()
        ts = timestep.to(**f32).reshape(-1)
        if ts.numel() != bs:
            ts = ts.expand(bs)
        ts = ts.contiguous()
        yy = y.to(**f32).contiguous()
        hw = data_info["img_hw"].to(**f32).reshape(-1, 2)
        ar = data_info["aspect_ratio"].to(**f32).reshape(-1)
        if hw.shape[0] != bs:
            hw = hw.repeat(bs // hw.shape[0], 1)
        if ar.shape[0] != bs:
            ar = ar.repeat(bs // ar.shape[0])
        hw, ar = hw.contiguous(), ar.contiguous()
        if y.dim() != 4 or y.shape[1] != 1 or y.shape[3] != self.caption_channels:
            raise ValueError(f"y must be (N,1,L,{self.caption_channels}), got {tuple(y.shape)}")
        y_index, kv_off, kv_len, sum_l, max_l, kv_total, reuse = self._caption_tables(y, mask, bs)
        out = torch.empty(bs, self.out_channels, H, W, **f32)
        with torch.cuda.device(dev):
            need = L.ir_dit_workspace_bytes(self._handle, bs, H, W, sum_l)
            ws = self._workspace(need, dev)
            _lib.check(L.ir_dit_forward(self._handle, xx.data_ptr(), _lib.ptr(cc), ts.data_ptr(), yy.data_ptr(),
                                        y_index.data_ptr(), kv_off.data_ptr(), kv_len.data_ptr(), hw.data_ptr(),
                                        ar.data_ptr(), out.data_ptr(), bs, H, W, sum_l, max_l, kv_total, int(reuse), ws.data_ptr(),
                                        ws.numel(), _lib.stream_ptr()), "ir_dit_forward")
        return out

    def forward_with_dpmsolver(self, x, t, y, data_info, c, **kwargs):
        """pixart_controlnet.py:141-143."""
        return self.forward(x, t, y, data_info=data_info, c=c, **kwargs).chunk(2, dim=1)[0]

    @torch.no_grad()
    def forward_c(self, c):
        """x_embedder(c) + pos_embed -> (N, T, D) (pixart_controlnet.py:78-87)."""
        if c is None:
            return c
        self.pack()
        bs, _, H, W = c.shape
        self.h, self.w = H // self.patch_size, W // self.patch_size
        cc = c.to(dtype=torch.float32).contiguous()
        out = torch.empty(bs, self.h * self.w, self.hidden_size, device=c.device, dtype=torch.float32)
        with torch.cuda.device(c.device):
            _lib.check(_lib.lib().ir_dit_patch_embed(self._handle, cc.data_ptr(), out.data_ptr(), bs, H, W,
                                                     _lib.stream_ptr()), "ir_dit_patch_embed")
        return out
