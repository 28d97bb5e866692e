"""Weight-format conversion between the diffusers `Transformer2DModel` layout of the released checkpoints and the PixArt
layout this package (and the reference's nets (A)) consumes -- SURVEY 8f row 3.

The forward direction (PixArt -> diffusers) is what the reference ships as tools/convert_pixart_to_diffusers.py:29-154:
q/k/v and k/v weights are split out of the fused `attn.qkv` / `cross_attn.kv_linear`, `t_embedder` / `t_block` /
`csize_embedder` / `ar_embedder` move under `adaln_single.*`, `x_embedder` becomes `pos_embed.proj`, `y_embedder.y_proj`
becomes `caption_projection`, `final_layer` becomes `proj_out` + `scale_shift_table`. `diffusers_to_pixart` inverts that
mapping (re-fusing q/k/v in the order the converter chunks them), for a bare transformer state dict and for the
`base_model.* / controlnet.N.copied_block.*` layout of the ControlNet wrappers
(diffusion/model/nets/transformer_controlnet.py:17-73 wraps `transformer_blocks[i]` the way pixart_controlnet.py wraps
`blocks[i]`; `before_proj` / `after_proj` keep their names). Pure host-side tensor bookkeeping: no arithmetic.

Only the weight mapping is claimed here; the flavour-(B) call signature differs from flavour (A) in how it treats the
caption mask (SURVEY 8f row 3 caveat) and is not reproduced.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Mapping

import torch

# (PixArt key, diffusers key) pairs that map one to one -- tools/convert_pixart_to_diffusers.py:30-78,152-154
_TOP = [
    ("x_embedder.proj.weight", "pos_embed.proj.weight"),
    ("x_embedder.proj.bias", "pos_embed.proj.bias"),
    ("y_embedder.y_proj.fc1.weight", "caption_projection.linear_1.weight"),
    ("y_embedder.y_proj.fc1.bias", "caption_projection.linear_1.bias"),
    ("y_embedder.y_proj.fc2.weight", "caption_projection.linear_2.weight"),
    ("y_embedder.y_proj.fc2.bias", "caption_projection.linear_2.bias"),
    ("t_embedder.mlp.0.weight", "adaln_single.emb.timestep_embedder.linear_1.weight"),
    ("t_embedder.mlp.0.bias", "adaln_single.emb.timestep_embedder.linear_1.bias"),
    ("t_embedder.mlp.2.weight", "adaln_single.emb.timestep_embedder.linear_2.weight"),
    ("t_embedder.mlp.2.bias", "adaln_single.emb.timestep_embedder.linear_2.bias"),
    ("csize_embedder.mlp.0.weight", "adaln_single.emb.resolution_embedder.linear_1.weight"),
    ("csize_embedder.mlp.0.bias", "adaln_single.emb.resolution_embedder.linear_1.bias"),
    ("csize_embedder.mlp.2.weight", "adaln_single.emb.resolution_embedder.linear_2.weight"),
    ("csize_embedder.mlp.2.bias", "adaln_single.emb.resolution_embedder.linear_2.bias"),
    ("ar_embedder.mlp.0.weight", "adaln_single.emb.aspect_ratio_embedder.linear_1.weight"),
    ("ar_embedder.mlp.0.bias", "adaln_single.emb.aspect_ratio_embedder.linear_1.bias"),
    ("ar_embedder.mlp.2.weight", "adaln_single.emb.aspect_ratio_embedder.linear_2.weight"),
    ("ar_embedder.mlp.2.bias", "adaln_single.emb.aspect_ratio_embedder.linear_2.bias"),
    ("t_block.1.weight", "adaln_single.linear.weight"),
    ("t_block.1.bias", "adaln_single.linear.bias"),
    ("final_layer.linear.weight", "proj_out.weight"),
    ("final_layer.linear.bias", "proj_out.bias"),
    ("final_layer.scale_shift_table", "scale_shift_table"),
]
# per block, one to one -- convert_pixart_to_diffusers.py:82-84,97-102,118-129,144-149
_BLOCK = [
    ("scale_shift_table", "scale_shift_table"),
    ("attn.proj.weight", "attn1.to_out.0.weight"),
    ("attn.proj.bias", "attn1.to_out.0.bias"),
    ("mlp.fc1.weight", "ff.net.0.proj.weight"),
    ("mlp.fc1.bias", "ff.net.0.proj.bias"),
    ("mlp.fc2.weight", "ff.net.2.weight"),
    ("mlp.fc2.bias", "ff.net.2.bias"),
    ("cross_attn.q_linear.weight", "attn2.to_q.weight"),
    ("cross_attn.q_linear.bias", "attn2.to_q.bias"),
    ("cross_attn.proj.weight", "attn2.to_out.0.weight"),
    ("cross_attn.proj.bias", "attn2.to_out.0.bias"),
]
# PixArt buffers the converter drops (:194-198) and diffusers buffers that have no PixArt parameter
_PIXART_ONLY = ("y_embedder.y_embedding", "pos_embed")
_DIFFUSERS_ONLY = ("pos_embed.pos_embed",)


def _block_to_diffusers(src: Mapping[str, torch.Tensor], sp: str, dst: Dict[str, torch.Tensor], dp: str) -> None:
    for a, b in _BLOCK:
        dst[dp + b] = src[sp + a]
    for wb in ("weight", "bias"):   # :88-95 torch.chunk(qkv, 3, dim=0); :133-142 torch.chunk(kv, 2, dim=0)
        q, k, v = torch.chunk(src[f"{sp}attn.qkv.{wb}"], 3, dim=0)
        dst[f"{dp}attn1.to_q.{wb}"], dst[f"{dp}attn1.to_k.{wb}"], dst[f"{dp}attn1.to_v.{wb}"] = q, k, v
        k2, v2 = torch.chunk(src[f"{sp}cross_attn.kv_linear.{wb}"], 2, dim=0)
        dst[f"{dp}attn2.to_k.{wb}"], dst[f"{dp}attn2.to_v.{wb}"] = k2, v2


def _block_to_pixart(src: Mapping[str, torch.Tensor], sp: str, dst: Dict[str, torch.Tensor], dp: str) -> None:
    for a, b in _BLOCK:
        dst[dp + a] = src[sp + b]
    for wb in ("weight", "bias"):
        dst[f"{dp}attn.qkv.{wb}"] = torch.cat([src[f"{sp}attn1.to_q.{wb}"], src[f"{sp}attn1.to_k.{wb}"],
                                               src[f"{sp}attn1.to_v.{wb}"]], dim=0)
        dst[f"{dp}cross_attn.kv_linear.{wb}"] = torch.cat([src[f"{sp}attn2.to_k.{wb}"], src[f"{sp}attn2.to_v.{wb}"]], dim=0)


def _count(sd: Mapping[str, torch.Tensor], prefix: str, tail: str) -> int:
    n = 0
    while f"{prefix}{n}.{tail}" in sd:
        n += 1
    return n


def pixart_to_diffusers(sd: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """PixArt layout (bare, or base_model.* + controlnet.*) -> diffusers Transformer2DModel layout."""
    wrapped = any(k.startswith("base_model.") for k in sd)
    bp = "base_model." if wrapped else ""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for a, b in _TOP:
        if bp + a in sd:   # csize / ar embedders exist only with micro_condition (:49-75)
            out[bp + b] = sd[bp + a]
    for i in range(_count(sd, f"{bp}blocks.", "scale_shift_table")):
        _block_to_diffusers(sd, f"{bp}blocks.{i}.", out, f"{bp}transformer_blocks.{i}.")
    for j in range(_count(sd, "controlnet.", "copied_block.scale_shift_table")):
        _block_to_diffusers(sd, f"controlnet.{j}.copied_block.", out, f"controlnet.{j}.copied_block.")
        for name in ("before_proj", "after_proj"):
            for wb in ("weight", "bias"):
                if f"controlnet.{j}.{name}.{wb}" in sd:
                    out[f"controlnet.{j}.{name}.{wb}"] = sd[f"controlnet.{j}.{name}.{wb}"]
    return out


def diffusers_to_pixart(sd: Mapping[str, torch.Tensor]) -> "OrderedDict[str, torch.Tensor]":
    """diffusers Transformer2DModel layout (bare, or base_model.* + controlnet.*) -> PixArt layout, ready for
    ControlPixArtMSHalf.load_state_dict (which also accepts bare PixArt keys, pixart_controlnet.py:151-163)."""
    wrapped = any(k.startswith("base_model.") for k in sd)
    bp = "base_model." if wrapped else ""
    out: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for a, b in _TOP:
        if bp + b in sd:
            out[bp + a] = sd[bp + b]
    nblocks = _count(sd, f"{bp}transformer_blocks.", "scale_shift_table")
    if nblocks == 0:
        raise KeyError("no transformer_blocks.N.scale_shift_table keys: not a diffusers PixArt Transformer2DModel state dict")
    for i in range(nblocks):
        _block_to_pixart(sd, f"{bp}transformer_blocks.{i}.", out, f"{bp}blocks.{i}.")
    for j in range(_count(sd, "controlnet.", "copied_block.scale_shift_table")):
        _block_to_pixart(sd, f"controlnet.{j}.copied_block.", out, f"controlnet.{j}.copied_block.")
        for name in ("before_proj", "after_proj"):
            for wb in ("weight", "bias"):
                if f"controlnet.{j}.{name}.{wb}" in sd:
                    out[f"controlnet.{j}.{name}.{wb}"] = sd[f"controlnet.{j}.{name}.{wb}"]
    return out


def is_diffusers_layout(sd: Mapping[str, torch.Tensor]) -> bool:
    return any(".transformer_blocks." in k or k.startswith("transformer_blocks.") for k in sd)
