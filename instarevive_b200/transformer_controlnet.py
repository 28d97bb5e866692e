"""Flavour (B) of the generator: the diffusers call signature the released CLI and the DMD scripts use
 on top of the same sm_100a kernels.

  * `Transformer2DModel` stands in for `diffusers.Transformer2DModel` as test_scripts/inference.py:238-242 uses it
    (`model(latents, timestep=..., encoder_hidden_states=..., encoder_attention_mask=..., added_cond_kwargs=...).sample`,
    `model.config.sample_size / out_channels`, `load_state_dict` of the released `InstaRevive_v1.ckpt` keys);
  * `ControlTransformerHalf(base_model, copy_blocks_num=13)` mirrors diffusion/model/nets/transformer_controlnet.py:56-173
    (`forward(..., c=...)` returns the bare tensor when `return_dict` is true and a 1-tuple otherwise -- the reference's
    own asymmetry, which scripts/DMD/transformer_train/generate.py:66-82 relies on).

Both are thin argument adapters over `ControlPixArtMSHalf.forward` (flavour (A), nets.py): the two flavours are the same
network key for key (tools/convert_pixart_to_diffusers.py:29-160; instarevive_b200.convert), so no arithmetic lives here.

Differences that are kept, on purpose:
  * `config.sample_size == 64` (the 512 px DMD checkpoint) has no resolution / aspect-ratio embedders in diffusers
    (`use_additional_conditions` false): the size embedders of the flavour-(A) container get zero output layers, which
    adds an exact 0 to the timestep embedding;
  * caption mask: a 2-D `(B, L)` mask is a hard mask in diffusers (converted to a -10000 bias) and here (valid tokens are
    packed, pixart_controlnet.py:222-228). The CLI passes a 3-D float mask of ones and zeros (inference.py:274-277),
    which diffusers 0.30 would add to the scores as a bias instead of masking (SURVEY 8f row 3 caveat; not verifiable
    offline). Here a 3-D mask is read as "non-zero = valid token" -- the behaviour flavour (A) was trained with.
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Any, Dict, Mapping, Optional

import torch
import torch.nn as nn

from . import convert
from .nets import ControlPixArtMSHalf, PixArtMS


class Transformer2DModelOutput:
    """diffusers.models.modeling_outputs.Transformer2DModelOutput: the one member callers read."""

    def __init__(self, sample: torch.Tensor):
        self.sample = sample


def _as_caption(encoder_hidden_states: torch.Tensor) -> torch.Tensor:
    """diffusers passes (B, L, 4096); flavour (A) wants (B, 1, L, 4096)."""
    if encoder_hidden_states is None:
        raise TypeError("encoder_hidden_states (the T5 caption embedding) is required")
    if encoder_hidden_states.dim() == 3:
        return encoder_hidden_states.unsqueeze(1)
    if encoder_hidden_states.dim() == 4 and encoder_hidden_states.shape[1] == 1:
        return encoder_hidden_states
    raise ValueError(f"encoder_hidden_states must be (B, L, C) or (B, 1, L, C), got {tuple(encoder_hidden_states.shape)}")


def _as_mask(encoder_attention_mask: Optional[torch.Tensor], n_tokens: int) -> Optional[torch.Tensor]:
    """(B, L) keep-mask or (B, 1, L) / (B, 1, 1, L) one/zero mask -> flavour (A)'s (B, 1, 1, L) integer mask (non-zero =
    valid)."""
    m = encoder_attention_mask
    if m is None:
        return None
    while m.dim() > 2 and m.shape[1] == 1:   # (B, 1, L) of the CLI, (B, 1, 1, L) of flavour (A) callers
        m = m[:, 0]
    if m.dim() != 2 or m.shape[1] != n_tokens:
        raise ValueError(f"encoder_attention_mask must be (B, {n_tokens}) or (B, 1, {n_tokens}), got "
                         f"{tuple(encoder_attention_mask.shape)}")
    return (m != 0).to(torch.int32)[:, None, None, :]


class _FlavourB(nn.Module):
    """Shared plumbing: a flavour-(A) ControlPixArtMSHalf inside, diffusers' keyword surface outside."""

    def __init__(self, sample_size: int, copy_blocks_num: int, depth: int = 28, caption_channels: int = 4096,
                 model_max_length: int = 120, interpolation_scale: Optional[float] = None):
        super().__init__()
        if sample_size not in (32, 64, 128, 256):
            raise ValueError("sample_size must be the latent side of a PixArt checkpoint (32, 64, 128 or 256)")
        # diffusers: interpolation_scale = max(sample_size // 64, 1); use_additional_conditions = sample_size == 128
        scale = float(interpolation_scale if interpolation_scale is not None else max(sample_size // 64, 1))
        base = PixArtMS(depth=depth, input_size=sample_size, pe_interpolation=scale, caption_channels=caption_channels,
                        model_max_length=model_max_length, micro_condition=True, init_weights=False)
        self.net = ControlPixArtMSHalf(base, copy_blocks_num).eval()
        self.config = SimpleNamespace(sample_size=sample_size, patch_size=2, in_channels=4, out_channels=8,
                                      num_layers=depth, num_attention_heads=16, attention_head_dim=72,
                                      caption_channels=caption_channels, norm_type="ada_norm_single",
                                      interpolation_scale=scale)
        self.use_additional_conditions = sample_size == 128
        if not self.use_additional_conditions:
            self._zero_size_embedders()

    # ------------------------------------------------------------------ weights
    def _zero_size_embedders(self) -> None:
        for e in (self.net.base_model.csize_embedder, self.net.base_model.ar_embedder):
            for p in e.parameters():
                nn.init.zeros_(p)
        self.net._packed_version = None

    def load_state_dict(self, state_dict: Mapping[str, Any], strict: bool = True):
        """Accepts the diffusers layout (bare `transformer_blocks.*` or `base_model.* / controlnet.*`) and the PixArt
        layout. Checkpoints of the 512 px model carry no resolution / aspect-ratio embedder; those stay zero."""
        sd = dict(state_dict)
        if convert.is_diffusers_layout(sd):
            sd = dict(convert.diffusers_to_pixart(sd))
        wrapped = any(k.startswith("base_model.") or k.startswith("controlnet.") for k in sd)
        own = self.net.state_dict() if wrapped else self.net.base_model.state_dict()
        for k, v in own.items():
            if k in sd:
                continue
            is_size = "csize_embedder." in k or "ar_embedder." in k
            is_buf = k.endswith("y_embedder.y_embedding") or k.split(".")[-1] == "pos_embed"
            if is_size and not self.use_additional_conditions:
                sd[k] = torch.zeros_like(v)
            elif is_buf:   # PixArt-only buffers the diffusers layout does not carry (convert_pixart_to_diffusers.py:194-198)
                sd[k] = v
        return self.net.load_state_dict(sd, strict)

    def state_dict(self, *a, **k):
        """diffusers layout, like the module this class stands in for."""
        sd = self.net.state_dict(*a, **k)
        if self.net.copy_blocks_num == 0:
            sd = {key[len("base_model."):]: v for key, v in sd.items() if key.startswith("base_model.")}
        out = convert.pixart_to_diffusers(sd)
        if not self.use_additional_conditions:
            out = type(out)((key, v) for key, v in out.items()
                            if "resolution_embedder" not in key and "aspect_ratio_embedder" not in key)
        return out

    # ------------------------------------------------------------------ device / dtype plumbing
    @property
    def dtype(self):
        return self.net.dtype

    @property
    def device(self):
        return self.net.device

    # ------------------------------------------------------------------ the call
    def _run(self, hidden_states, encoder_hidden_states, timestep, added_cond_kwargs, encoder_attention_mask, c):
        if timestep is None:
            raise TypeError("timestep is required")
        y = _as_caption(encoder_hidden_states)
        mask = _as_mask(encoder_attention_mask, y.shape[2])
        bs, _, h, w = hidden_states.shape
        dev = hidden_states.device
        res = ar = None
        if self.use_additional_conditions:
            if added_cond_kwargs is None or added_cond_kwargs.get("resolution") is None \
                    or added_cond_kwargs.get("aspect_ratio") is None:
                # diffusers raises the same way for the 1024 px model (Transformer2DModel.forward, ada_norm_single)
                raise ValueError("`added_cond_kwargs` with 'resolution' and 'aspect_ratio' is required when sample_size is 128")
            res, ar = added_cond_kwargs["resolution"], added_cond_kwargs["aspect_ratio"]
        else:  # size embedders are zero: any finite value gives +0
            res = torch.tensor([[float(8 * h), float(8 * w)]], device=dev).repeat(bs, 1)
            ar = torch.tensor([[float(h) / float(w)]], device=dev).repeat(bs, 1)
        ts = torch.as_tensor(timestep, device=dev).reshape(-1).float()
        if ts.numel() == 1:
            ts = ts.expand(bs)
        data_info = {"img_hw": res.to(dev).float().reshape(-1, 2), "aspect_ratio": ar.to(dev).float().reshape(-1, 1)}
        return self.net(hidden_states, ts, y, mask=mask, data_info=data_info, c=c)


class Transformer2DModel(_FlavourB):
    """The PixArt configuration of diffusers.Transformer2DModel (norm_type ada_norm_single, 28 x 16 x 72, patch 2,
    caption_channels 4096) as test_scripts/inference.py:238-242 and test_scripts/test_dmd*.py instantiate it."""

    def __init__(self, sample_size: int = 64, num_layers: int = 28, caption_channels: int = 4096,
                 model_max_length: int = 120, interpolation_scale: Optional[float] = None, **unused):
        super().__init__(sample_size, 0, depth=num_layers, caption_channels=caption_channels,
                         model_max_length=model_max_length, interpolation_scale=interpolation_scale)

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                timestep: Optional[torch.Tensor] = None, added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None,
                class_labels=None, cross_attention_kwargs: Optional[Dict[str, Any]] = None, attention_mask=None,
                encoder_attention_mask: Optional[torch.Tensor] = None, return_dict: bool = True):
        if attention_mask is not None or class_labels is not None or cross_attention_kwargs:
            raise NotImplementedError("attention_mask / class_labels / cross_attention_kwargs are unused by the PixArt "
                                      "configuration and by every caller in the reference")
        out = self._run(hidden_states, encoder_hidden_states, timestep, added_cond_kwargs, encoder_attention_mask, None)
        return Transformer2DModelOutput(out) if return_dict else (out,)


class ControlTransformerHalf(_FlavourB):
    """diffusion/model/nets/transformer_controlnet.py:56-173. `base_model` is a Transformer2DModel of this module; its
    weights are copied (the first `copy_blocks_num` blocks also into the control branch, :69-70)."""

    def __init__(self, base_model: Transformer2DModel, copy_blocks_num: int = 13) -> None:
        cfg = base_model.config
        super().__init__(cfg.sample_size, copy_blocks_num, depth=cfg.num_layers, caption_channels=cfg.caption_channels,
                         model_max_length=base_model.net.base_model.y_embedder.y_embedding.shape[0],
                         interpolation_scale=cfg.interpolation_scale)
        self.copy_blocks_num = copy_blocks_num
        self.total_blocks_num = cfg.num_layers
        src = base_model.net.base_model
        self.net.base_model.load_state_dict(src.state_dict(), strict=True)
        for i, blk in enumerate(self.net.controlnet):
            blk.copied_block.load_state_dict(src.blocks[i].state_dict(), strict=True)
        self.net.to(base_model.device)
        self.net._packed_version = None

    @torch.no_grad()
    def forward_c(self, c):
        """:77-87 -- patch embedding + position table of the control latent."""
        return self.net.forward_c(c)

    @torch.no_grad()
    def forward(self, hidden_states: torch.Tensor, encoder_hidden_states: Optional[torch.Tensor] = None,
                timestep: Optional[torch.Tensor] = None, added_cond_kwargs: Optional[Dict[str, torch.Tensor]] = None,
                class_labels=None, cross_attention_kwargs: Optional[Dict[str, Any]] = None, attention_mask=None,
                encoder_attention_mask: Optional[torch.Tensor] = None, c: Optional[torch.Tensor] = None,
                return_dict: bool = True):
        if attention_mask is not None or class_labels is not None or cross_attention_kwargs:
            raise NotImplementedError("attention_mask / class_labels / cross_attention_kwargs are unused by the PixArt "
                                      "configuration and by every caller in the reference")
        out = self._run(hidden_states, encoder_hidden_states, timestep, added_cond_kwargs, encoder_attention_mask, c)
        return out if return_dict else (out,)   # the reference returns the bare tensor here (:170-173)
