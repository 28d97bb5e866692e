"""One-step generator math with the reference's function names (scripts/DMD/transformer_train/generate.py:22-87),
adapted to operator surface (A): `model` is a ControlPixArtMSHalf whose control input is the degraded latent itself
(the authors' own usage, test_scripts/test_controlnet.py:137-139)."""
from __future__ import annotations

import torch

from . import _lib


class DDPMSchedulerLite:
    """The only piece of diffusers' DDPMScheduler the path reads: `alphas_cumprod` (fp32 cumprod of a linear beta
    schedule 1e-4..2e-2 over 1000 steps; in-tree twin: get_named_beta_schedule("linear"), gaussian_diffusion.py:99-116)."""

    def __init__(self, num_train_timesteps=1000, beta_start=1e-4, beta_end=2e-2):
        betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)


def eps_to_mu(scheduler, model_output, sample, timesteps):
    """generate.py:44-51 on the (B,8,H,W) model output: keeps channels [0,4) (generate.py:84-85) and solves for x0."""
    t = int(timesteps.reshape(-1)[0]) if torch.is_tensor(timesteps) else int(timesteps)  # tensor on GPU: one sync
    a = float(scheduler.alphas_cumprod.to(dtype=sample.dtype)[t])
    B, Cc = sample.shape[:2]
    hw = sample.shape[2] * sample.shape[3]
    s = sample.to(torch.float32).contiguous()
    mo = model_output.to(torch.float32).contiguous()
    if mo.shape[1] != 2 * Cc:
        raise ValueError("model output must carry the learned-sigma half (2*C channels)")
    out = torch.empty_like(s)
    with torch.cuda.device(s.device):
        _lib.check(_lib.lib().ir_eps_to_x0(s.data_ptr(), mo.data_ptr(), out.data_ptr(), B, Cc, hw, a ** 0.5,
                                          (1.0 - a) ** 0.5, _lib.stream_ptr()), "ir_eps_to_x0")
    return out


def _is_flavour_b(model) -> bool:
    cfg = getattr(model, "config", None)
    return cfg is not None and hasattr(cfg, "sample_size")


def forward_model(model, latents, timestep, prompt_embeds, prompt_attention_masks=None, c=None):
    """generate.py:54-87: micro-conditioning from the latent size, timestep expanded to the batch. Returns the full
    (B,8,h,w) output; eps_to_mu keeps channels [0,4) (generate.py:84-85).

    Flavour (A) models (ControlPixArtMSHalf) are called with the native signature; flavour (B) models
    (instarevive_b200.Transformer2DModel / ControlTransformerHalf, recognised by `config.sample_size` exactly as the
    reference does at :56) with the diffusers keywords of :66-82."""
    B, _, h, w = latents.shape
    ts = timestep.to(latents.device).float().expand(B)
    if _is_flavour_b(model):
        added_cond_kwargs = {"resolution": None, "aspect_ratio": None}
        if model.config.sample_size == 128:   # :56-62 -- note: latent height / width, as the reference passes them
            added_cond_kwargs = {
                "resolution": torch.tensor([float(h), float(w)], device=latents.device).repeat(B, 1),
                "aspect_ratio": torch.tensor([float(h / w)], device=latents.device).repeat(B, 1),
            }
        kw = dict(timestep=ts, encoder_hidden_states=prompt_embeds, encoder_attention_mask=prompt_attention_masks,
                  added_cond_kwargs=added_cond_kwargs)
        if c is None:
            return model(latents, **kw).sample
        return model(latents, c=c, **kw)
    return model(latents, ts, prompt_embeds, mask=prompt_attention_masks, data_info=_micro_conditions(B, h, w, latents.device), c=c)


_mc_cache: dict = {}


def _micro_conditions(B: int, h: int, w: int, device):
    """img_hw / aspect_ratio of generate.py:56-62 in pixels, cached per (batch, latent size, device): the reference
    builds (and uploads) them on every call."""
    key = (B, h, w, str(device))
    hit = _mc_cache.get(key)
    if hit is None:
        if len(_mc_cache) > 64:
            _mc_cache.clear()
        hit = _mc_cache[key] = {
            "img_hw": torch.tensor([[float(h * 8), float(w * 8)]], device=device).repeat(B, 1),
            "aspect_ratio": torch.tensor([[float(h) / float(w)]], device=device).repeat(B, 1),
        }
    return hit


def generate_sample_1step(model, scheduler, latents, maxt, prompt_embeds, prompt_attention_masks=None, c=None,
                          use_control: bool = False):
    """generate.py:22-42: one forward at t = maxt, eps -> x0. c=None runs the plain 28-block path, exactly as in the
    reference (pixart_controlnet.py:245-247; every reference caller passes no c). use_control=True (keyword-only
    extension, the north-star configuration) feeds the degraded latent itself to the ControlNet-Half branch: c = latents,
    the authors' own ControlNet usage (test_scripts/test_controlnet.py:137-139)."""
    t = torch.full((1,), maxt, device=latents.device).long()
    if c is None and use_control and getattr(model, "copy_blocks_num", 0) > 0:
        c = latents
    out = forward_model(model, latents, t, prompt_embeds, prompt_attention_masks, c=c)
    return eps_to_mu(scheduler, out, latents, int(maxt))
