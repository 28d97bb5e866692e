"""Host-side mirror of the VAE surface that process() uses (test_scripts/inference.py:104-117,142):
`vae.config.scaling_factor`, `vae.decode(z)` (diffusers style: object with `.sample`; ldm style via `decode_tensor`)
and `vae.encode(x).latent_dist.mode()`. The decoder (`post_quant_conv` + ldm `Decoder`, ldm/models/autoencoder.py:88-91,
ldm/modules/diffusionmodules/model.py:549-655) runs in libinstarevive_b200.so.

`AutoencoderKLDecoder` holds the decoder only (the hot path; an encoder callable can be injected);
`AutoencoderKL` also holds `encoder.*` / `quant_conv.*` and runs AutoencoderKL.encode (autoencoder.py:82-86,
Encoder.forward model.py:521-546) on the same CUDA kernels -- SURVEY 8f row 1, the step right before the hot path."""
from __future__ import annotations

import ctypes as C
from types import SimpleNamespace
from typing import Callable, Mapping, Optional

import torch

from . import _lib


class DecoderOutput:
    def __init__(self, sample):
        self.sample = sample


class DiagonalGaussianDistribution:
    """ldm/modules/distributions/distributions.py:24-62 (the members process() and diffusers callers use)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise


class EncoderOutput:
    def __init__(self, latent_dist):
        self.latent_dist = latent_dist


class AutoencoderKLDecoder:
    """`post_quant_conv` + `decoder.*` weights (reference key names) on a CUDA device, decode through the C ABI."""

    _WITH_ENCODER = False

    def __init__(self, state_dict: Mapping[str, torch.Tensor], device="cuda", scaling_factor: float = 0.18215,
                 ch: int = 128, ch_mult=(1, 2, 4, 4), num_res_blocks: int = 2, z_channels: int = 4, out_ch: int = 3,
                 encoder: Optional[Callable] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: the VAE decoder needs a CUDA device")
        self.config = SimpleNamespace(scaling_factor=scaling_factor)
        self._encoder = encoder
        self._ws = None
        L = _lib.lib()
        self._ws_enc = None
        cfg = _lib.VaeConfig(ch, z_channels, out_ch, num_res_blocks, (C.c_int * 4)(*ch_mult), 1 if self._WITH_ENCODER else 0)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.ir_vae_create(C.byref(cfg), C.byref(h)), "ir_vae_create")
            self._handle = h.value
            name = C.create_string_buffer(256)
            numel = C.c_longlong()
            for i in range(L.ir_vae_num_params(self._handle)):
                _lib.check(L.ir_vae_param_info(self._handle, i, name, 256, C.byref(numel)), "ir_vae_param_info")
                key = name.value.decode()
                if key not in state_dict:
                    raise KeyError(f"VAE state_dict lacks '{key}'")
                t = state_dict[key].detach().to(device=self.device, dtype=torch.float32).contiguous()
                if t.numel() != numel.value:
                    raise ValueError(f"{key}: {t.numel()} elements, library expects {numel.value}")
                _lib.check(L.ir_vae_load_param(self._handle, key.encode(), t.data_ptr(), t.numel(), _lib.stream_ptr()),
                           f"load {key}")
            torch.cuda.current_stream().synchronize()

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ir_vae_destroy(self._handle)
        except Exception:
            pass

    @torch.no_grad()
    def set_cuda_graphs(self, enable: bool) -> None:
        """CUDA-graph replay of decode / encode calls (default on: from the third call with a given batch / size the ~110
        launches of a call are replayed as one graph). Off = plain stream-ordered launches."""
        _lib.check(_lib.lib().ir_vae_set_graphs(self._handle, 1 if enable else 0), "ir_vae_set_graphs")

    def decode_tensor(self, z: torch.Tensor, in_scale: float = 1.0, out_scale: float = 1.0, out_shift: float = 0.0):
        """(B,4,h,w) latents -> (B,3,8h,8w) fp32 image = Decoder(post_quant_conv(z*in_scale))*out_scale + out_shift."""
        if z.device.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: latents must be CUDA tensors")
        L = _lib.lib()
        zz = z.to(dtype=torch.float32).contiguous()
        B, _, h, w = zz.shape
        out = torch.empty(B, 3, 8 * h, 8 * w, device=z.device, dtype=torch.float32)
        with torch.cuda.device(z.device):
            need = L.ir_vae_workspace_bytes(self._handle, B, h, w)
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=z.device)
            _lib.check(L.ir_vae_decode(self._handle, zz.data_ptr(), out.data_ptr(), B, h, w, in_scale, out_scale,
                                       out_shift, self._ws.data_ptr(), self._ws.numel(), _lib.stream_ptr()),
                       "ir_vae_decode")
        return out

    def decode(self, z: torch.Tensor) -> DecoderOutput:
        return DecoderOutput(self.decode_tensor(z))

    def encode(self, x: torch.Tensor):
        if self._encoder is None:
            raise NotImplementedError("the VAE encoder is outside the restoration hot path (SURVEY 8f); pass encoder=")
        return self._encoder(x)


class AutoencoderKL(AutoencoderKLDecoder):
    """Decoder + encoder (`encoder.*`, `quant_conv.*`) on the device; `encode(x).latent_dist.mode()` as process() calls
    it (test_scripts/inference.py:104-109). x: (B,3,H,W) in [-1,1], H and W multiples of 16."""

    _WITH_ENCODER = True

    @torch.no_grad()
    def encode_moments(self, x: torch.Tensor) -> torch.Tensor:
        if x.device.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: images must be CUDA tensors")
        L = _lib.lib()
        xx = x.to(dtype=torch.float32).contiguous()
        B, ch, H, W = xx.shape
        if ch != 3 or H % 16 or W % 16:
            raise ValueError(f"encode expects (B,3,H,W) with H, W multiples of 16, got {tuple(xx.shape)}")
        out = torch.empty(B, 8, H // 8, W // 8, device=x.device, dtype=torch.float32)
        with torch.cuda.device(x.device):
            need = L.ir_vae_encode_workspace_bytes(self._handle, B, H, W)
            if self._ws_enc is None or self._ws_enc.numel() < need:
                self._ws_enc = None
                self._ws_enc = torch.empty(need, dtype=torch.uint8, device=x.device)
            _lib.check(L.ir_vae_encode(self._handle, xx.data_ptr(), out.data_ptr(), B, H, W, self._ws_enc.data_ptr(),
                                       self._ws_enc.numel(), _lib.stream_ptr()), "ir_vae_encode")
        return out

    def encode(self, x: torch.Tensor) -> EncoderOutput:
        return EncoderOutput(DiagonalGaussianDistribution(self.encode_moments(x)))
