"""Stage-1 SwinIR (`preprocess_model` of test_scripts/inference.py:92-103,245-248) behind the reference's call shape:
`SwinIR(state_dict)(x)` with x (B,3,H,W) in [0,1] -> (B,3,H,W). The arithmetic is diffusion/model/swinir.py:867-905 with
the parameters of configs/swinir.yaml, run by libinstarevive_b200.so (csrc/swinir.cu) -- SURVEY 8f row 2."""
from __future__ import annotations

import ctypes as C
from typing import Mapping

import torch

from . import _lib


class SwinIR:
    """Weights under the reference's parameter names (buffers such as relative_position_index / attn_mask and the lpips
    metric of the training module are ignored) on a CUDA device."""

    def __init__(self, state_dict: Mapping[str, torch.Tensor], device="cuda"):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: SwinIR needs a CUDA device")
        self._ws = None
        L = _lib.lib()
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(L.ir_swinir_create(C.byref(h)), "ir_swinir_create")
            self._handle = h.value
            name = C.create_string_buffer(256)
            numel = C.c_longlong()
            for i in range(L.ir_swinir_num_params(self._handle)):
                _lib.check(L.ir_swinir_param_info(self._handle, i, name, 256, C.byref(numel)), "ir_swinir_param_info")
                key = name.value.decode()
                if key not in state_dict:
                    raise KeyError(f"SwinIR state_dict lacks '{key}'")
                t = state_dict[key].detach().to(device=self.device, dtype=torch.float32).contiguous()
                if t.numel() != numel.value:
                    raise ValueError(f"{key}: {t.numel()} elements, library expects {numel.value}")
                _lib.check(L.ir_swinir_load_param(self._handle, key.encode(), t.data_ptr(), t.numel(), _lib.stream_ptr()),
                           f"load {key}")
            torch.cuda.current_stream().synchronize()

    def __del__(self):
        try:
            if getattr(self, "_handle", None):
                _lib.lib().ir_swinir_destroy(self._handle)
        except Exception:
            pass

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if x.device.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: images must be CUDA tensors")
        xx = x.to(dtype=torch.float32).contiguous()
        B, ch, H, W = xx.shape
        if ch != 3 or H % 64 or W % 64:
            raise ValueError(f"SwinIR expects (B,3,H,W) with H, W multiples of 64, got {tuple(xx.shape)}")
        L = _lib.lib()
        out = torch.empty_like(xx)
        with torch.cuda.device(x.device):
            need = L.ir_swinir_workspace_bytes(self._handle, B, H, W)
            if self._ws is None or self._ws.numel() < need:
                self._ws = None
                self._ws = torch.empty(need, dtype=torch.uint8, device=x.device)
            _lib.check(L.ir_swinir_forward(self._handle, xx.data_ptr(), out.data_ptr(), B, H, W, self._ws.data_ptr(),
                                           self._ws.numel(), _lib.stream_ptr()), "ir_swinir_forward")
        return out

    __call__ = forward
