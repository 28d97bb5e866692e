"""ctypes binding of libinstarevive_b200.so (the C ABI declared in include/instarevive_b200.h).

The product path has no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libinstarevive_b200.so"

_vp, _i, _ll, _f, _sz = C.c_void_p, C.c_int, C.c_longlong, C.c_float, C.c_size_t


class DitConfig(C.Structure):
    _fields_ = [
        ("depth", _i), ("copy_blocks", _i), ("hidden", _i), ("heads", _i), ("patch", _i),
        ("in_channels", _i), ("out_channels", _i), ("caption_channels", _i), ("mlp_ratio", _i),
        ("base_size", _i), ("pe_interpolation", _f),
    ]


class VaeConfig(C.Structure):
    _fields_ = [("ch", _i), ("z_channels", _i), ("out_ch", _i), ("num_res_blocks", _i), ("ch_mult", _i * 4),
                ("with_encoder", _i)]


# name -> (restype, argtypes); every symbol of include/instarevive_b200.h must appear here
PROTOTYPES = {
    "ir_last_error": (C.c_char_p, []),
    "ir_version": (C.c_char_p, []),
    "ir_launch_count": (_ll, []),
    "ir_profile_begin": (None, []),
    "ir_profile_end": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(_ll)]),
    "ir_profile_records": (_ll, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_i), C.POINTER(_f), _ll]),
    "ir_profile_calibrate": (_i, [_i, _vp]),
    "ir_dit_create": (_i, [C.POINTER(DitConfig), C.POINTER(_vp)]),
    "ir_dit_destroy": (None, [_vp]),
    "ir_dit_num_params": (_i, [_vp]),
    "ir_dit_param_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_ll), C.POINTER(_i), C.POINTER(_i)]),
    "ir_dit_load_param": (_i, [_vp, C.c_char_p, _vp, _ll, _vp]),
    "ir_dit_workspace_bytes": (_sz, [_vp, _i, _i, _i, _i]),
    "ir_dit_reserve": (_i, [_vp, _i, _i]),
    "ir_dit_set_graphs": (_i, [_vp, _i]),
    "ir_dit_set_dual_chain": (_i, [_vp, _i]),
    "ir_vae_set_graphs": (_i, [_vp, _i]),
    "ir_dit_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _ll, _i, _vp, _sz, _vp]),
    "ir_dit_patch_embed": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ir_eps_to_x0": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _vp]),
    "ir_gemm_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _ll, _ll, _ll, _i, _f, _vp, _vp, _vp, _vp, _ll, _i, _i, _vp]),
    "ir_conv3x3_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp]),
    "ir_upsample_conv3x3_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "ir_conv3x3_s2_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "ir_conv1x1_bf16": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _vp]),
    "ir_gemm_attn_pass": (_i, [_vp, _vp, _i, _i, _i, _ll, _ll, _i, _f, _vp, _vp, _vp, _ll, _i, _vp]),
    "ir_attention_bf16": (_i, [_vp, _vp, _vp, _vp, _ll, _ll, _ll, _ll, _i, _i, _i, _i, _i, _vp, _vp, _f, _vp]),
    "ir_gemm_qkv_heads": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "ir_attention_tc_bf16": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _i, _i, _f, _vp]),
    "ir_cross_attention_vt_bytes": (_sz, [_i, _i]),
    "ir_cross_attention_tc_bf16": (_i, [_vp, _vp, _vp, _vp, _ll, _ll, _ll, _i, _i, _i, _i, _i, _vp, _vp, _i, _f, _vp]),
    "ir_debug_attention_trace": (_i, [_vp]),
    "ir_debug_gemm_trace": (_i, [_vp, _i]),
    "ir_ln_modulate": (_i, [_vp, _vp, _vp, _vp, _ll, _i, _i, _i, _vp]),
    "ir_pos_embed": (_i, [_vp, _i, _i, _i, _i, _f, _vp]),
    "ir_lincomb3": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _vp]),
    "ir_swinir_create": (_i, [C.POINTER(_vp)]),
    "ir_swinir_destroy": (None, [_vp]),
    "ir_swinir_num_params": (_i, [_vp]),
    "ir_swinir_param_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_ll)]),
    "ir_swinir_load_param": (_i, [_vp, C.c_char_p, _vp, _ll, _vp]),
    "ir_swinir_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "ir_swinir_forward": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "ir_vae_create": (_i, [C.POINTER(VaeConfig), C.POINTER(_vp)]),
    "ir_vae_destroy": (None, [_vp]),
    "ir_vae_num_params": (_i, [_vp]),
    "ir_vae_param_info": (_i, [_vp, _i, C.c_char_p, _i, C.POINTER(_ll)]),
    "ir_vae_load_param": (_i, [_vp, C.c_char_p, _vp, _ll, _vp]),
    "ir_vae_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "ir_vae_decode": (_i, [_vp, _vp, _vp, _i, _i, _i, _f, _f, _f, _vp, _sz, _vp]),
    "ir_vae_encode_workspace_bytes": (_sz, [_vp, _i, _i, _i]),
    "ir_vae_encode": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _sz, _vp]),
    "ir_tile_gather": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ir_tile_blend": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ir_wavelet_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "ir_wavelet_reconstruction": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _sz, _vp]),
    "ir_adain": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ir_to_uint8": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
}

_lib = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library with typed prototypes. Fails loudly if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m instarevive_b200.csrc.build` "
            "(there is no CPU or PyTorch fallback for the restoration path)")
    l = C.CDLL(str(LIB_PATH))
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(l, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = l
    return l


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().ir_last_error().decode(errors="replace")
        raise RuntimeError(f"instarevive_b200 {what} failed (status {status}): {msg}")


def ptr(t) -> int:
    """Device (or host) address of a torch tensor, None -> NULL."""
    return None if t is None else t.data_ptr()


def stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(lib().ir_launch_count())
