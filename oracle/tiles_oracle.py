"""ORACLE (test infrastructure, not product code): CPU restatement of the reference's tiled-latent scheduler and
pixel post-processing: _sliding_windows / process() (test_scripts/inference.py:40-166) and wavelet_reconstruction /
adaptive_instance_normalization (utils/image/align_color.py:44-119).

Parity status: PINNED. oracle/make_goldens.py lifts `_sliding_windows` and `process` from the reference file with `ast`
(unmodified source text), runs them with injected one-step-generator / VAE adapters and stores window tables, count
masks and uint8 images under tests/golden/; tests/test_oracle.py checks this restatement against them.
Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""
from __future__ import annotations

from typing import Callable, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F


def sliding_windows(h: int, w: int, tile_size: int, tile_stride: int) -> List[Tuple[int, int, int, int]]:
    """_sliding_windows, test_scripts/inference.py:40-53 (integer, bit-exact): row-major (hi, hi_end, wi, wi_end)."""
    hi_list = list(range(0, h - tile_size + 1, tile_stride))
    if (h - tile_size) % tile_stride != 0:
        hi_list.append(h - tile_size)
    wi_list = list(range(0, w - tile_size + 1, tile_stride))
    if (w - tile_size) % tile_stride != 0:
        wi_list.append(w - tile_size)
    return [(hi, hi + tile_size, wi, wi + tile_size) for hi in hi_list for wi in wi_list]


def count_mask(h: int, w: int, coords) -> np.ndarray:
    """How many tiles cover each latent pixel (the `count` buffer of inference.py:124,134)."""
    cnt = np.zeros((h, w), dtype=np.int64)
    for hi, hi_end, wi, wi_end in coords:
        cnt[hi:hi_end, wi:wi_end] += 1
    return cnt


# ------------------------------------------------------------------------------------------------ colour fix
def wavelet_blur(image: torch.Tensor, radius: int) -> torch.Tensor:
    """wavelet_blur, align_color.py:73-92: replicate pad by `radius`, depthwise 3x3 [1 2 1]^2/16 with dilation."""
    k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]],
                     dtype=image.dtype)[None, None].repeat(3, 1, 1, 1)
    image = F.pad(image, (radius, radius, radius, radius), mode="replicate")
    return F.conv2d(image, k, groups=3, dilation=radius)


def wavelet_decomposition(image: torch.Tensor, levels: int = 5):
    """wavelet_decomposition, align_color.py:94-106."""
    high = torch.zeros_like(image)
    low = image
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high = high + (image - low)
        image = low
    return high, low


def wavelet_reconstruction(content: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
    """wavelet_reconstruction, align_color.py:108-119: high(content) + low(style)."""
    ch, _ = wavelet_decomposition(content)
    _, sl = wavelet_decomposition(style)
    return ch + sl


def adaptive_instance_normalization(content: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
    """adaptive_instance_normalization, align_color.py:44-71 (unbiased variance + 1e-5)."""
    def ms(f):
        b, c = f.shape[:2]
        var = f.reshape(b, c, -1).var(dim=2) + 1e-5
        return f.reshape(b, c, -1).mean(dim=2).reshape(b, c, 1, 1), var.sqrt().reshape(b, c, 1, 1)
    sm, ss = ms(style)
    cm, cs = ms(content)
    return (content - cm) / cs * ss + sm


# ------------------------------------------------------------------------------------------------ process()
@torch.no_grad()
def process(control: torch.Tensor, init_noise: torch.Tensor, one_step: Callable, decode: Callable,
            scaling_factor: float, tiled: bool, tile_size: int = 512, tile_stride: int = 448,
            color_fix_type: str = "wavelet"):
    """The part of process() after VAE-encode (test_scripts/inference.py:111-163).

    control: (N,3,H,W) in [0,1]; init_noise: (N,4,H/8,W/8) = c_latent * scaling_factor;
    one_step(latents) -> x0 latents (generate_sample_1step); decode(z) -> image in ~[-1,1] (vae.decode(...).sample).
    Returns (uint8 NHWC images, blended latent buffer or x0 latents).
    """
    n, _, height, width = control.shape
    h, w = height // 8, width // 8
    if not tiled:
        latents = one_step(init_noise)                                              # :114
        img = decode(latents / scaling_factor) / 2 + 0.5                            # :116-117
        lat_out = latents
    else:
        coords = sliding_windows(h, w, tile_size // 8, tile_stride // 8)            # :120
        count = torch.zeros((n, 4, h, w), dtype=torch.long).to(init_noise)          # :124 (float after .to)
        noise_buffer = torch.zeros_like(init_noise)                                 # :126
        for hi, hi_end, wi, wi_end in coords:                                       # :128-134
            tile = one_step(init_noise[:, :, hi:hi_end, wi:wi_end])
            noise_buffer[:, :, hi:hi_end, wi:wi_end] += tile
            count[:, :, hi:hi_end, wi:wi_end] += 1
        noise_buffer.div_(count)                                                    # :136
        img = torch.zeros_like(control)
        pcount = torch.zeros_like(control, dtype=torch.long)                        # :137
        for hi, hi_end, wi, wi_end in coords:                                       # :139-152
            tile_img = decode(noise_buffer[:, :, hi:hi_end, wi:wi_end] / scaling_factor) / 2 + 0.5
            cond = control[:, :, hi * 8:hi_end * 8, wi * 8:wi_end * 8]
            if color_fix_type == "adain":
                tile_img = adaptive_instance_normalization(tile_img, cond)
            elif color_fix_type == "wavelet":
                tile_img = wavelet_reconstruction(tile_img, cond)
            img[:, :, hi * 8:hi_end * 8, wi * 8:wi_end * 8] += tile_img
            pcount[:, :, hi * 8:hi_end * 8, wi * 8:wi_end * 8] += 1
        img.div_(pcount)                                                            # :153
        lat_out = noise_buffer
    x = img.clamp(0, 1)                                                             # :159-160
    x = (x.permute(0, 2, 3, 1) * 255).cpu().numpy().clip(0, 255).astype(np.uint8)
    return x, lat_out
