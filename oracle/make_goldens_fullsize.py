"""Mint the goldens at the sizes the benchmark is quoted on ("pin parity at the benchmark's own sizes") by
EXECUTING THE UNMODIFIED REFERENCE on the CPU of the build container. Run:  python oracle/make_goldens_fullsize.py [case ...]

Cases (each a few minutes of CPU time, fp32):
  dit128   full 28+13-block ControlPixArtMSHalf.forward at a 128x128 latent (T = 4096, L = 77) + x0 via eps_to_mu
           (diffusion/model/nets/pixart_controlnet.py:191-251; BASELINE configs[1])
  vae128   Decoder at a 128x128 latent, mid-attention over P = 16384 tokens
           (ldm/modules/diffusionmodules/model.py:622-655, :181-205)
  tiled2048  the reference's own process() (test_scripts/inference.py:56-166, lifted with ast) on a 2048x2048 image,
           25 tiles of 512/448, wavelet colour fix, reduced-depth generator (BASELINE configs[3])
  b8       batch-8 forward at a 128x128 latent with ragged captions, reduced depth (BASELINE configs[2]: M = 32768 rows)

The outputs are too large to commit whole (12 MB per fp32 1024^2 image), so each file stores fp32 crops (including
crops that straddle kernel-tile and restoration-tile seams), 8x8 block means of the whole tensor and, for images, the
uint8 picture or strips of it; tests/test_gpu_fullsize.py compares the CUDA path with exactly these.
"""
from __future__ import annotations

import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent))
import make_goldens as mg  # noqa: E402  (sets up the shims and imports the reference modules)
from instarevive_b200 import weights  # noqa: E402

GOLD = mg.GOLD
torch.set_grad_enabled(False)

# crops of a 1024x1024 decoded image: (y0, x0), 128x128 each. (448,448) straddles the centre; (0,0) / (896,896) are
# corners (conv zero padding); (192,704) and (640,64) straddle the 16x8-pixel conv tiles and 128/256-row M tiles
VAE_CROPS = ((0, 0), (448, 448), (896, 896), (192, 704), (640, 64), (896, 320))


def case_dit128():
    net = mg.build_dit(28, 13, seed=1)
    x, ts, y, mask, info = weights.make_inputs(1, 128, 128, seed=0, lens=(77,))
    t0 = time.time()
    out = net(x, ts, y, mask=mask, data_info=info, c=x.clone())
    dt = time.time() - t0
    betas = mg.get_named_beta_schedule("linear", 1000)
    sched = types.SimpleNamespace(alphas_cumprod=torch.from_numpy(np.cumprod(1.0 - betas, axis=0)))
    x0 = mg.eps_to_mu(sched, out.chunk(2, dim=1)[0], x, torch.full((1,), 400).long())
    np.savez_compressed(GOLD / "dit_full_b1_128x128.npz", out=out.numpy(), x0=x0.numpy(), depth=28, copy_blocks=13, wseed=1,
                        B=1, h=128, w=128, lens=np.array((77,)), iseed=0, use_mask=True, use_c=True)
    print(f"dit_full_b1_128x128: out std {out.std():.4f} absmax {out.abs().max():.3f} x0 std {x0.std():.4f} ({dt:.1f}s)", flush=True)


def case_b8():
    net = mg.build_dit(4, 2, seed=11)
    lens = (77, 120, 33, 1, 100, 64, 120, 5)
    x, ts, y, mask, info = weights.make_inputs(8, 128, 128, seed=6, lens=lens)
    t0 = time.time()
    out = net(x, ts, y, mask=mask, data_info=info, c=x.clone())
    dt = time.time() - t0
    np.savez_compressed(GOLD / "dit_small_b8_128x128_ragged.npz", eps=out[:, :4].numpy(), sigma_means=mg.block_means(out[:, 4:], 8).numpy(),
                        sigma_s7=out[7, 4:].numpy(), depth=4, copy_blocks=2, wseed=11, B=8, h=128, w=128, lens=np.array(lens), iseed=6,
                        use_mask=True, use_c=True)
    print(f"dit_small_b8_128x128_ragged: out std {out.std():.4f} absmax {out.abs().max():.3f} ({dt:.1f}s)", flush=True)


def case_vae128():
    decode = mg.build_decoder(seed=2)
    z = torch.randn(1, 4, 128, 128, generator=torch.Generator().manual_seed(8)) / 0.18215 * 0.6
    t0 = time.time()
    img = decode(z)
    dt = time.time() - t0
    u8 = ((img / 2 + 0.5).clamp(0, 1).permute(0, 2, 3, 1) * 255).numpy().clip(0, 255).astype(np.uint8)
    crops = np.stack([img[0, :, y0:y0 + 128, x0:x0 + 128].numpy() for (y0, x0) in VAE_CROPS])
    np.savez_compressed(GOLD / "vae_b1_128x128.npz", means=mg.block_means(img).numpy(), crops=crops, crop_origins=np.array(VAE_CROPS),
                        u8=u8[0], wseed=2, zseed=8, mean=img.mean().item(), std=img.std().item())
    print(f"vae_b1_128x128: mean {img.mean():.3f} std {img.std():.3f} min {img.min():.3f} max {img.max():.3f} ({dt:.1f}s)", flush=True)


def case_tiled2048():
    import ast
    import einops
    from typing import List, Tuple
    from tqdm import tqdm
    src = (mg.REF / "test_scripts" / "inference.py").read_text()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("_sliding_windows", "process")]
    assert len(fns) == 2
    ns = dict(torch=torch, np=np, einops=einops, tqdm=tqdm, List=List, Tuple=Tuple,
              wavelet_reconstruction=mg.align_color.wavelet_reconstruction,
              adaptive_instance_normalization=mg.align_color.adaptive_instance_normalization)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "<reference test_scripts/inference.py>", "exec"), ns)
    betas = mg.get_named_beta_schedule("linear", 1000)
    sched = types.SimpleNamespace(alphas_cumprod=torch.from_numpy(np.cumprod(1.0 - betas, axis=0)))
    tiny = mg.build_dit(2, 1, seed=21)
    tiny.device = torch.device("cpu")

    def gen_1step(model, scheduler, latents, maxt, y, y_mask, c=None):
        # adapter of generate_sample_1step (generate.py:22-42) to operator surface (A); x = c = degraded latent
        B, _, hh, ww = latents.shape
        info = {"img_hw": torch.tensor([[hh * 8.0, ww * 8.0]] * B), "aspect_ratio": torch.tensor([[hh / ww]] * B)}
        t = torch.full((B,), float(maxt))
        out = model(latents, t, y, mask=y_mask, data_info=info, c=latents)
        return mg.eps_to_mu(scheduler, out.chunk(2, dim=1)[0], latents, torch.full((1,), maxt).long())

    ns["generate_sample_1step"] = gen_1step
    ns["noise_scheduler"] = sched
    vae = weights.SyntheticVAE(mg.build_decoder(seed=2))
    _, _, y, mask, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))
    H = W = 2048
    img = weights.synthetic_degraded_image(H, W, seed=0)   # the image bench.py --workload tiled restores
    assert len(ns["_sliding_windows"](H // 8, W // 8, 64, 56)) == 25
    t0 = time.time()
    preds, _ = ns["process"](tiny, [img], strength=1, color_fix_type="wavelet", disable_preprocess_model=True, tiled=True,
                             tile_size=512, tile_stride=448, vae=vae, y=y, y_mask=mask)
    pred = preds[0]
    dt = time.time() - t0
    means = pred.astype(np.float32).reshape(H // 8, 8, W // 8, 8, 3).mean(axis=(1, 3))
    crops_at = ((0, 0), (960, 960), (1400, 300), (1900, 1900), (440, 1500), (1330, 880))   # 128x128, several on tile seams
    crops = np.stack([pred[y0:y0 + 128, x0:x0 + 128] for (y0, x0) in crops_at])
    np.savez_compressed(GOLD / "process_tiled_2048x2048_wavelet.npz", means=means, row_strip=pred[440:520], row_strip_y=440,
                        col_strip=pred[:, 1530:1610], col_strip_x=1530, crops=crops, crop_origins=np.array(crops_at),
                        H=H, W=W, tiled=True, fix="wavelet", dit_seed=21, vae_seed=2, img_seed=0, cap_seed=9,
                        mean=float(pred.mean()), std=float(pred.std()))
    print(f"process_tiled_2048x2048_wavelet: mean {pred.mean():.1f} std {pred.std():.1f} ({dt:.1f}s)", flush=True)


CASES = {"dit128": case_dit128, "vae128": case_vae128, "tiled2048": case_tiled2048, "b8": case_b8}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(CASES)):
        CASES[name]()
