"""Mint tests/golden/*.npz by EXECUTING THE UNMODIFIED REFERENCE in the build container (/root/reference is read-only and
does not exist on the GPU box, so the vectors are committed). Run:  python oracle/make_goldens.py

What runs: the reference's own ControlPixArtMSHalf / PixArtMS (diffusion/model/nets), Decoder
(ldm/modules/diffusionmodules/model.py), eps_to_mu (scripts/DMD/transformer_train/generate.py),
get_named_beta_schedule (diffusion/model/gaussian_diffusion.py), wavelet_reconstruction (utils/image/align_color.py) and
`_sliding_windows` / `process` lifted with `ast` from test_scripts/inference.py -- third-party imports the container
lacks are satisfied by oracle/shims/. Weights come from instarevive_b200/weights.py (seeded, recorded in each file).
"""
from __future__ import annotations

import ast
import importlib.util
import sys
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle" / "shims"))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT))
for name in ("diffusion",):  # bare package: skip diffusion/__init__.py (pulls samplers that are out of scope)
    pkg = types.ModuleType(name)
    pkg.__path__ = [str(REF / name)]
    sys.modules[name] = pkg

from diffusion.model.nets.PixArtMS import PixArtMS  # noqa: E402
from diffusion.model.nets.pixart_controlnet import ControlPixArtMSHalf  # noqa: E402
from diffusion.model.gaussian_diffusion import get_named_beta_schedule  # noqa: E402
import ldm.xformers_state as _xs  # noqa: E402
_xs.disable_xformers()  # use the vanilla AttnBlock (model.py:154-205), the a24 row of SURVEY section 8
from ldm.modules.diffusionmodules.model import Decoder  # noqa: E402
from scripts.DMD.transformer_train.generate import eps_to_mu  # noqa: E402

from instarevive_b200 import weights  # noqa: E402

GOLD = ROOT / "tests" / "golden"
GOLD.mkdir(parents=True, exist_ok=True)
torch.set_grad_enabled(False)


def load_file_module(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


align_color = load_file_module("ref_align_color", REF / "utils" / "image" / "align_color.py")


def build_dit(depth, copy_blocks, seed, input_size=64):
    base = PixArtMS(depth=depth, hidden_size=1152, patch_size=2, num_heads=16, input_size=input_size,
                    micro_condition=True, model_max_length=120)
    net = ControlPixArtMSHalf(base, copy_blocks_num=copy_blocks).eval()
    sd = weights.make_dit_state_dict(depth=depth, copy_blocks=copy_blocks, seed=seed, input_size=input_size)
    missing = net.load_state_dict(sd, strict=True)  # strict: the factory's keys/shapes ARE the reference's
    assert not missing.missing_keys and not missing.unexpected_keys
    return net


def build_decoder(seed):
    dec = Decoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                  in_channels=3, resolution=256, z_channels=4, double_z=True).eval()
    pq = torch.nn.Conv2d(4, 4, 1).eval()
    sd = weights.make_vae_decoder_state_dict(seed=seed)
    dec.load_state_dict({k[len("decoder."):]: v for k, v in sd.items() if k.startswith("decoder.")}, strict=True)
    pq.load_state_dict({k[len("post_quant_conv."):]: v for k, v in sd.items() if k.startswith("post_quant_conv.")})
    return lambda z: dec(pq(z))  # AutoencoderKL.decode, ldm/models/autoencoder.py:88-91


def dit_case(tag, net, depth, copy_blocks, wseed, B, h, w, lens, iseed, use_mask=True, use_c=True):
    x, ts, y, mask, info = weights.make_inputs(B, h, w, seed=iseed, lens=lens)
    t0 = time.time()
    out = net(x, ts, y, mask=mask if use_mask else None, data_info=info, c=x.clone() if use_c else None)
    dt = time.time() - t0
    np.savez_compressed(GOLD / f"dit_{tag}.npz", out=out.numpy(), depth=depth, copy_blocks=copy_blocks, wseed=wseed,
                        B=B, h=h, w=w, lens=np.array(lens), iseed=iseed, use_mask=use_mask, use_c=use_c)
    print(f"dit_{tag}: out {tuple(out.shape)} std {out.std():.4f} absmax {out.abs().max():.3f}  ({dt:.1f}s)", flush=True)


def block_means(img, k=8):
    b, c, h, w = img.shape
    return img.reshape(b, c, h // k, k, w // k, k).mean(dim=(3, 5))


def main():
    # ------------------------------------------------------------------ known answers
    betas = get_named_beta_schedule("linear", 1000)
    abar = np.cumprod(1.0 - betas, axis=0)
    np.savez(GOLD / "alphas.npz", abar400=abar[400], abar=abar)
    print("abar[400] =", repr(abar[400]))

    # ------------------------------------------------------------------ DiT + ControlNet forward
    small = build_dit(4, 2, seed=11)
    dit_case("small_b1_64x64", small, 4, 2, 11, 1, 64, 64, (77,), 0)
    dit_case("small_b2_64x96_ragged", small, 4, 2, 11, 2, 64, 96, (120, 33), 1)
    dit_case("small_b1_32x32_nomask", small, 4, 2, 11, 1, 32, 32, (120,), 2, use_mask=False)
    dit_case("small_b1_64x64_noc", small, 4, 2, 11, 1, 64, 64, (77,), 3, use_c=False)
    dit_case("small_b1_40x72", small, 4, 2, 11, 1, 40, 72, (50,), 4)
    del small
    full = build_dit(28, 13, seed=1)
    dit_case("full_b1_64x64", full, 28, 13, 1, 1, 64, 64, (77,), 0)
    # eps -> x0 with the reference's eps_to_mu and its in-tree schedule
    x, ts, y, mask, info = weights.make_inputs(1, 64, 64, seed=0, lens=(77,))
    out = full(x, ts, y, mask=mask, data_info=info, c=x.clone())
    sched = types.SimpleNamespace(alphas_cumprod=torch.from_numpy(abar))
    x0 = eps_to_mu(sched, out.chunk(2, dim=1)[0], x, torch.full((1,), 400).long())
    np.savez_compressed(GOLD / "x0_full_b1_64x64.npz", x0=x0.numpy())
    print("x0 std", x0.std().item())
    del full

    # ------------------------------------------------------------------ VAE decoder
    decode = build_decoder(seed=2)
    for tag, (B, h, w), seed in (("b1_32x32", (1, 32, 32), 5), ("b2_16x24", (2, 16, 24), 6)):
        z = torch.randn(B, 4, h, w, generator=torch.Generator().manual_seed(seed)) / 0.18215 * 0.6
        t0 = time.time()
        img = decode(z)
        np.savez_compressed(GOLD / f"vae_{tag}.npz", img=img.numpy(), wseed=2, zseed=seed, B=B, h=h, w=w)
        print(f"vae_{tag}: {tuple(img.shape)} mean {img.mean():.3f} std {img.std():.3f} min {img.min():.3f} max {img.max():.3f} ({time.time() - t0:.1f}s)", flush=True)
    z = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(7)) / 0.18215 * 0.6
    img = decode(z)
    np.savez_compressed(GOLD / "vae_b1_64x64_blockmeans.npz", means=block_means(img).numpy(), crop=img[:, :, 224:288, 224:288].numpy(),
                        wseed=2, zseed=7, mean=img.mean().item(), std=img.std().item())
    print(f"vae_b1_64x64: std {img.std():.3f}", flush=True)

    # ------------------------------------------------------------------ tile scheduler: tables and masks
    src = (REF / "test_scripts" / "inference.py").read_text()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("_sliding_windows", "process")]
    assert len(fns) == 2
    import einops
    from typing import List, Tuple
    from tqdm import tqdm
    ns = dict(torch=torch, np=np, einops=einops, tqdm=tqdm, List=List, Tuple=Tuple,
              wavelet_reconstruction=align_color.wavelet_reconstruction,
              adaptive_instance_normalization=align_color.adaptive_instance_normalization)
    exec(compile(ast.Module(body=fns, type_ignores=[]), "<reference test_scripts/inference.py>", "exec"), ns)
    tables = {}
    for h in (64, 72, 104, 120, 128, 184, 256):
        for w in (64, 96, 128, 200, 256):
            coords = ns["_sliding_windows"](h, w, 64, 56)
            tables[f"{h}x{w}"] = np.array(coords, dtype=np.int64)
    for (h, w, t, s) in ((40, 56, 32, 24), (32, 32, 32, 24), (100, 36, 32, 32)):
        tables[f"{h}x{w}_t{t}_s{s}"] = np.array(ns["_sliding_windows"](h, w, t, s), dtype=np.int64)
    np.savez_compressed(GOLD / "sliding_windows.npz", **tables)
    print("sliding windows:", {k: len(v) for k, v in list(tables.items())[:6]}, "...")

    # ------------------------------------------------------------------ process(): end to end through the reference loop
    tiny = build_dit(2, 1, seed=21)
    tiny.device = torch.device("cpu")

    def gen_1step(model, scheduler, latents, maxt, y, y_mask, c=None):
        # adapter of generate_sample_1step (generate.py:22-42) to operator surface (A); x = c = degraded latent
        B, _, hh, ww = latents.shape
        info = {"img_hw": torch.tensor([[hh * 8.0, ww * 8.0]] * B), "aspect_ratio": torch.tensor([[hh / ww]] * B)}
        t = torch.full((B,), float(maxt))
        out = model(latents, t, y, mask=y_mask, data_info=info, c=latents)
        return eps_to_mu(scheduler, out.chunk(2, dim=1)[0], latents, torch.full((1,), maxt).long())

    ns["generate_sample_1step"] = gen_1step
    ns["noise_scheduler"] = sched
    vae = weights.SyntheticVAE(decode)
    _, _, y, mask, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))
    for tag, (H, W), tiled, fix in (("untiled_256x320", (256, 320), False, "wavelet"),
                                    ("tiled_512x576_wavelet", (512, 576), True, "wavelet"),
                                    ("tiled_512x576_adain", (512, 576), True, "adain")):
        img = weights.synthetic_degraded_image(H, W, seed=4)
        t0 = time.time()
        preds, stage1 = ns["process"](tiny, [img], strength=1, color_fix_type=fix, disable_preprocess_model=True,
                                      tiled=tiled, tile_size=512, tile_stride=448, vae=vae, y=y, y_mask=mask)
        np.savez_compressed(GOLD / f"process_{tag}.npz", pred=preds[0], H=H, W=W, tiled=tiled, fix=fix, dit_seed=21,
                            vae_seed=2, img_seed=4, cap_seed=9)
        print(f"process_{tag}: {preds[0].shape} mean {preds[0].mean():.1f} std {preds[0].std():.1f} ({time.time() - t0:.1f}s)", flush=True)
    # standalone colour fixes
    g = torch.Generator().manual_seed(12)
    a, b = torch.rand(1, 3, 96, 128, generator=g), torch.rand(1, 3, 96, 128, generator=g)
    np.savez_compressed(GOLD / "color_fix.npz", content=a.numpy(), style=b.numpy(),
                        wavelet=align_color.wavelet_reconstruction(a, b).numpy(),
                        adain=align_color.adaptive_instance_normalization(a, b).numpy())
    print("done")


if __name__ == "__main__":
    main()
