#!/usr/bin/env python
"""README CLI of the reference (`python inference.py --ckpt .. --input .. --output .. [--tiled] --sr_scale N`) on the
B200-native path. The argparse surface is the one of the reference's test_scripts/inference.py:169-205; the loop body
follows its main() (:230-347): resize by --sr_scale, auto-resize / pad to x64, process(), crop, LANCZOS back, save PNG.

What differs, by construction of the hot-path scope:
  * the generator runs on hand-written sm_100a kernels; `--ckpt` is a torch.save'd state_dict -- the diffusers
    Transformer2DModel keys of the released InstaRevive_v1.ckpt (flavour (B): instarevive_b200.Transformer2DModel /
    ControlTransformerHalf) or the reference's PixArt keys (base_model.* / controlnet.* or bare; flavour (A):
    ControlPixArtMSHalf) -- or the literal `random:<seed>` for the seeded random-init weights used in the parity tests
    (no checkpoint ships offline);
  * the VAE weights come from `--vae_ckpt` (a state_dict with the AutoencoderKL keys post_quant_conv.*, decoder.*,
    encoder.*, quant_conv.*) or `random:<seed>`; encode and decode both run on the sm_100a kernels
    (instarevive_b200.AutoencoderKL). A checkpoint that holds only the decoder keys falls back to the synthetic stride-8
    projection of instarevive_b200.weights.SyntheticVAE for the encoder. The SwinIR stage-1 model (reference:
    ./weights/general_swinir_v1.ckpt, inference.py:245-248) runs on the sm_100a kernels too (instarevive_b200.SwinIR) from
    `--swinir_ckpt` (a state_dict with the reference's keys, or `random:<seed>`); without it
    `--disable_preprocess_model` is implied;
  * the caption embedding is read from `--caption_embeds` (a .pth with 'caption_embeds' and 'emb_mask', as the
    reference loads at :256-259) or synthesised.
"""
from __future__ import annotations

import math
import os
import sys
from argparse import ArgumentParser, Namespace
from pathlib import Path

import numpy as np
import torch
from PIL import Image

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

import instarevive_b200 as ir  # noqa: E402
from instarevive_b200 import weights  # noqa: E402


def parse_args() -> Namespace:
    p = ArgumentParser()
    p.add_argument("--ckpt", required=True, type=str, help="state_dict path or random:<seed>")
    p.add_argument("--input", type=str, required=True)
    p.add_argument("--sr_scale", type=float, default=1)
    p.add_argument("--repeat_times", type=int, default=1)
    p.add_argument("--disable_preprocess_model", action="store_true")
    # patch-based sampling
    p.add_argument("--tiled", action="store_true")
    p.add_argument("--tile_size", type=int, default=512)
    p.add_argument("--tile_stride", type=int, default=448)
    # latent image guidance (accepted for CLI compatibility; the one-step path never used them)
    p.add_argument("--use_guidance", action="store_true")
    p.add_argument("--g_scale", type=float, default=0.0)
    p.add_argument("--g_t_start", type=int, default=1001)
    p.add_argument("--g_t_stop", type=int, default=-1)
    p.add_argument("--g_space", type=str, default="latent")
    p.add_argument("--g_repeat", type=int, default=5)
    p.add_argument("--color_fix_type", type=str, default="wavelet", choices=["wavelet", "adain", "none"])
    p.add_argument("--output", type=str, required=True)
    p.add_argument("--show_lq", action="store_true")
    p.add_argument("--skip_if_exist", action="store_true")
    p.add_argument("--seed", type=int, default=231)
    p.add_argument("--device", type=str, default="cuda", choices=["cpu", "cuda", "mps"])
    p.add_argument("--use_prompt", action="store_true")
    p.add_argument("--use_center_crop", action="store_true")
    # additions of this implementation
    p.add_argument("--vae_ckpt", type=str, default="random:2")
    p.add_argument("--caption_embeds", type=str, default=None)
    p.add_argument("--swinir_ckpt", type=str, default=None,
                   help="stage-1 SwinIR state_dict (reference keys) or random:<seed>; omitted = stage 1 disabled")
    return p.parse_args()


IMG_EXT = (".jpg", ".jpeg", ".png", ".bmp", ".webp")


def list_image_files(root: str):
    out = []
    for d, _, files in sorted(os.walk(root, followlinks=True)):
        out += [os.path.join(d, f) for f in sorted(files) if f.lower().endswith(IMG_EXT)]
    return out


def auto_resize(img: Image.Image, size: int) -> Image.Image:
    """utils/image/common.py:229-239."""
    short = min(img.size)
    if short < size:
        r = size / short
        return img.resize(tuple(math.ceil(x * r) for x in img.size), Image.BICUBIC)
    return img.copy()


def pad(img: np.ndarray, scale: int) -> np.ndarray:
    """utils/image/common.py:242-249."""
    h, w = img.shape[:2]
    ph = 0 if h % scale == 0 else math.ceil(h / scale) * scale - h
    pw = 0 if w % scale == 0 else math.ceil(w / scale) * scale - w
    return np.pad(img, ((0, ph), (0, pw), (0, 0)), mode="constant", constant_values=0)


def center_crop_arr(pil_image: Image.Image, image_size: int) -> Image.Image:
    """utils/image/common.py:12-36 (ADM centre crop)."""
    while min(*pil_image.size) >= 2 * image_size:
        pil_image = pil_image.resize(tuple(x // 2 for x in pil_image.size), resample=Image.BOX)
    scale = image_size / min(*pil_image.size)
    pil_image = pil_image.resize(tuple(round(x * scale) for x in pil_image.size), resample=Image.BICUBIC)
    arr = np.array(pil_image)
    cy, cx = (arr.shape[0] - image_size) // 2, (arr.shape[1] - image_size) // 2
    return Image.fromarray(arr[cy:cy + image_size, cx:cx + image_size])


def _load_sd(spec: str, maker):
    if spec.startswith("random:"):
        return maker(int(spec.split(":")[1]))
    sd = torch.load(spec, map_location="cpu")
    return sd.get("state_dict", sd)


def build_generator(sd):
    """The generator class follows the checkpoint: the released `InstaRevive_v1.ckpt` is a bare diffusers
    Transformer2DModel state dict (reference: inference.py:238-242) -> flavour (B), 28 blocks, no control branch;
    `base_model.* / controlnet.*` checkpoints get the ControlNet-Half wrapper of their flavour (13 copied blocks,
    pixart_controlnet.py:188 / transformer_controlnet.py:58)."""
    from instarevive_b200 import convert
    n_ctrl = 0
    while f"controlnet.{n_ctrl}.after_proj.weight" in sd:
        n_ctrl += 1
    if convert.is_diffusers_layout(sd):
        model = ir.Transformer2DModel(sample_size=64)
        if n_ctrl:
            model = ir.ControlTransformerHalf(model, n_ctrl)
    else:
        model = ir.ControlPixArtMSHalf(ir.PixArtMS_XL_2(input_size=64, micro_condition=True, init_weights=False), n_ctrl)
    model.load_state_dict(sd, strict=True)
    return model.eval()


def main() -> None:
    args = parse_args()
    torch.manual_seed(args.seed)
    np.random.seed(args.seed)
    if args.device != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("instarevive_b200 runs on CUDA (sm_100a) only; there is no CPU/MPS path")
    dev = torch.device("cuda")

    model = build_generator(_load_sd(args.ckpt, lambda s: weights.make_dit_state_dict(28, 13, seed=s))).to(dev)

    vae_sd = _load_sd(args.vae_ckpt, lambda s: weights.make_vae_state_dict(dec_seed=s))
    if any(k.startswith("encoder.") for k in vae_sd):
        vae = ir.AutoencoderKL(vae_sd, device=dev)   # encode + decode on the device (reference: inference.py:104-117)
    else:
        vae = ir.AutoencoderKLDecoder(vae_sd, device=dev, encoder=weights.SyntheticVAE(None).encode)

    preprocess_model = None
    if args.swinir_ckpt and not args.disable_preprocess_model:
        preprocess_model = ir.SwinIR(_load_sd(args.swinir_ckpt, lambda s: weights.make_swinir_state_dict(seed=s)), device=dev)
    disable_pre = args.disable_preprocess_model or preprocess_model is None

    if args.caption_embeds:
        cap = torch.load(args.caption_embeds, map_location="cpu")
        y_null, y_null_mask = cap["caption_embeds"].to(dev), cap["emb_mask"].to(dev)
    else:
        _, _, y_s, m_s, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))
        y_null, y_null_mask = y_s[0, 0].to(dev), m_s[0, 0, 0].to(dev)
    y = y_null.reshape(1, 1, -1, y_null.shape[-1]).to(torch.float32)           # (1, 1, L, 4096)
    y_mask = y_null_mask.reshape(1, 1, 1, -1).to(torch.float32)                # (1, 1, 1, L)

    assert os.path.isdir(args.input)
    for file_path in list_image_files(args.input):
        lq = Image.open(file_path).convert("RGB")
        if args.sr_scale != 1:
            lq = lq.resize(tuple(math.ceil(x * args.sr_scale) for x in lq.size), Image.BICUBIC)
        if not args.tiled:
            if args.use_center_crop:
                lq_resized = center_crop_arr(lq, 512)
                x = np.array(lq_resized)
            else:
                lq_resized = auto_resize(lq, 512)
                x = pad(np.array(lq_resized), scale=64)
        else:
            lq_resized = auto_resize(lq, args.tile_size)
            x = pad(np.array(lq_resized), scale=64)
        for i in range(args.repeat_times):
            save_path = os.path.join(args.output, os.path.relpath(file_path, args.input))
            parent, name = os.path.split(save_path)
            stem = os.path.splitext(name)[0]
            save_path = os.path.join(parent, f"{stem}_{i}.png")
            if os.path.exists(save_path) and args.skip_if_exist:
                print(f"skip {save_path}")
                continue
            os.makedirs(parent, exist_ok=True)
            preds, stage1_preds = ir.process(
                model, [x], strength=1, color_fix_type=args.color_fix_type, disable_preprocess_model=disable_pre,
                tiled=args.tiled, tile_size=args.tile_size, tile_stride=args.tile_stride, vae=vae,
                preprocess_model=preprocess_model, y=y, y_mask=y_mask,
                # a checkpoint that carries a ControlNet-Half branch is run WITH it (c = degraded latent); a plain
                # generator checkpoint takes the reference's literal c=None path (inference.py:114,131)
                use_control=getattr(model, "copy_blocks_num", 0) > 0)
            pred, stage1_pred = preds[0], stage1_preds[0]
            if not args.use_center_crop:  # remove padding
                pred = pred[:lq_resized.height, :lq_resized.width, :]
                stage1_pred = stage1_pred[:lq_resized.height, :lq_resized.width, :]
            if args.show_lq:
                if not args.use_center_crop:
                    pred = np.array(Image.fromarray(pred).resize(lq.size, Image.LANCZOS))
                    stage1_pred = np.array(Image.fromarray(stage1_pred).resize(lq.size, Image.LANCZOS))
                    lq_arr = np.array(lq)
                else:
                    lq_arr = x
                images = [lq_arr, pred] if disable_pre else [lq_arr, stage1_pred, pred]
                Image.fromarray(np.concatenate(images, axis=1)).save(save_path)
            elif not args.use_center_crop:
                Image.fromarray(pred).resize(lq.size, Image.LANCZOS).save(save_path)
            else:
                Image.fromarray(pred).save(save_path)
            print(f"save to {save_path}")


if __name__ == "__main__":
    main()
