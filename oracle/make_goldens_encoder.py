"""Mint tests/golden/vae_enc_*.npz by EXECUTING THE UNMODIFIED REFERENCE Encoder in the build container
(/root/reference does not exist on the GPU box, so the vectors are committed).
Run:  python oracle/make_goldens_encoder.py

What runs: ldm.modules.diffusionmodules.model.Encoder (ddconfig of configs/cldm.yaml:69-84), a Conv2d(8, 8, 1) composed as
AutoencoderKL.encode does (ldm/models/autoencoder.py:82-86; the class itself needs pytorch_lightning) and
ldm.modules.distributions.distributions.DiagonalGaussianDistribution.mode(). Weights: instarevive_b200/weights.py.
"""
from __future__ import annotations

import sys
import time
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle" / "shims"))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT))

import ldm.xformers_state as _xs  # noqa: E402
_xs.disable_xformers()
from ldm.modules.diffusionmodules.model import Encoder  # noqa: E402
from ldm.modules.distributions.distributions import DiagonalGaussianDistribution  # noqa: E402

from instarevive_b200 import weights  # noqa: E402

GOLD = ROOT / "tests" / "golden"
torch.set_grad_enabled(False)
ENC_SEED = 5


def build_encoder(seed):
    enc = Encoder(ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                  in_channels=3, resolution=256, z_channels=4, double_z=True).eval()
    qc = torch.nn.Conv2d(8, 8, 1).eval()
    sd = weights.make_vae_encoder_state_dict(seed=seed)
    enc.load_state_dict({k[len("encoder."):]: v for k, v in sd.items() if k.startswith("encoder.")}, strict=True)
    qc.load_state_dict({k[len("quant_conv."):]: v for k, v in sd.items() if k.startswith("quant_conv.")}, strict=True)
    return lambda x: DiagonalGaussianDistribution(qc(enc(x)))  # AutoencoderKL.encode


def image(B, H, W, seed):
    imgs = [weights.synthetic_degraded_image(H, W, seed=seed + i) for i in range(B)]
    x = torch.from_numpy(np.stack(imgs)).float().div(255.0).permute(0, 3, 1, 2).contiguous()
    return x * 2 - 1   # test_scripts/inference.py:106


def main():
    encode = build_encoder(ENC_SEED)
    for tag, (B, H, W), seed in (("b1_128x128", (1, 128, 128), 20), ("b2_96x160", (2, 96, 160), 21),
                                 ("b1_256x256", (1, 256, 256), 23)):
        t0 = time.time()
        post = encode(image(B, H, W, seed))
        moments = post.parameters
        assert torch.equal(post.mode(), moments[:, :4])
        np.savez_compressed(GOLD / f"vae_enc_{tag}.npz", moments=moments.numpy(), wseed=ENC_SEED, img_seed=seed, B=B, H=H, W=W)
        print(f"vae_enc_{tag}: {tuple(moments.shape)} mean-part std {moments[:, :4].std():.3f} max {moments[:, :4].abs().max():.3f} "
              f"logvar-part std {moments[:, 4:].std():.3f} ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
