"""Mint tests/golden/swinir_*.npz by EXECUTING THE UNMODIFIED REFERENCE SwinIR (diffusion/model/swinir.py with the
parameters of configs/swinir.yaml) in the build container -- SURVEY 8f row 2. Third-party imports the container lacks
(pytorch_lightning, timm helpers, lpips) come from oracle/shims. Run:  python oracle/make_goldens_swinir.py"""
from __future__ import annotations

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
for name in ("diffusion", "diffusion.model"):   # bare packages: skip the __init__ files (they pull unrelated modules)
    pkg = types.ModuleType(name)
    pkg.__path__ = [str(REF / name.replace(".", "/"))]
    sys.modules[name] = pkg

from diffusion.model.swinir import SwinIR  # noqa: E402

from instarevive_b200 import weights  # noqa: E402

GOLD = ROOT / "tests" / "golden"
torch.set_grad_enabled(False)
SEED = 7


def main():
    net = SwinIR(img_size=64, patch_size=1, in_chans=3, embed_dim=180, depths=[6] * 8, num_heads=[6] * 8, window_size=8,
                 mlp_ratio=2, sf=8, img_range=1.0, upsampler="nearest+conv", resi_connection="1conv", unshuffle=True,
                 unshuffle_scale=8).eval()
    missing = net.load_state_dict(weights.make_swinir_state_dict(seed=SEED), strict=False)
    assert not missing.unexpected_keys
    assert all(k.endswith("relative_position_index") or k.endswith("attn_mask") or k.startswith("lpips") for k in missing.missing_keys), missing.missing_keys
    for tag, (B, H, W), seed in (("b1_128x128", (1, 128, 128), 30), ("b2_64x192", (2, 64, 192), 31), ("b1_256x256", (1, 256, 256), 33)):
        imgs = [weights.synthetic_degraded_image(H, W, seed=seed + i) for i in range(B)]
        x = torch.from_numpy(np.stack(imgs)).float().div(255.0).permute(0, 3, 1, 2).contiguous()
        t0 = time.time()
        y = net(x)
        np.savez_compressed(GOLD / f"swinir_{tag}.npz", out=y.numpy(), wseed=SEED, img_seed=seed, B=B, H=H, W=W)
        print(f"swinir_{tag}: {tuple(y.shape)} mean {y.mean():.3f} std {y.std():.3f} min {y.min():.3f} max {y.max():.3f} "
              f"input std {x.std():.3f} ({time.time() - t0:.1f}s)", flush=True)


if __name__ == "__main__":
    main()
