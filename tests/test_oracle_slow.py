"""Full-size oracle pins (a few tens of seconds of CPU): the 28+13-block forward and the end-to-end process() loop."""
import types

import numpy as np
import torch

from instarevive_b200 import weights
from oracle import dit_oracle, tiles_oracle, vae_oracle

torch.set_grad_enabled(False)


def test_dit_oracle_full_model_and_x0(golden_dir):
    g = np.load(golden_dir / "dit_full_b1_64x64.npz")
    sd = weights.make_dit_state_dict(depth=28, copy_blocks=13, seed=1)
    assert len(sd) == 668  # SURVEY 8b weight contract
    x, ts, y, mask, info = weights.make_inputs(1, 64, 64, seed=0, lens=(77,))
    out = dit_oracle.control_pixart_forward(sd, x, ts, y, mask, info, c=x.clone())
    assert (out - torch.from_numpy(g["out"])).abs().max().item() < 5e-4
    x0 = dit_oracle.eps_to_mu(out.chunk(2, dim=1)[0], x, 400)
    gx = np.load(golden_dir / "x0_full_b1_64x64.npz")["x0"]
    assert (x0 - torch.from_numpy(gx)).abs().max().item() < 2e-3


def test_process_matches_reference_loop(golden_dir):
    """oracle process() vs the reference's own process() (lifted by ast) on the same adapters: uint8 images."""
    dit_sd = weights.make_dit_state_dict(depth=2, copy_blocks=1, seed=21)
    vae_sd = weights.make_vae_decoder_state_dict(seed=2)
    vae = weights.SyntheticVAE(lambda z: vae_oracle.vae_decode(vae_sd, z))
    _, _, y, mask, _ = weights.make_inputs(1, 8, 8, seed=9, lens=(77,))

    def one_step(lat):
        return dit_oracle.generate_sample_1step(dit_sd, lat, y, mask, depth=2, copy_blocks=1)

    for tag in ("untiled_256x320", "tiled_512x576_wavelet"):
        g = np.load(golden_dir / f"process_{tag}.npz")
        img = weights.synthetic_degraded_image(int(g["H"]), int(g["W"]), seed=int(g["img_seed"]))
        control = torch.tensor(np.stack([img]) / 255.0, dtype=torch.float32).clamp_(0, 1).permute(0, 3, 1, 2).contiguous()
        init = vae.encode(control * 2 - 1).latent_dist.mode().float() * vae.config.scaling_factor
        pred, _ = tiles_oracle.process(control, init, one_step, lambda z: vae.decode(z).sample,
                                       vae.config.scaling_factor, bool(g["tiled"]), 512, 448, str(g["fix"]))
        diff = np.abs(pred[0].astype(np.int32) - g["pred"].astype(np.int32))
        assert diff.max() <= 1, (tag, diff.max())          # fp32 summation-order noise can flip a truncation
        assert (diff > 0).mean() < 1e-3


def test_dit_oracle_full_model_1024(golden_dir):
    """The oracle port at the benchmark's own size (128x128 latent, T = 4096): pins the restatement that bench.py's
    reference arm / cpu_baseline leg times at 1024x1024 to the unmodified reference's output."""
    g = np.load(golden_dir / "dit_full_b1_128x128.npz")
    sd = weights.make_dit_state_dict(depth=28, copy_blocks=13, seed=1)
    x, ts, y, mask, info = weights.make_inputs(1, 128, 128, seed=0, lens=(77,))
    out = dit_oracle.control_pixart_forward(sd, x, ts, y, mask, info, c=x.clone())
    assert (out - torch.from_numpy(g["out"])).abs().max().item() < 5e-4
    x0 = dit_oracle.eps_to_mu(out.chunk(2, dim=1)[0], x, 400)
    assert (x0 - torch.from_numpy(g["x0"])).abs().max().item() < 2e-3


def test_vae_oracle_1024(golden_dir):
    """Decoder oracle at a 128x128 latent (mid-attention over 16384 tokens) vs the reference Decoder's crops / block means."""
    g = np.load(golden_dir / "vae_b1_128x128.npz")
    sd = weights.make_vae_decoder_state_dict(seed=2)
    z = torch.randn(1, 4, 128, 128, generator=torch.Generator().manual_seed(8)) / 0.18215 * 0.6
    img = vae_oracle.vae_decode(sd, z)
    for (y0, x0), ref in zip(g["crop_origins"], g["crops"]):
        assert (img[0, :, y0:y0 + 128, x0:x0 + 128] - torch.from_numpy(ref)).abs().max().item() < 2e-4
    means = img.reshape(1, 3, 128, 8, 128, 8).mean(dim=(3, 5))
    assert (means - torch.from_numpy(g["means"])).abs().max().item() < 1e-4
