"""Pins the oracle (oracle/*.py) against the golden vectors minted by EXECUTING THE REFERENCE (oracle/make_goldens.py).
CPU only. The reference itself ships no tests or fixtures for this path, so these files are the pin."""
import numpy as np
import pytest
import torch

from instarevive_b200 import weights
from oracle import dit_oracle, swinir_oracle, tiles_oracle, vae_oracle

torch.set_grad_enabled(False)


def _load(golden_dir, name):
    return np.load(golden_dir / name, allow_pickle=False)


def test_alphas_cumprod_known_answer(golden_dir):
    g = _load(golden_dir, "alphas.npz")
    abar = dit_oracle.alphas_cumprod()
    assert abar[400] == float(g["abar400"]) == 0.19357200966664662  # SURVEY a15 known answer
    np.testing.assert_array_equal(abar, g["abar"])


@pytest.mark.parametrize("tag", ["small_b1_64x64", "small_b2_64x96_ragged", "small_b1_32x32_nomask",
                                 "small_b1_64x64_noc", "small_b1_40x72"])
def test_dit_oracle_matches_reference_small(golden_dir, tag):
    g = _load(golden_dir, f"dit_{tag}.npz")
    depth, cb = int(g["depth"]), int(g["copy_blocks"])
    sd = weights.make_dit_state_dict(depth=depth, copy_blocks=cb, seed=int(g["wseed"]))
    x, ts, y, mask, info = weights.make_inputs(int(g["B"]), int(g["h"]), int(g["w"]), seed=int(g["iseed"]),
                                               lens=tuple(int(v) for v in g["lens"]))
    out = dit_oracle.control_pixart_forward(sd, x, ts, y, mask if bool(g["use_mask"]) else None, info,
                                            c=x.clone() if bool(g["use_c"]) else None, depth=depth, copy_blocks=cb)
    ref = torch.from_numpy(g["out"])
    assert out.shape == ref.shape
    # same fp32 arithmetic up to summation order (per-head attention / packed captions)
    assert (out - ref).abs().max().item() < 2e-4


def test_sliding_windows_bit_exact(golden_dir):
    g = _load(golden_dir, "sliding_windows.npz")
    for key in g.files:
        parts = key.split("_")
        h, w = (int(v) for v in parts[0].split("x"))
        t, s = (int(parts[1][1:]), int(parts[2][1:])) if len(parts) == 3 else (64, 56)
        got = np.array(tiles_oracle.sliding_windows(h, w, t, s), dtype=np.int64)
        np.testing.assert_array_equal(got, g[key], err_msg=key)
    # probe values quoted in SURVEY a16
    assert len(tiles_oracle.sliding_windows(64, 64, 64, 56)) == 1
    assert [c[0] for c in tiles_oracle.sliding_windows(128, 64, 64, 56)] == [0, 56, 64]
    assert len(tiles_oracle.sliding_windows(256, 256, 64, 56)) == 25
    cnt = tiles_oracle.count_mask(256, 256, tiles_oracle.sliding_windows(256, 256, 64, 56))
    assert cnt.min() == 1 and cnt.max() == 4


def test_color_fix_matches_reference(golden_dir):
    g = _load(golden_dir, "color_fix.npz")
    a, b = torch.from_numpy(g["content"]), torch.from_numpy(g["style"])
    np.testing.assert_allclose(tiles_oracle.wavelet_reconstruction(a, b).numpy(), g["wavelet"], atol=1e-6)
    np.testing.assert_allclose(tiles_oracle.adaptive_instance_normalization(a, b).numpy(), g["adain"], atol=1e-5)


@pytest.mark.parametrize("tag,shape,seed", [("b1_32x32", (1, 32, 32), 5), ("b2_16x24", (2, 16, 24), 6)])
def test_vae_oracle_matches_reference(golden_dir, tag, shape, seed):
    g = _load(golden_dir, f"vae_{tag}.npz")
    sd = weights.make_vae_decoder_state_dict(seed=int(g["wseed"]))
    B, h, w = shape
    z = torch.randn(B, 4, h, w, generator=torch.Generator().manual_seed(seed)) / 0.18215 * 0.6
    img = vae_oracle.vae_decode(sd, z)
    assert (img - torch.from_numpy(g["img"])).abs().max().item() < 1e-4


def _enc_image(B, H, W, seed):
    imgs = [weights.synthetic_degraded_image(H, W, seed=seed + i) for i in range(B)]
    return torch.from_numpy(np.stack(imgs)).float().div(255.0).permute(0, 3, 1, 2).contiguous() * 2 - 1


@pytest.mark.parametrize("tag", ["b1_128x128", "b2_96x160", "b1_256x256"])
def test_vae_encoder_oracle_matches_reference(golden_dir, tag):
    """Encoder.forward + quant_conv restatement vs the reference Encoder's own output (moments)."""
    g = _load(golden_dir, f"vae_enc_{tag}.npz")
    sd = weights.make_vae_encoder_state_dict(seed=int(g["wseed"]))
    moments = vae_oracle.vae_encode_moments(sd, _enc_image(int(g["B"]), int(g["H"]), int(g["W"]), int(g["img_seed"])))
    ref = torch.from_numpy(g["moments"])
    assert moments.shape == ref.shape
    assert (moments - ref).abs().max().item() < 1e-4


@pytest.mark.parametrize("tag", ["b1_128x128", "b2_64x192", "b1_256x256"])
def test_swinir_oracle_matches_reference(golden_dir, tag):
    """SwinIR.forward restatement vs the reference class's own output (window shift + mask, relative
    position bias, RSTB convs, nearest+conv upsampler; batch 2 and a non-square image)."""
    g = _load(golden_dir, f"swinir_{tag}.npz")
    sd = weights.make_swinir_state_dict(seed=int(g["wseed"]))
    B, H, W = int(g["B"]), int(g["H"]), int(g["W"])
    imgs = [weights.synthetic_degraded_image(H, W, seed=int(g["img_seed"]) + i) for i in range(B)]
    x = torch.from_numpy(np.stack(imgs)).float().div(255.0).permute(0, 3, 1, 2).contiguous()
    out = swinir_oracle.swinir_forward(sd, x)
    assert out.shape == (B, 3, H, W)
    assert (out - torch.from_numpy(g["out"])).abs().max().item() < 1e-4
