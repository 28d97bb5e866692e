"""Mint tests/golden/dpm_*.npz by EXECUTING THE REFERENCE's multi-step sampler:
diffusion/dpm_solver.py:DPMS + diffusion/model/dpm_solver.py (NoiseScheduleVP, model_wrapper, DPM_Solver), imported
unmodified from /root/reference. Run:  python oracle/make_goldens_sampler.py

 * dpm_plan_*.npz     the solver's scalars (time steps, alpha, sigma, lambda) from the reference's NoiseScheduleVP, and
                      the trajectory of DPMS(...).sample() on an analytic toy noise model (pins the solver arithmetic)
 * dpm_dit_5step.npz  DPMS(ControlPixArtMSHalf.forward_with_dpmsolver, ...).sample(z, steps=5, order=2) with the
                      reference's own network (depth 4 + 2 control blocks, seeded weights) on CPU in fp32
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(ROOT / "oracle" / "shims"))
sys.path.insert(0, str(REF))
sys.path.insert(0, str(ROOT))
pkg = types.ModuleType("diffusion")   # bare package: skip diffusion/__init__.py (pulls unrelated samplers)
pkg.__path__ = [str(REF / "diffusion")]
sys.modules["diffusion"] = pkg

from diffusion.dpm_solver import DPMS  # noqa: E402
from diffusion.model import gaussian_diffusion as gd  # noqa: E402
from diffusion.model.dpm_solver import NoiseScheduleVP  # noqa: E402
from diffusion.model.nets.PixArtMS import PixArtMS  # noqa: E402
from diffusion.model.nets.pixart_controlnet import ControlPixArtMSHalf  # noqa: E402

from instarevive_b200 import weights  # noqa: E402

GOLD = ROOT / "tests" / "golden"
torch.set_grad_enabled(False)


def toy_model(x, t_input, cond, scale=1.0):
    """Analytic stand-in for a noise-prediction network: smooth in x and in the model-input time."""
    t = t_input.view(-1, 1, 1, 1) / 1000.0
    return scale * (0.6 * x * torch.cos(2.0 * t) + 0.25 * torch.sin(3.0 * x + t) + 0.1 * cond.view(-1, 1, 1, 1))


def main():
    betas = torch.tensor(gd.get_named_beta_schedule("linear", 1000))
    ns = NoiseScheduleVP(schedule="discrete", betas=betas)
    for steps in (5, 20):
        ts = torch.linspace(ns.T, 1.0 / ns.total_N, steps + 1)
        z = torch.randn(2, 4, 8, 8, generator=torch.Generator().manual_seed(steps))
        cond, uncond = torch.tensor([0.3, -0.7]), torch.tensor([0.0, 0.0])
        out = {}
        for cfg in (1.0, 4.5):
            solver = DPMS(toy_model, condition=cond, uncondition=uncond, cfg_scale=cfg, model_kwargs=dict(scale=0.9))
            x_end, inter = solver.sample(z, steps=steps, order=2, skip_type="time_uniform", method="multistep",
                                         return_intermediate=True)
            out[f"x_end_cfg{cfg}"] = x_end.numpy()
            out[f"inter_cfg{cfg}"] = torch.stack(inter).numpy()
        np.savez_compressed(GOLD / f"dpm_plan_{steps}.npz", steps=steps, timesteps=ts.numpy(),
                            alpha=ns.marginal_alpha(ts).numpy(), sigma=ns.marginal_std(ts).numpy(),
                            lam=ns.marginal_lambda(ts).numpy(), total_N=ns.total_N, z=z.numpy(), **out)
        print(f"dpm_plan_{steps}: total_N {ns.total_N}, x_end std {out['x_end_cfg1.0'].std():.4f}")

    depth, cb = 4, 2
    net = ControlPixArtMSHalf(PixArtMS(depth=depth, input_size=64, micro_condition=True), copy_blocks_num=cb).eval()
    net.load_state_dict(weights.make_dit_state_dict(depth=depth, copy_blocks=cb, seed=11), strict=True)
    x, _, y, mask, info = weights.make_inputs(1, 32, 32, seed=6, lens=(77,))
    c = torch.randn(1, 4, 32, 32, generator=torch.Generator().manual_seed(61))
    solver = DPMS(net.forward_with_dpmsolver, condition=y, uncondition=None, cfg_scale=1.0,
                  model_kwargs=dict(data_info=info, mask=mask, c=c))
    x_end = solver.sample(x, steps=5, order=2, skip_type="time_uniform", method="multistep")
    np.savez_compressed(GOLD / "dpm_dit_5step.npz", x_end=x_end.numpy(), c=c.numpy(), depth=depth, copy_blocks=cb, wseed=11,
                        iseed=6, steps=5)
    print("dpm_dit_5step: x_end std", x_end.std().item(), "max", x_end.abs().max().item())


if __name__ == "__main__":
    main()
