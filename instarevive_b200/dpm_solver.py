"""Multi-step sampler of the reference's 20-step ControlNet evaluation:
`DPMS(model.forward_with_dpmsolver, condition, uncondition, cfg_scale, model_kwargs).sample(z, steps=20, order=2,
skip_type="time_uniform", method="multistep")` -- diffusion/dpm_solver.py:6-35, called at
test_scripts/test_controlnet.py:141-152.

What is restated (diffusion/model/dpm_solver.py): the discrete VP noise schedule (:5-170, piecewise-linear log-alpha
over t = n/N, `interpolate_fn` :1285-1324), the discrete-time model wrapper for a noise-prediction model with
classifier-free guidance (:172-336), the data-prediction conversion of DPM-Solver++ (:435-444) and the multistep
solver of order 1 / 2 with `lower_order_final` (:551-597, :805-863, :1201-1243). Singlestep / adaptive / order-3
variants, thresholding and the correcting hooks are not part of that call and are not provided.

The schedule scalars are host-side float64; the latent arithmetic (x0 = (x - sigma*eps)/alpha, the guidance mix and the
fused state update x <- ca*x + c0*m0 + c1*m1) runs in one small CUDA kernel through the C ABI (`ir_lincomb3`).
No CPU path.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence

import numpy as np
import torch

from . import _lib


def get_named_beta_schedule_linear(num_diffusion_timesteps: int = 1000) -> np.ndarray:
    """diffusion/model/gaussian_diffusion.py:99-116, schedule "linear" (float64)."""
    scale = 1000 / num_diffusion_timesteps
    return np.linspace(scale * 0.0001, scale * 0.02, num_diffusion_timesteps, dtype=np.float64)


class NoiseScheduleVP:
    """dpm_solver.py:5-170, `schedule='discrete'` given betas."""

    def __init__(self, betas: np.ndarray, clipped_lambda: float = -5.1):
        log_alphas = 0.5 * np.cumsum(np.log(1.0 - np.asarray(betas, dtype=np.float64)))
        # numerical_clip_alpha (:114-125): drop the tail whose half-logSNR is below clipped_lambda
        log_sigmas = 0.5 * np.log(1.0 - np.exp(2.0 * log_alphas))
        lambs = log_alphas - log_sigmas
        idx = int(np.searchsorted(lambs[::-1], clipped_lambda))
        if idx > 0:
            log_alphas = log_alphas[:-idx]
        self.T = 1.0
        self.log_alpha_array = log_alphas
        self.total_N = len(log_alphas)
        self.t_array = np.linspace(0.0, 1.0, self.total_N + 1)[1:]

    def marginal_log_mean_coeff(self, t: float) -> float:
        """Piecewise-linear in t over the keypoints, linearly extrapolated beyond them (interpolate_fn :1285-1324)."""
        xp, yp = self.t_array, self.log_alpha_array
        k = int(np.searchsorted(xp, t, side="left"))   # number of keypoints strictly below t
        i0 = min(max(k - 1, 0), len(xp) - 2)
        return float(yp[i0] + (t - xp[i0]) * (yp[i0 + 1] - yp[i0]) / (xp[i0 + 1] - xp[i0]))

    def marginal_alpha(self, t: float) -> float:
        return math.exp(self.marginal_log_mean_coeff(t))

    def marginal_std(self, t: float) -> float:
        return math.sqrt(1.0 - math.exp(2.0 * self.marginal_log_mean_coeff(t)))

    def marginal_lambda(self, t: float) -> float:
        lm = self.marginal_log_mean_coeff(t)
        return lm - 0.5 * math.log(1.0 - math.exp(2.0 * lm))


def multistep_coefficients(ns: NoiseScheduleVP, steps: int, order: int = 2, t_start: Optional[float] = None,
                           t_end: Optional[float] = None, lower_order_final: bool = True) -> List[Dict[str, float]]:
    """The scalar plan of `sample(method='multistep', skip_type='time_uniform')` (:1181-1243): for every model
    evaluation its continuous time, model-input time, alpha, sigma; for every state update x <- ca*x + c0*m0 + c1*m1.
    Entry i describes the evaluation at timesteps[i] and the update that leads to timesteps[i+1]."""
    if order not in (1, 2) or steps < order:
        raise ValueError("multistep DPM-Solver++ of order 1 or 2 with steps >= order")
    t_0 = 1.0 / ns.total_N if t_end is None else t_end
    t_T = ns.T if t_start is None else t_start
    ts = np.linspace(t_T, t_0, steps + 1)
    plan = []
    for i in range(steps):
        s, t = float(ts[i]), float(ts[i + 1])
        lam_s, lam_t = ns.marginal_lambda(s), ns.marginal_lambda(t)
        h = lam_t - lam_s
        ca = ns.marginal_std(t) / ns.marginal_std(s)
        b = ns.marginal_alpha(t) * math.expm1(-h)        # alpha_t * phi_1
        step = i + 1                                      # index of the update in the reference's loop
        if step < order:
            step_order = step                             # warm-up by lower order (:1215-1224)
        else:
            step_order = min(order, steps + 1 - step) if lower_order_final else order   # (:1226-1232)
        entry = {"t": s, "t_next": t, "t_input": (s - 1.0 / ns.total_N) * 1000.0, "alpha": ns.marginal_alpha(s),
                 "sigma": ns.marginal_std(s), "order": step_order}
        if step_order == 1:      # dpm_solver_first_update, dpmsolver++ branch (:572-583)
            entry.update(ca=ca, c0=-b, c1=0.0)
        else:                    # multistep_dpm_solver_second_update, 'dpmsolver' type (:833-846)
            lam_p = ns.marginal_lambda(float(ts[i - 1]))
            r0 = (lam_s - lam_p) / h
            entry.update(ca=ca, c0=-b - 0.5 * b / r0, c1=0.5 * b / r0)
        plan.append(entry)
    return plan


class DPM_Solver_pp:
    """The object DPMS() returns: `.sample(x, steps, order, skip_type, method)`."""

    def __init__(self, model: Callable, ns: NoiseScheduleVP, condition, uncondition, cfg_scale: float, model_kwargs: dict):
        self.model, self.ns = model, ns
        self.condition, self.uncondition, self.cfg_scale = condition, uncondition, float(cfg_scale)
        self.model_kwargs = model_kwargs

    def _noise(self, x: torch.Tensor, t_input: float) -> torch.Tensor:
        """model_wrapper.model_fn, guidance_type 'classifier-free' (:322-331)."""
        B = x.shape[0]
        t = torch.full((B,), t_input, device=x.device, dtype=torch.float32)
        if self.cfg_scale == 1.0 or self.uncondition is None:
            return self.model(x, t, self.condition, **self.model_kwargs)
        x_in = torch.cat([x] * 2)
        c_in = torch.cat([self.uncondition, self.condition])
        kw = {k: (torch.cat([v] * 2) if torch.is_tensor(v) and v.shape[:1] == x.shape[:1] else v)
              for k, v in self.model_kwargs.items()}
        out = self.model(x_in, torch.cat([t] * 2), c_in, **kw)
        n_un, n_c = out.chunk(2)
        L = _lib.lib()
        res = torch.empty_like(n_c, memory_format=torch.contiguous_format)
        n_un, n_c = n_un.contiguous(), n_c.contiguous()
        with torch.cuda.device(x.device):   # noise_uncond + s * (noise - noise_uncond)
            _lib.check(L.ir_lincomb3(n_un.data_ptr(), n_c.data_ptr(), None, res.data_ptr(), res.numel(),
                                     1.0 - self.cfg_scale, self.cfg_scale, 0.0, _lib.stream_ptr()), "ir_lincomb3")
        return res

    @torch.no_grad()
    def sample(self, x: torch.Tensor, steps: int = 20, t_start=None, t_end=None, order: int = 2,
               skip_type: str = "time_uniform", method: str = "multistep", lower_order_final: bool = True,
               return_intermediate: bool = False):
        if skip_type != "time_uniform" or method != "multistep":
            raise NotImplementedError("only skip_type='time_uniform', method='multistep' (the reference's call) is provided")
        if x.device.type != "cuda":
            raise RuntimeError("instarevive_b200 has no CPU path: the sampler state must be a CUDA tensor")
        L = _lib.lib()
        plan = multistep_coefficients(self.ns, steps, order, t_start, t_end, lower_order_final)
        x = x.to(torch.float32).contiguous().clone()
        m_prev = None
        inter = []
        with torch.cuda.device(x.device):
            for e in plan:
                eps = self._noise(x, e["t_input"]).to(torch.float32).contiguous()
                if eps.shape != x.shape:
                    raise ValueError("the wrapped model must return a noise prediction of the state's shape")
                # data prediction x0 = (x - sigma * eps) / alpha (:435-444)
                m0 = torch.empty_like(x)
                _lib.check(L.ir_lincomb3(x.data_ptr(), eps.data_ptr(), None, m0.data_ptr(), x.numel(), 1.0 / e["alpha"],
                                         -e["sigma"] / e["alpha"], 0.0, _lib.stream_ptr()), "ir_lincomb3")
                m1 = m_prev if e["order"] == 2 else None
                _lib.check(L.ir_lincomb3(x.data_ptr(), m0.data_ptr(), m1.data_ptr() if m1 is not None else None,
                                         x.data_ptr(), x.numel(), e["ca"], e["c0"], e["c1"], _lib.stream_ptr()),
                           "ir_lincomb3")
                m_prev = m0
                if return_intermediate:
                    inter.append(x.clone())
        return (x, inter) if return_intermediate else x


def DPMS(model: Callable, condition, uncondition, cfg_scale: float, model_type: str = "noise",
         noise_schedule: str = "linear", guidance_type: str = "classifier-free", model_kwargs: Optional[dict] = None,
         diffusion_steps: int = 1000) -> DPM_Solver_pp:
    """diffusion/dpm_solver.py:6-35 (noise-prediction model, linear betas, classifier-free guidance, dpmsolver++)."""
    if model_type != "noise" or noise_schedule != "linear" or guidance_type != "classifier-free":
        raise NotImplementedError("DPMS: the reference's call uses model_type='noise', 'linear' betas, classifier-free guidance")
    ns = NoiseScheduleVP(get_named_beta_schedule_linear(diffusion_steps))
    return DPM_Solver_pp(model, ns, condition, uncondition, cfg_scale, model_kwargs or {})
