"""DDIM and UniPC schedulers restated (oracle; test infrastructure).  SURVEY.md A.5.

The reference calls `scheduler.set_timesteps / scale_model_input / step`
(/root/reference/model/edgestyle_pipeline.py:448,520,705-709) on diffusers 0.26.3 schedulers
(DDIMScheduler in the stock SD1.5 repo config; UniPCMultistepScheduler.from_config at
/root/reference/app.py:118).  SD1.5 scheduler config: scaled_linear betas 0.00085..0.012, 1000 train
steps, steps_offset=1, set_alpha_to_one=False, clip_sample=False, epsilon prediction.
"""
from __future__ import annotations

import math
from typing import List, Optional

import numpy as np
import torch


def alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> torch.Tensor:
    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0)


class DDIMScheduler:
    """eta = 0, leading spacing, steps_offset = 1."""

    init_noise_sigma = 1.0
    order = 1

    def __init__(self, num_train: int = 1000):
        self.num_train = num_train
        self.alphas_cumprod = alphas_cumprod(num_train)
        self.final_alpha_cumprod = self.alphas_cumprod[0]  # set_alpha_to_one=False
        self.timesteps: Optional[torch.Tensor] = None

    def set_timesteps(self, n: int, device=None):
        self.num_inference_steps = n
        ratio = self.num_train // n
        ts = (np.arange(0, n) * ratio).round()[::-1].copy().astype(np.int64) + 1
        self.timesteps = torch.from_numpy(ts).to(device)
        return self.timesteps

    def scale_model_input(self, sample, t=None):
        return sample

    def coefficients(self, t: int):
        """(a_t, a_prev): alpha-bar at t and at the previous timestep."""
        prev_t = t - self.num_train // self.num_inference_steps
        a_t = self.alphas_cumprod[t]
        a_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.final_alpha_cumprod
        return a_t, a_prev

    def step(self, eps, t, sample, eta: float = 0.0, variance_noise=None):
        """diffusers 0.26.3 DDIMScheduler.step: sigma_t = eta sqrt((1 - a') / (1 - a) (1 - a / a')), direction
        sqrt(1 - a' - sigma_t^2) eps, plus sigma_t z when eta > 0 (formulas (12), (16) of the DDIM paper)."""
        a_t, a_prev = self.coefficients(int(t))
        a_t = a_t.to(sample.dtype)
        a_prev = a_prev.to(sample.dtype)
        x0 = (sample - (1 - a_t) ** 0.5 * eps) / a_t ** 0.5
        std = eta * ((1 - a_prev) / (1 - a_t) * (1 - a_t / a_prev)) ** 0.5
        prev = a_prev ** 0.5 * x0 + (1 - a_prev - std ** 2) ** 0.5 * eps
        if eta > 0:
            prev = prev + std * variance_noise
        return prev


class UniPCMultistepScheduler:
    """UniPC-bh2, solver_order 2, predict_x0, lower_order_final, no Karras sigmas.

    `timestep_spacing` is an explicit parameter: from_config(PNDM config) most likely inherits
    "leading" (unverified against a real install); UniPC's own default is "linspace".
    """

    init_noise_sigma = 1.0
    order = 1

    def __init__(self, num_train: int = 1000, solver_order: int = 2, timestep_spacing: str = "leading",
                 steps_offset: int = 1, dtype=torch.float32):
        self.num_train = num_train
        self.solver_order = solver_order
        self.spacing = timestep_spacing
        self.steps_offset = steps_offset
        self.alphas_cumprod = alphas_cumprod(num_train)
        self.dtype = dtype

    def set_timesteps(self, n: int, device=None):
        T = self.num_train
        if self.spacing == "linspace":
            ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ratio = T // (n + 1)
            ts = (np.arange(0, n + 1) * ratio).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        else:
            raise ValueError(self.spacing)
        ac = self.alphas_cumprod.numpy().astype(np.float64)
        sig = ((1 - ac) / ac) ** 0.5
        sig_t = np.interp(ts, np.arange(0, len(sig)), sig)
        sigma_last = ((1 - ac[0]) / ac[0]) ** 0.5
        self.sigmas = torch.from_numpy(np.concatenate([sig_t, [sigma_last]]).astype(np.float32))
        self.timesteps = torch.from_numpy(ts).to(device)
        self.num_inference_steps = n
        self.model_outputs: List[Optional[torch.Tensor]] = [None] * self.solver_order
        self.lower_order_nums = 0
        self.last_sample = None
        self.step_index = 0
        self.this_order = 1
        return self.timesteps

    def scale_model_input(self, sample, t=None):
        return sample

    @staticmethod
    def _alpha_sigma(sigma):
        alpha_t = 1.0 / (sigma ** 2 + 1) ** 0.5
        return alpha_t, sigma * alpha_t

    def _lam(self, idx):
        a, s = self._alpha_sigma(self.sigmas[idx])
        return a, s, torch.log(a) - torch.log(s)

    def _rb(self, rks, order, hh):
        h_phi_1 = torch.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = torch.expm1(hh)  # bh2
        R, b = [], []
        fact = 1
        rks_t = torch.stack(rks)
        for i in range(1, order + 1):
            R.append(rks_t ** (i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return torch.stack(R), torch.stack(b), h_phi_1, B_h

    def _uni_p(self, sample, order):
        m0 = self.model_outputs[-1]
        alpha_t, sigma_t, lam_t = self._lam(self.step_index + 1)
        alpha_s0, sigma_s0, lam_s0 = self._lam(self.step_index)
        h = lam_t - lam_s0
        rks, D1s = [], []
        for i in range(1, order):
            mi = self.model_outputs[-(i + 1)]
            _, _, lam_si = self._lam(self.step_index - i)
            rk = (lam_si - lam_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(torch.tensor(1.0))
        R, b, h_phi_1, B_h = self._rb(rks, order, -h)
        x_t_ = sigma_t / sigma_s0 * sample - alpha_t * h_phi_1 * m0
        if D1s:
            rhos_p = torch.tensor([0.5]) if order == 2 else torch.linalg.solve(R[:-1, :-1], b[:-1])
            pred_res = sum(r * d for r, d in zip(rhos_p, D1s))
        else:
            pred_res = 0
        return (x_t_ - alpha_t * B_h * pred_res).to(sample.dtype)

    def _uni_c(self, this_model_output, last_sample, this_sample, order):
        m0 = self.model_outputs[-1]
        alpha_t, sigma_t, lam_t = self._lam(self.step_index)
        alpha_s0, sigma_s0, lam_s0 = self._lam(self.step_index - 1)
        h = lam_t - lam_s0
        rks, D1s = [], []
        for i in range(1, order):
            mi = self.model_outputs[-(i + 1)]
            _, _, lam_si = self._lam(self.step_index - (i + 1))
            rk = (lam_si - lam_s0) / h
            rks.append(rk)
            D1s.append((mi - m0) / rk)
        rks.append(torch.tensor(1.0))
        R, b, h_phi_1, B_h = self._rb(rks, order, -h)
        rhos_c = torch.tensor([0.5]) if order == 1 else torch.linalg.solve(R, b)
        x_t_ = sigma_t / sigma_s0 * last_sample - alpha_t * h_phi_1 * m0
        corr_res = sum(r * d for r, d in zip(rhos_c[:-1], D1s)) if D1s else 0
        D1_t = this_model_output - m0
        return (x_t_ - alpha_t * B_h * (corr_res + rhos_c[-1] * D1_t)).to(this_sample.dtype)

    def step(self, eps, t, sample):
        use_corrector = self.step_index > 0 and self.last_sample is not None
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[self.step_index])
        x0 = (sample - sigma_t * eps) / alpha_t
        if use_corrector:
            sample = self._uni_c(x0, self.last_sample, sample, self.this_order)
        for i in range(self.solver_order - 1):
            self.model_outputs[i] = self.model_outputs[i + 1]
        self.model_outputs[-1] = x0
        this_order = min(self.solver_order, len(self.timesteps) - self.step_index)  # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self.last_sample = sample
        prev = self._uni_p(sample, self.this_order)
        if self.lower_order_nums < self.solver_order:
            self.lower_order_nums += 1
        self.step_index += 1
        return prev
