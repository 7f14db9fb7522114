"""Host side of the schedulers: timestep tables and per-step coefficients (the update itself runs on the
device in es_cfg_ddim).  Mirrors the diffusers scheduler surface the reference pipeline touches
(/root/reference/model/edgestyle_pipeline.py:448,520-522,668-711): set_timesteps, timesteps,
init_noise_sigma, scale_model_input, order."""
from __future__ import annotations

import numpy as np


def alphas_cumprod(num_train: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012) -> np.ndarray:
    """scaled_linear betas, fp32 torch.cumprod -- the same arithmetic (and rounding) as diffusers."""
    import torch

    betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train, dtype=torch.float32) ** 2
    return torch.cumprod(1.0 - betas, dim=0).numpy()


class DDIMScheduler:
    """DDIM, eta = 0, epsilon prediction, leading spacing, steps_offset = 1, set_alpha_to_one = False,
    clip_sample = False (the SD1.5 scheduler config)."""

    init_noise_sigma = 1.0
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 steps_offset: int = 1):
        self.num_train_timesteps = num_train_timesteps
        self.steps_offset = steps_offset
        self.alphas_cumprod = alphas_cumprod(num_train_timesteps, beta_start, beta_end)
        self.final_alpha_cumprod = self.alphas_cumprod[0]
        self.timesteps = None
        self.num_inference_steps = None

    def set_timesteps(self, num_inference_steps: int, device=None):
        if num_inference_steps > self.num_train_timesteps:
            raise ValueError("num_inference_steps exceeds num_train_timesteps")
        self.num_inference_steps = num_inference_steps
        ratio = self.num_train_timesteps // num_inference_steps
        self.timesteps = (np.arange(0, num_inference_steps) * ratio).round()[::-1].astype(np.int64) + self.steps_offset
        return self.timesteps

    def scale_model_input(self, sample, timestep=None):
        return sample

    def coefficients(self, t: int):
        prev_t = t - self.num_train_timesteps // self.num_inference_steps
        a_t = float(self.alphas_cumprod[t])
        a_prev = float(self.alphas_cumprod[prev_t]) if prev_t >= 0 else float(self.final_alpha_cumprod)
        return a_t, a_prev

    def device_step_ddim(self, engine, latents, t: int):
        a_t, a_prev = self.coefficients(int(t))
        return engine.cfg_ddim_update(latents, a_t, a_prev)


class UniPCMultistepScheduler:
    """UniPC-bh2, solver_order 2, predict_x0, lower_order_final, no Karras sigmas -- the reference's default
    (UniPCMultistepScheduler.from_config(pipeline.scheduler.config), /root/reference/app.py:118).  The host keeps
    the sigma table and turns every predictor / corrector update into scalar coefficients (float64); the tensor
    arithmetic runs on the device (es_cfg_x0 + es_lincomb4).  `timestep_spacing` is explicit because which one
    from_config inherits is unverified without a diffusers install."""

    init_noise_sigma = 1.0
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, solver_order: int = 2, timestep_spacing: str = "leading",
                 steps_offset: int = 1):
        if solver_order != 2:
            raise NotImplementedError("only solver_order = 2 (the reference default)")
        self.num_train_timesteps = num_train_timesteps
        self.spacing, self.steps_offset = timestep_spacing, steps_offset
        self.alphas_cumprod = alphas_cumprod(num_train_timesteps)
        self.timesteps = None

    def set_timesteps(self, n: int, device=None):
        T = self.num_train_timesteps
        if self.spacing == "linspace":
            ts = np.linspace(0, T - 1, n + 1).round()[::-1][:-1].copy().astype(np.int64)
        elif self.spacing == "leading":
            ts = (np.arange(0, n + 1) * (T // (n + 1))).round()[::-1][:-1].copy().astype(np.int64) + self.steps_offset
        else:
            raise ValueError(self.spacing)
        ac = self.alphas_cumprod.astype(np.float64)
        sig = ((1 - ac) / ac) ** 0.5
        self.sigmas = np.concatenate([np.interp(ts, np.arange(len(sig)), sig), [((1 - ac[0]) / ac[0]) ** 0.5]]).astype(np.float32).astype(np.float64)
        self.timesteps = ts
        self.num_inference_steps = n
        self.step_index, self.lower_order_nums, self.this_order = 0, 0, 1
        self._m = [None, None]      # device buffers of the last two x0 predictions (model_outputs)
        self._last_sample = None
        self._have_last = False
        return ts

    def scale_model_input(self, sample, timestep=None):
        return sample

    @staticmethod
    def _alpha_sigma(sigma):
        a = 1.0 / (sigma * sigma + 1.0) ** 0.5
        return a, sigma * a

    def _lam(self, i):
        a, s = self._alpha_sigma(self.sigmas[i])
        return a, s, np.log(a) - np.log(s)

    @staticmethod
    def _rb(rks, order, hh):
        h_phi_1 = np.expm1(hh)
        h_phi_k = h_phi_1 / hh - 1
        B_h = np.expm1(hh)
        R, b, fact = [], [], 1
        for i in range(1, order + 1):
            R.append(np.power(rks, i - 1))
            b.append(h_phi_k * fact / B_h)
            fact *= i + 1
            h_phi_k = h_phi_k / hh - 1 / fact
        return np.stack(R), np.array(b), h_phi_1, B_h

    def device_step(self, eps, latents, guidance):
        """One scheduler.step on the device: `eps` [2*imgs, ...] raw UNet output (or [imgs, ...] with `guidance` None:
        no classifier-free guidance), `latents` updated in place."""
        import torch

        from . import ops

        i = self.step_index
        if self._m[0] is None:
            self._m = [torch.empty_like(latents), torch.empty_like(latents)]
            self._last_sample = torch.empty_like(latents)
            self._x0 = torch.empty_like(latents)
        alpha_t, sigma_t = self._alpha_sigma(self.sigmas[i])
        if guidance is None:                                                # x0 = (x - sigma eps) / alpha
            ops.lincomb(self._x0, [(1.0 / alpha_t, latents), (-sigma_t / alpha_t, eps)])
        else:
            ops.cfg_x0(eps, latents, guidance, alpha_t, sigma_t, self._x0)  # model_output_convert (on the pre-corrector sample)
        m_prev, m_prev2 = self._m[1], self._m[0]                            # model_outputs[-1], [-2] before the shift
        if i > 0 and self._have_last:                                       # UniC corrector with the previous order
            order = self.this_order
            a_t, s_t, lam_t = self._lam(i)
            a_s0, s_s0, lam_s0 = self._lam(i - 1)
            h = lam_t - lam_s0
            rks, have_d1 = [], order >= 2
            if have_d1:
                _, _, lam_si = self._lam(i - 2)
                rks.append((lam_si - lam_s0) / h)
            rks.append(1.0)
            R, b, h_phi_1, B_h = self._rb(np.array(rks), order, -h)
            rhos = np.array([0.5]) if order == 1 else np.linalg.solve(R, b)
            # x = s_t/s_s0 * last - a_t*h_phi_1*m0 - a_t*B_h*( rho0*(m1-m0)/rk + rho_last*(m_t - m0) ),  m0 = m_prev
            c_last = s_t / s_s0
            c_m0 = -a_t * h_phi_1 + a_t * B_h * rhos[-1]
            c_mt = -a_t * B_h * rhos[-1]
            c_m1 = 0.0
            if have_d1:
                c_m1 = -a_t * B_h * rhos[0] / rks[0]
                c_m0 += a_t * B_h * rhos[0] / rks[0]
            ops.lincomb(latents, [(c_last, self._last_sample), (c_m0, m_prev), (c_mt, self._x0),
                                  (c_m1, m_prev2 if have_d1 else None)])
        # shift the model outputs: [-2] <- [-1], [-1] <- x0
        self._m[0], self._m[1] = self._m[1], self._m[0]
        self._m[1].copy_(self._x0)
        this_order = min(2, len(self.timesteps) - i)                        # lower_order_final
        self.this_order = min(this_order, self.lower_order_nums + 1)
        self._last_sample.copy_(latents)
        self._have_last = True
        # UniP predictor
        order = self.this_order
        a_t, s_t, lam_t = self._lam(i + 1)
        a_s0, s_s0, lam_s0 = self._lam(i)
        h = lam_t - lam_s0
        m0, m1 = self._m[1], self._m[0]
        c_x = s_t / s_s0
        h_phi_1, B_h = np.expm1(-h), np.expm1(-h)
        c_m0 = -a_t * h_phi_1
        c_m1 = 0.0
        if order == 2:
            _, _, lam_si = self._lam(i - 1)
            rk = (lam_si - lam_s0) / h
            c_m1 = -a_t * B_h * 0.5 / rk
            c_m0 += a_t * B_h * 0.5 / rk
        ops.lincomb(latents, [(c_x, latents), (c_m0, m0), (c_m1, m1 if order == 2 else None)])
        if self.lower_order_nums < 2:
            self.lower_order_nums += 1
        self.step_index += 1
        return latents
