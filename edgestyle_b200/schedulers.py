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
