"""Euler sampler and the classifier-free-guidance denoising step, restated.

Reference (under /root/reference/src/refiners/foundationals/latent_diffusion/):
  solvers/solver.py:113-435   schedule (quadratic betas 8.5e-4..1.2e-2 over 1000 steps), linspace timesteps
  solvers/euler.py:13-100     sigmas = interp(noise_std / cumulative_scale), scale_model_input, update
  model.py:128-159            LatentDiffusionModel.forward: cat(x, x) -> scale -> unet -> CFG -> solver
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""

from __future__ import annotations

from typing import Callable

import numpy as np
import torch
from torch import Tensor


class EulerSchedule:
    def __init__(self, num_inference_steps: int, dtype: torch.dtype = torch.float32) -> None:
        betas = torch.linspace(8.5e-4**0.5, 1.2e-2**0.5, 1000) ** 2          # solver.py:394-416 (power 2)
        alphas_cumprod = (1 - betas).cumprod(dim=0)
        cumulative_scale = torch.sqrt(alphas_cumprod)                         # solver.py:160
        noise_std = torch.sqrt(1.0 - alphas_cumprod)                          # solver.py:161
        self.timesteps = torch.tensor(np.linspace(0, 999, num_inference_steps), dtype=torch.float32).flip(0)  # :218-223
        table = noise_std / cumulative_scale
        sig = torch.tensor(np.interp(self.timesteps, np.arange(0, 1000), table))  # euler.py:56-61
        self.sigmas = torch.cat([sig, torch.tensor([0.0])]).to(dtype)
        self.num_inference_steps = num_inference_steps

    @property
    def init_noise_sigma(self) -> Tensor:
        return self.sigmas.max()

    def scale_model_input(self, x: Tensor, step: int) -> Tensor:
        """euler.py:63-78."""
        if step == -1:
            return x * self.init_noise_sigma
        return x / ((self.sigmas[step] ** 2 + 1) ** 0.5)

    def update(self, x: Tensor, predicted_noise: Tensor, step: int) -> Tensor:
        """euler.py:80-100 (noise prediction)."""
        return x + predicted_noise * (self.sigmas[step + 1] - self.sigmas[step])


def denoise_step(
    unet: Callable[[Tensor, Tensor], Tensor], schedule: EulerSchedule, x: Tensor, step: int, condition_scale: float
) -> Tensor:
    """model.py:128-159 with classifier-free guidance: ``unet(latents, timestep)`` receives the
    doubled, sigma-scaled batch (unconditional half first)."""
    timestep = schedule.timesteps[step].unsqueeze(0)
    latents = schedule.scale_model_input(torch.cat((x, x)), step)
    uncond, cond = unet(latents, timestep).chunk(2)
    eps = uncond + condition_scale * (cond - uncond)
    return schedule.update(x, eps, step)
