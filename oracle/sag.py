"""Self-attention guidance and the DDIM sampler it is used with, restated.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Reference (under /root/reference/src/refiners/foundationals/latent_diffusion/):
  self_attention_guidance.py:77-98   mask from the middle block's attention map, blur, re-noise
  model.py:128-159                   the step: CFG combine, then ``+ sag.scale * (uncond - unet(degraded))``
  stable_diffusion_1/model.py:175-213, stable_diffusion_xl/model.py:194-250   the extra, unconditional pass
  solvers/solver.py:226-228,244-266,300-318   LEADING timesteps, add_noise / remove_noise
  solvers/ddim.py:58-99              the deterministic DDIM update
  fluxion/utils.py:65-113            gaussian_blur
"""

from __future__ import annotations

from typing import Callable

import torch
import torch.nn.functional as F
from torch import Tensor

from oracle import unet as ounet


class DDIMSchedule:
    def __init__(self, num_inference_steps: int) -> None:
        betas = torch.linspace(8.5e-4**0.5, 1.2e-2**0.5, 1000) ** 2
        self.signal = torch.sqrt((1 - betas).cumprod(dim=0))       # cumulative_scale_factors
        self.noise = torch.sqrt(1 - self.signal**2)                  # noise_std
        self.timesteps = (torch.arange(num_inference_steps) * (1000 // num_inference_steps) + 1).flip(0)
        self.n = num_inference_steps

    def add_noise(self, x: Tensor, noise: Tensor, step: int) -> Tensor:
        t = self.timesteps[step]
        return self.signal[t] * x + self.noise[t] * noise

    def remove_noise(self, x: Tensor, noise: Tensor, step: int) -> Tensor:
        t = self.timesteps[step]
        return (x - self.noise[t] * noise) / self.signal[t]

    def update(self, x: Tensor, eps: Tensor, step: int) -> Tensor:
        last = step == self.n - 1
        a_t = self.signal[self.timesteps[step]]
        a_prev = self.signal[0] if last else self.signal[self.timesteps[step + 1]]
        x0 = (x - torch.sqrt(1 - a_t**2) * eps) / a_t
        return a_prev * x0 + (0.0 if last else torch.sqrt(1 - a_prev**2)) * eps


def gaussian_blur(x: Tensor, kernel_size: int, sigma: float) -> Tensor:
    half = (kernel_size - 1) * 0.5
    taps = torch.exp(-0.5 * (torch.linspace(-half, half, kernel_size, dtype=x.dtype) / sigma) ** 2)
    taps = taps / taps.sum()
    window = (taps[:, None] * taps[None, :]).expand(x.shape[1], 1, kernel_size, kernel_size)
    p = kernel_size // 2
    return F.conv2d(F.pad(x, (p, p, p, p), mode="reflect"), window, groups=x.shape[1])


def sag_mask(attention_map: Tensor, map_size: tuple[int, int], latents: Tensor) -> Tensor:
    """``attention_map``: [2B, heads, S, S] of the guided pass; the unconditional half decides."""
    probs = attention_map.chunk(2)[0]
    b, c, h, w = latents.shape
    attended = (probs.mean(dim=1).sum(dim=1) > 1.0).reshape(b, 1, *map_size).to(probs.dtype).expand(b, c, *map_size)
    return F.interpolate(attended, size=(h, w), mode="nearest")


def denoise_step(
    unet: Callable[[Tensor, Tensor, bool], Tensor], schedule: DDIMSchedule, x: Tensor, step: int, condition_scale: float,
    sag_scale: float, kernel_size: int = 9, sigma: float = 1.0,
) -> Tensor:
    """``unet(latents, timestep, guided)``: guided = the doubled (uncond | cond) batch with the full conditioning,
    otherwise the unconditional half only.  DDIM does not rescale the model input."""
    timestep = schedule.timesteps[step].unsqueeze(0)
    ounet.middle_probe = {}
    try:
        uncond, cond = unet(torch.cat((x, x)), timestep, True).chunk(2)
        probe = ounet.middle_probe
    finally:
        ounet.middle_probe = None
    eps = uncond + condition_scale * (cond - uncond)
    mask = sag_mask(probe["map"], probe["shape"], x)
    clean = schedule.remove_noise(x, uncond, step)
    degraded = schedule.add_noise(gaussian_blur(clean, kernel_size, sigma) * mask + clean * (1 - mask), uncond, step)
    eps = eps + sag_scale * (uncond - unet(degraded, timestep, False))
    return schedule.update(x, eps, step)
