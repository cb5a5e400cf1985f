"""Euler solver (k-diffusion style): x_{t+1} = x_t + eps * (sigma_{t+1} - sigma_t).

Follows /root/reference/src/refiners/foundationals/latent_diffusion/solvers/euler.py:13-100.
"""

from __future__ import annotations

import numpy as np
import torch
from torch import Generator, Tensor

from refiners_b200.foundationals.latent_diffusion.solvers.solver import (
    BaseSolverParams,
    ModelPredictionType,
    NoiseSchedule,
    Solver,
)


class Euler(Solver):
    def __init__(
        self, num_inference_steps: int, first_inference_step: int = 0, params: BaseSolverParams | None = None,
        device: torch.device | str = "cpu", dtype: torch.dtype = torch.float32,
    ) -> None:
        if params and params.noise_schedule not in (NoiseSchedule.QUADRATIC, None):
            raise NotImplementedError
        if params and params.sde_variance != 0.0:
            raise NotImplementedError("Euler does not support sde_variance != 0.0 yet")
        super().__init__(
            num_inference_steps=num_inference_steps,
            first_inference_step=first_inference_step,
            params=params,
            device=device,
            dtype=dtype,
        )
        self.sigmas = self._generate_sigmas()

    @property
    def init_noise_sigma(self) -> Tensor:
        return self.sigmas.max()

    def _generate_sigmas(self) -> Tensor:
        table = (self.noise_std / self.cumulative_scale_factors).cpu()
        at_steps = torch.tensor(np.interp(self.timesteps.cpu(), np.arange(0, len(table)), table))
        return torch.cat([at_steps, torch.tensor([0.0])]).to(device=self.device, dtype=self.dtype)

    def scale_model_input(self, x: Tensor, step: int) -> Tensor:
        if step == -1:
            return x * self.init_noise_sigma
        return x / ((self.sigmas[step] ** 2 + 1) ** 0.5)

    def __call__(self, x: Tensor, predicted_noise: Tensor, step: int, generator: Generator | None = None) -> Tensor:
        assert self.first_inference_step <= step < self.num_inference_steps, f"invalid step {step}"
        if self.params.model_prediction_type == ModelPredictionType.SAMPLE:
            ratio = self.sigmas[step + 1] / self.sigmas[step]
            return ratio * x + (1 - ratio) * predicted_noise  # the model predicted x0
        assert self.params.model_prediction_type == ModelPredictionType.NOISE
        return x + predicted_noise * (self.sigmas[step + 1] - self.sigmas[step])
