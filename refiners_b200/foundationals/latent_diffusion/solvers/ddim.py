"""DDIM solver (deterministic, eta = 0).

Follows /root/reference/src/refiners/foundationals/latent_diffusion/solvers/ddim.py:15-95:
leading timestep spacing with offset 1 by default.
"""

from __future__ import annotations

import dataclasses

import torch
from torch import Generator, Tensor

from refiners_b200.foundationals.latent_diffusion.solvers.solver import (
    BaseSolverParams,
    ModelPredictionType,
    Solver,
    TimestepSpacing,
)


class DDIM(Solver):
    default_params = dataclasses.replace(
        Solver.default_params,
        timesteps_spacing=TimestepSpacing.LEADING,
        timesteps_offset=1,
    )

    def __init__(
        self, num_inference_steps: int, first_inference_step: int = 0, params: BaseSolverParams | None = None,
        device: torch.device | str = "cpu", dtype: torch.dtype = torch.float32,
    ) -> None:
        if params and params.model_prediction_type not in (ModelPredictionType.NOISE, None):
            raise NotImplementedError
        if params and params.sde_variance != 0.0:
            raise NotImplementedError("DDIM does not support sde_variance != 0.0 yet")
        super().__init__(
            num_inference_steps=num_inference_steps,
            first_inference_step=first_inference_step,
            params=params,
            device=device,
            dtype=dtype,
        )

    def __call__(self, x: Tensor, predicted_noise: Tensor, step: int, generator: Generator | None = None) -> Tensor:
        assert self.first_inference_step <= step < self.num_inference_steps, f"invalid step {step}"
        last = step == self.num_inference_steps - 1
        t_now = self.timesteps[step]
        t_next = torch.tensor([0], device=self.device, dtype=self.dtype) if last else self.timesteps[step + 1]
        a_now = self.cumulative_scale_factors[t_now]
        a_next = self.cumulative_scale_factors[t_next] if t_next > 0 else self.cumulative_scale_factors[0]
        x0 = (x - torch.sqrt(1 - a_now**2) * predicted_noise) / a_now
        # no noise re-injection at the last step (avoids visual artefacts)
        noise_factor = 0 if last else torch.sqrt(1 - a_next**2)
        return a_next * x0 + noise_factor * predicted_noise
