"""DDIM solver (deterministic, eta = 0).

Contract from /root/reference/src/refiners/foundationals/latent_diffusion/solvers/ddim.py:15-95: leading timestep
spacing with offset 1 by default; noise prediction only; no noise re-injection on the last step.
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
    default_params = dataclasses.replace(Solver.default_params, timesteps_spacing=TimestepSpacing.LEADING, timesteps_offset=1)

    def __init__(
        self, num_inference_steps: int, first_inference_step: int = 0, params: "BaseSolverParams | None" = None,  # type: ignore[valid-type]
        device: torch.device | str = "cpu", dtype: torch.dtype = torch.float32,
    ) -> None:
        if params is not None:
            if params.model_prediction_type not in (None, ModelPredictionType.NOISE):
                raise NotImplementedError
            if params.sde_variance != 0.0:
                raise NotImplementedError("DDIM does not support sde_variance != 0.0 yet")
        super().__init__(num_inference_steps, first_inference_step, params=params, device=device, dtype=dtype)

    def _signal_scale_after(self, step: int) -> Tensor:
        """sqrt(alpha-bar) at the timestep the update lands on: the next inference timestep, or training timestep 0
        after the last step (also when the next timestep is not positive)."""
        if step == self.num_inference_steps - 1:
            return self.cumulative_scale_factors[0]
        landing = self.timesteps[step + 1]
        return self.cumulative_scale_factors[landing] if landing > 0 else self.cumulative_scale_factors[0]

    def __call__(self, x: Tensor, predicted_noise: Tensor, step: int, generator: Generator | None = None) -> Tensor:
        assert self.first_inference_step <= step < self.num_inference_steps, f"invalid step {step}"
        signal_now = self.cumulative_scale_factors[self.timesteps[step]]
        signal_next = self._signal_scale_after(step)
        clean = (x - torch.sqrt(1 - signal_now**2) * predicted_noise) / signal_now
        if step == self.num_inference_steps - 1:
            return signal_next * clean + 0 * predicted_noise  # the final step lands on the clean estimate
        return signal_next * clean + torch.sqrt(1 - signal_next**2) * predicted_noise
