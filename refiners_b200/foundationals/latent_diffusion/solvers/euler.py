"""Euler solver in sigma space (k-diffusion style): ``x <- x + eps * (sigma[next] - sigma[now])``.

Contract and numerics (bit-exact tables, see tests/test_models_golden.py::test_euler_host) from
/root/reference/src/refiners/foundationals/latent_diffusion/solvers/euler.py:13-100.
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


def _reject_unsupported(params: "BaseSolverParams | None") -> None:  # type: ignore[valid-type]
    if params is None:
        return
    if params.noise_schedule not in (None, NoiseSchedule.QUADRATIC):
        raise NotImplementedError
    if params.sde_variance != 0.0:
        raise NotImplementedError("Euler does not support sde_variance != 0.0 yet")


class Euler(Solver):
    """``sigmas`` has one entry per inference step plus a trailing 0 (the clean sample)."""

    def __init__(
        self, num_inference_steps: int, first_inference_step: int = 0, params: "BaseSolverParams | None" = None,  # type: ignore[valid-type]
        device: torch.device | str = "cpu", dtype: torch.dtype = torch.float32,
    ) -> None:
        _reject_unsupported(params)
        super().__init__(num_inference_steps, first_inference_step, params=params, device=device, dtype=dtype)
        self.sigmas = self._generate_sigmas()

    def _generate_sigmas(self) -> Tensor:
        """sigma = noise_std / signal scale, linearly interpolated (in float64) at the - possibly fractional -
        inference timesteps, then a final 0."""
        # .float(): a solver built directly in a 16-bit dtype (``rebuild`` of a cast solver) still interpolates in
        # float64 from float32 values - the reference raises there (numpy has no bfloat16)
        per_train_step = (self.noise_std / self.cumulative_scale_factors).float().cpu().numpy()
        at_inference_steps = np.interp(self.timesteps.cpu().numpy(), np.arange(per_train_step.size), per_train_step)
        return torch.from_numpy(np.append(at_inference_steps, 0.0)).to(device=self.device, dtype=self.dtype)

    @property
    def init_noise_sigma(self) -> Tensor:
        return self.sigmas.max()

    def scale_model_input(self, x: Tensor, step: int) -> Tensor:
        """``step == -1`` scales pure noise up to the first sigma; otherwise the variance-preserving input scaling."""
        if step < 0:
            assert step == -1, f"invalid step {step}"
            return x * self.init_noise_sigma
        sigma = self.sigmas[step]
        return x / ((sigma**2 + 1) ** 0.5)

    def __call__(self, x: Tensor, predicted_noise: Tensor, step: int, generator: Generator | None = None) -> Tensor:
        assert self.first_inference_step <= step < self.num_inference_steps, f"invalid step {step}"
        now, nxt = self.sigmas[step], self.sigmas[step + 1]
        kind = self.params.model_prediction_type
        if kind == ModelPredictionType.NOISE:
            return x + predicted_noise * (nxt - now)
        assert kind == ModelPredictionType.SAMPLE
        keep = nxt / now  # the model predicted the clean sample: move towards it along the sigma ratio
        return keep * x + (1 - keep) * predicted_noise
