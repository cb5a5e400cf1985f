"""Noise schedules and the solver interface.

Semantics follow /root/reference/src/refiners/foundationals/latent_diffusion/solvers/solver.py
(`SolverParams` :63-120, `Solver` :113-435): quadratic beta schedule by default, linspace
timesteps, all schedule tensors cast to the model dtype (timesteps keep theirs).
"""

from __future__ import annotations

import dataclasses
from abc import ABC, abstractmethod
from enum import Enum
from typing import TypeVar

import numpy as np
import torch
from torch import Generator, Tensor

from refiners_b200.fluxion import layers as fl

T = TypeVar("T", bound="Solver")
Device = torch.device
DType = torch.dtype


class NoiseSchedule(str, Enum):
    UNIFORM = "uniform"
    QUADRATIC = "quadratic"
    KARRAS = "karras"


class TimestepSpacing(str, Enum):
    LINSPACE = "linspace"
    LINSPACE_ROUNDED = "linspace_rounded"
    LEADING = "leading"
    TRAILING = "trailing"
    CUSTOM = "custom"


class ModelPredictionType(str, Enum):
    NOISE = "noise"
    SAMPLE = "sample"


@dataclasses.dataclass(kw_only=True, frozen=True)
class BaseSolverParams:
    num_train_timesteps: int | None
    timesteps_spacing: TimestepSpacing | None
    timesteps_offset: int | None
    initial_diffusion_rate: float | None
    final_diffusion_rate: float | None
    noise_schedule: NoiseSchedule | None
    sigma_schedule: NoiseSchedule | None
    model_prediction_type: ModelPredictionType | None
    sde_variance: float


@dataclasses.dataclass(kw_only=True, frozen=True)
class SolverParams(BaseSolverParams):
    """User-facing parameters; ``None`` means "use the solver's default"."""

    num_train_timesteps: int | None = None
    timesteps_spacing: TimestepSpacing | None = None
    timesteps_offset: int | None = None
    initial_diffusion_rate: float | None = None
    final_diffusion_rate: float | None = None
    noise_schedule: NoiseSchedule | None = None
    sigma_schedule: NoiseSchedule | None = None
    model_prediction_type: ModelPredictionType | None = None
    sde_variance: float = 0.0


@dataclasses.dataclass(kw_only=True, frozen=True)
class ResolvedSolverParams(BaseSolverParams):
    num_train_timesteps: int
    timesteps_spacing: TimestepSpacing
    timesteps_offset: int
    initial_diffusion_rate: float
    final_diffusion_rate: float
    noise_schedule: NoiseSchedule
    sigma_schedule: NoiseSchedule | None
    model_prediction_type: ModelPredictionType
    sde_variance: float


_SCHEDULE_POWER = {NoiseSchedule.UNIFORM: 1, NoiseSchedule.QUADRATIC: 2, NoiseSchedule.KARRAS: 7}


class Solver(fl.Module, ABC):
    timesteps: Tensor
    params: ResolvedSolverParams

    default_params = ResolvedSolverParams(
        num_train_timesteps=1000,
        timesteps_spacing=TimestepSpacing.LINSPACE,
        timesteps_offset=0,
        initial_diffusion_rate=8.5e-4,
        final_diffusion_rate=1.2e-2,
        noise_schedule=NoiseSchedule.QUADRATIC,
        sigma_schedule=None,
        model_prediction_type=ModelPredictionType.NOISE,
        sde_variance=0.0,
    )

    def __init__(
        self, num_inference_steps: int, first_inference_step: int = 0, params: BaseSolverParams | None = None,
        device: Device | str = "cpu", dtype: DType = torch.float32,
    ) -> None:
        super().__init__()
        self.num_inference_steps, self.first_inference_step = num_inference_steps, first_inference_step
        self.params = self.resolve_params(params)
        self.scale_factors = self.sample_noise_schedule()
        alphas_cumprod = self.scale_factors.cumprod(dim=0)
        self.cumulative_scale_factors = torch.sqrt(alphas_cumprod)
        self.noise_std = torch.sqrt(1.0 - alphas_cumprod)
        self.signal_to_noise_ratios = torch.log(self.cumulative_scale_factors) - torch.log(self.noise_std)
        self.timesteps = self._generate_timesteps()
        self.to(device=device, dtype=dtype)

    def resolve_params(self, params: BaseSolverParams | None) -> ResolvedSolverParams:
        if params is None:
            return dataclasses.replace(self.default_params)
        overrides = {k: v for k, v in dataclasses.asdict(params).items() if v is not None}
        return dataclasses.replace(self.default_params, **overrides)

    @abstractmethod
    def __call__(self, x: Tensor, predicted_noise: Tensor, step: int, generator: Generator | None = None) -> Tensor: ...

    @staticmethod
    def generate_timesteps(
        spacing: TimestepSpacing, num_inference_steps: int, num_train_timesteps: int = 1000, offset: int = 0,
    ) -> Tensor:
        top = num_train_timesteps - 1 + offset
        if spacing is TimestepSpacing.LINSPACE:
            return torch.tensor(np.linspace(offset, top, num_inference_steps), dtype=torch.float32).flip(0)
        if spacing is TimestepSpacing.LINSPACE_ROUNDED:
            return torch.tensor(np.linspace(offset, top, num_inference_steps).round().astype(int)).flip(0)
        if spacing is TimestepSpacing.LEADING:
            ratio = num_train_timesteps // num_inference_steps
            return (torch.arange(0, num_inference_steps, 1) * ratio + offset).flip(0)
        if spacing is TimestepSpacing.TRAILING:
            ratio = num_train_timesteps // num_inference_steps
            return torch.arange(top, offset, -ratio)
        raise RuntimeError("generate_timesteps called with custom spacing")

    def _generate_timesteps(self) -> Tensor:
        return self.generate_timesteps(
            spacing=self.params.timesteps_spacing,
            num_inference_steps=self.num_inference_steps,
            num_train_timesteps=self.params.num_train_timesteps,
            offset=self.params.timesteps_offset,
        )

    def _add_noise(self, x: Tensor, noise: Tensor, step: int) -> Tensor:
        t = self.timesteps[step]
        return self.cumulative_scale_factors[t] * x + self.noise_std[t] * noise

    def add_noise(self, x: Tensor, noise: Tensor, step: int | list[int]) -> Tensor:
        if isinstance(step, list):
            assert len(x) == len(noise) == len(step), "x, noise, and step must have the same length"
            return torch.stack([self._add_noise(x[i], noise[i], step[i]) for i in range(x.shape[0])], dim=0)
        return self._add_noise(x, noise, step)

    def remove_noise(self, x: Tensor, noise: Tensor, step: int) -> Tensor:
        t = self.timesteps[step]
        return (x - self.noise_std[t] * noise) / self.cumulative_scale_factors[t]

    @property
    def all_steps(self) -> list[int]:
        return list(range(self.num_inference_steps))

    @property
    def inference_steps(self) -> list[int]:
        return self.all_steps[self.first_inference_step :]

    @property
    def device(self) -> Device:
        return self.scale_factors.device

    @device.setter
    def device(self, device: Device | str | None = None) -> None:
        self.to(device=device)

    @property
    def dtype(self) -> DType:
        return self.scale_factors.dtype

    @dtype.setter
    def dtype(self, dtype: DType | None = None) -> None:
        self.to(dtype=dtype)

    def rebuild(self: T, num_inference_steps: int | None, first_inference_step: int | None = None) -> T:
        return self.__class__(
            num_inference_steps=self.num_inference_steps if num_inference_steps is None else num_inference_steps,
            first_inference_step=self.first_inference_step if first_inference_step is None else first_inference_step,
            params=dataclasses.replace(self.params),
            device=self.device,
            dtype=self.dtype,
        )

    def scale_model_input(self, x: Tensor, step: int) -> Tensor:
        return x

    def sample_power_distribution(self, power: float = 2, /) -> Tensor:
        return (
            torch.linspace(
                start=self.params.initial_diffusion_rate ** (1 / power),
                end=self.params.final_diffusion_rate ** (1 / power),
                steps=self.params.num_train_timesteps,
            )
            ** power
        )

    def sample_noise_schedule(self) -> Tensor:
        return 1 - self.sample_power_distribution(_SCHEDULE_POWER[self.params.noise_schedule])

    def to(self, device: Device | str | None = None, dtype: DType | None = None) -> "Solver":  # type: ignore[override]
        super().to(device=device, dtype=dtype)
        for name, value in list(self.__dict__.items()):
            if isinstance(value, Tensor):
                # timesteps keep their dtype (they index the schedules / feed the sinusoid)
                setattr(self, name, value.to(device=device) if name == "timesteps" else value.to(device=device, dtype=dtype))
        return self
