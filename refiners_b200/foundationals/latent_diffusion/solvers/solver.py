"""Noise schedules and the solver interface of the denoising loop.

Public contract (names, constructor signatures, attribute names, numerics to the last bit - the schedule tables are
compared with ``torch.equal`` against the reference's in tests/test_models_golden.py::test_euler_host) from
/root/reference/src/refiners/foundationals/latent_diffusion/solvers/solver.py (`SolverParams` :63-120, `Solver` :113-435).

Organisation (this module's own): the three parameter records are generated from ONE field table; a solver owns a set of
named schedule tables built by ``noise_tables`` and moved / cast as a group; timestep spacings are a dispatch table of
small generators.
"""

from __future__ import annotations

import dataclasses
from abc import ABC, abstractmethod
from enum import Enum
from typing import Any, Callable, TypeVar

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


# ------------------------------------------------------------------------------------ parameter records
# (field, resolved type, library default).  `sigma_schedule` stays optional even when resolved; `sde_variance` is the
# only field a user record carries a concrete default for.
_FIELDS: tuple[tuple[str, Any, Any], ...] = (
    ("num_train_timesteps", int, 1000),
    ("timesteps_spacing", TimestepSpacing, TimestepSpacing.LINSPACE),
    ("timesteps_offset", int, 0),
    ("initial_diffusion_rate", float, 8.5e-4),
    ("final_diffusion_rate", float, 1.2e-2),
    ("noise_schedule", NoiseSchedule, NoiseSchedule.QUADRATIC),
    ("sigma_schedule", NoiseSchedule | None, None),
    ("model_prediction_type", ModelPredictionType, ModelPredictionType.NOISE),
    ("sde_variance", float, 0.0),
)
_OPTIONAL_WHEN_RESOLVED = {"sigma_schedule"}
_CONCRETE_IN_USER_RECORD = {"sde_variance"}


def _record(name: str, fields: list[tuple[Any, ...]], bases: tuple[type, ...] = ()) -> type:
    cls = dataclasses.make_dataclass(name, fields, bases=bases, kw_only=True, frozen=True)
    cls.__module__ = __name__
    return cls


BaseSolverParams = _record(
    "BaseSolverParams",
    [(f, t if f in _CONCRETE_IN_USER_RECORD else (t | None)) for f, t, _ in _FIELDS],
)
BaseSolverParams.__doc__ = "Common shape of the solver parameter records."
# user-facing: every field optional, ``None`` = "take the solver's default"
SolverParams = _record(
    "SolverParams",
    [(f, t if f in _CONCRETE_IN_USER_RECORD else (t | None), dataclasses.field(default=d if f in _CONCRETE_IN_USER_RECORD else None))
     for f, t, d in _FIELDS],
    bases=(BaseSolverParams,),
)
SolverParams.__doc__ = "User-facing parameters; ``None`` means \"use the solver's default\"."
# what a constructed solver holds: every field concrete
ResolvedSolverParams = _record(
    "ResolvedSolverParams",
    [(f, t) for f, t, _ in _FIELDS],
    bases=(BaseSolverParams,),
)
ResolvedSolverParams.__doc__ = "Fully resolved parameters of a constructed solver."

LIBRARY_DEFAULTS = ResolvedSolverParams(**{f: d for f, _, d in _FIELDS})


# ------------------------------------------------------------------------------------------- schedules
_POWER_OF = {NoiseSchedule.UNIFORM: 1, NoiseSchedule.QUADRATIC: 2, NoiseSchedule.KARRAS: 7}
# tables every solver owns; they follow the solver's (device, dtype).  `timesteps` follows the device only: it
# indexes the tables / feeds the sinusoidal embedding and must keep its own dtype.
_TABLES = ("scale_factors", "cumulative_scale_factors", "noise_std", "signal_to_noise_ratios")


def diffusion_rates(initial: float, final: float, count: int, power: float) -> Tensor:
    """``count`` rates from ``initial`` to ``final``, equally spaced in the 1/power domain."""
    root = 1 / power
    return torch.linspace(start=initial**root, end=final**root, steps=count) ** power


def noise_tables(scale_factors: Tensor) -> dict[str, Tensor]:
    """The per-training-timestep tables derived from the per-step signal scale factors (1 - beta_t)."""
    kept = scale_factors.cumprod(dim=0)                      # alpha-bar
    signal, noise = torch.sqrt(kept), torch.sqrt(1.0 - kept)
    return {
        "scale_factors": scale_factors,
        "cumulative_scale_factors": signal,
        "noise_std": noise,
        "signal_to_noise_ratios": torch.log(signal) - torch.log(noise),
    }


def _linspace_steps(count: int, train: int, offset: int) -> Tensor:
    return torch.tensor(np.linspace(offset, train - 1 + offset, count), dtype=torch.float32).flip(0)


def _rounded_linspace_steps(count: int, train: int, offset: int) -> Tensor:
    return torch.tensor(np.linspace(offset, train - 1 + offset, count).round().astype(int)).flip(0)


def _leading_steps(count: int, train: int, offset: int) -> Tensor:
    return (torch.arange(0, count, 1) * (train // count) + offset).flip(0)


def _trailing_steps(count: int, train: int, offset: int) -> Tensor:
    return torch.arange(train - 1 + offset, offset, -(train // count))


_SPACINGS: dict[TimestepSpacing, Callable[[int, int, int], Tensor]] = {
    TimestepSpacing.LINSPACE: _linspace_steps,
    TimestepSpacing.LINSPACE_ROUNDED: _rounded_linspace_steps,
    TimestepSpacing.LEADING: _leading_steps,
    TimestepSpacing.TRAILING: _trailing_steps,
}


class Solver(fl.Module, ABC):
    """A sampler of the reverse diffusion: ``solver(x, predicted_noise=..., step=...)`` -> the next latents."""

    timesteps: Tensor
    params: ResolvedSolverParams  # type: ignore[valid-type]
    default_params = LIBRARY_DEFAULTS

    def __init__(
        self, num_inference_steps: int, first_inference_step: int = 0, params: BaseSolverParams | None = None,  # type: ignore[valid-type]
        device: Device | str = "cpu", dtype: DType = torch.float32,
    ) -> None:
        super().__init__()
        self.num_inference_steps = num_inference_steps
        self.first_inference_step = first_inference_step
        self.params = self.resolve_params(params)
        for name, table in noise_tables(self.sample_noise_schedule()).items():
            setattr(self, name, table)
        self.timesteps = self._generate_timesteps()
        self.to(device=device, dtype=dtype)

    # -- parameters -----------------------------------------------------------------------------
    def resolve_params(self, params: BaseSolverParams | None) -> ResolvedSolverParams:  # type: ignore[valid-type]
        """The class defaults overridden by every field the caller actually set."""
        chosen = {} if params is None else {f.name: getattr(params, f.name) for f in dataclasses.fields(params)}
        return dataclasses.replace(self.default_params, **{k: v for k, v in chosen.items() if v is not None})

    def rebuild(self: T, num_inference_steps: int | None, first_inference_step: int | None = None) -> T:
        """A fresh solver of the same class, parameters, device and dtype with another step count / first step."""
        keep = lambda new, old: old if new is None else new  # noqa: E731
        return type(self)(
            num_inference_steps=keep(num_inference_steps, self.num_inference_steps),
            first_inference_step=keep(first_inference_step, self.first_inference_step),
            params=dataclasses.replace(self.params),
            device=self.device,
            dtype=self.dtype,
        )

    # -- schedules ------------------------------------------------------------------------------
    def sample_power_distribution(self, power: float = 2, /) -> Tensor:
        p = self.params
        return diffusion_rates(p.initial_diffusion_rate, p.final_diffusion_rate, p.num_train_timesteps, power)

    def sample_noise_schedule(self) -> Tensor:
        return 1 - self.sample_power_distribution(_POWER_OF[self.params.noise_schedule])

    @staticmethod
    def generate_timesteps(
        spacing: TimestepSpacing, num_inference_steps: int, num_train_timesteps: int = 1000, offset: int = 0,
    ) -> Tensor:
        if spacing not in _SPACINGS:
            raise RuntimeError("generate_timesteps called with custom spacing")
        return _SPACINGS[spacing](num_inference_steps, num_train_timesteps, offset)

    def _generate_timesteps(self) -> Tensor:
        p = self.params
        return self.generate_timesteps(p.timesteps_spacing, self.num_inference_steps, p.num_train_timesteps, p.timesteps_offset)

    # -- the step -------------------------------------------------------------------------------
    @abstractmethod
    def __call__(self, x: Tensor, predicted_noise: Tensor, step: int, generator: Generator | None = None) -> Tensor: ...

    def scale_model_input(self, x: Tensor, step: int) -> Tensor:
        """What the model must be fed for ``step`` (identity unless a solver works in sigma space)."""
        return x

    def _signal_and_noise(self, step: int) -> tuple[Tensor, Tensor]:
        at = self.timesteps[step]
        return self.cumulative_scale_factors[at], self.noise_std[at]

    def _add_noise(self, x: Tensor, noise: Tensor, step: int) -> Tensor:
        signal, sigma = self._signal_and_noise(step)
        return signal * x + sigma * noise

    def add_noise(self, x: Tensor, noise: Tensor, step: int | list[int]) -> Tensor:
        """Forward diffusion of ``x`` to inference step ``step`` (one step per batch row if a list is given)."""
        if not isinstance(step, list):
            return self._add_noise(x, noise, step)
        assert len(x) == len(noise) == len(step), "x, noise, and step must have the same length"
        return torch.stack([self._add_noise(xi, ni, si) for xi, ni, si in zip(x, noise, step)], dim=0)

    def remove_noise(self, x: Tensor, noise: Tensor, step: int) -> Tensor:
        signal, sigma = self._signal_and_noise(step)
        return (x - sigma * noise) / signal

    # -- bookkeeping ----------------------------------------------------------------------------
    @property
    def all_steps(self) -> list[int]:
        return [*range(self.num_inference_steps)]

    @property
    def inference_steps(self) -> list[int]:
        return [*range(self.first_inference_step, self.num_inference_steps)]

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

    def to(self, device: Device | str | None = None, dtype: DType | None = None) -> "Solver":  # type: ignore[override]
        """Move / cast every schedule tensor the solver holds (sub-classes add theirs as plain attributes);
        ``timesteps`` only moves."""
        super().to(device=device, dtype=dtype)
        held = [name for name, value in vars(self).items() if isinstance(value, Tensor)]
        for name in held:
            moved = getattr(self, name).to(device=device, dtype=None if name == "timesteps" else dtype)
            setattr(self, name, moved)
        return self
