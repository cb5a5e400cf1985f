from refiners_b200.foundationals.latent_diffusion.solvers.ddim import DDIM
from refiners_b200.foundationals.latent_diffusion.solvers.euler import Euler
from refiners_b200.foundationals.latent_diffusion.solvers.solver import (
    ModelPredictionType,
    NoiseSchedule,
    Solver,
    SolverParams,
    TimestepSpacing,
)

__all__ = ["Solver", "SolverParams", "DDIM", "Euler", "NoiseSchedule", "TimestepSpacing", "ModelPredictionType"]
