"""``fl`` namespace: same exported names as refiners.fluxion.layers
(/root/reference/src/refiners/fluxion/layers/__init__.py:58-117)."""

from refiners_b200.fluxion.layers.base import ContextModule, Module, ModuleTree, WeightedModule
from refiners_b200.fluxion.layers.composites import (
    Attention,
    Downsample,
    Interpolate,
    MultiLinear,
    SelfAttention,
    SelfAttention2d,
    Upsample,
)
from refiners_b200.fluxion.layers.graph import (
    Breakpoint,
    Chain,
    ChainError,
    Concatenate,
    Distribute,
    Lambda,
    Matmul,
    Parallel,
    Passthrough,
    Residual,
    Return,
    ReturnException,
    SetContext,
    Sum,
    UseContext,
    generate_unique_names,
    structural_copy,
)
from refiners_b200.fluxion.layers.leaves import (
    GLU,
    Activation,
    Conv2d,
    ConvTranspose2d,
    Embedding,
    GeLU,
    GeLUApproximation,
    GroupNorm,
    InstanceNorm2d,
    LayerNorm,
    LayerNorm2d,
    Linear,
    ReLU,
    ScaledDotProductAttention,
    Sigmoid,
    SiLU,
)
from refiners_b200.fluxion.layers.shape_ops import (
    Converter,
    Cos,
    Flatten,
    GetArg,
    Identity,
    MaxPool1d,
    MaxPool2d,
    Multiply,
    Parameter,
    Permute,
    PixelUnshuffle,
    ReflectionPad2d,
    Reshape,
    Sin,
    Slicing,
    Squeeze,
    Transpose,
    Unflatten,
    Unsqueeze,
)

__all__ = [
    "Embedding", "LayerNorm", "GroupNorm", "LayerNorm2d", "InstanceNorm2d", "Activation", "GeLU",
    "GeLUApproximation", "GLU", "SiLU", "ReLU", "Sigmoid", "Attention", "ScaledDotProductAttention",
    "SelfAttention", "SelfAttention2d", "Identity", "GetArg", "Flatten", "Unflatten", "Transpose", "Permute",
    "Squeeze", "Unsqueeze", "Reshape", "Slicing", "Parameter", "Sin", "Cos", "Multiply", "Matmul", "Lambda",
    "Return", "Sum", "Residual", "Chain", "UseContext", "SetContext", "Parallel", "Distribute", "Passthrough",
    "Breakpoint", "Concatenate", "Conv2d", "ConvTranspose2d", "Linear", "MultiLinear", "Downsample", "Upsample",
    "Module", "WeightedModule", "ContextModule", "Interpolate", "ReflectionPad2d", "PixelUnshuffle", "Converter",
    "MaxPool1d", "MaxPool2d", "ChainError", "ModuleTree",
]

# reference-layout import paths (refiners.fluxion.layers.chain, .basics, .linear ...)
import sys as _sys

from refiners_b200.fluxion.layers import _paths, base as _base, composites as _composites, graph as _graph, leaves as _leaves, shape_ops as _shape_ops

_paths.register(_sys.modules[__name__], (_graph, _base, _leaves, _shape_ops, _composites))
