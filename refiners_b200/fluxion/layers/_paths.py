"""Import-path compatibility with the reference's layout of ``fluxion.layers``.

The reference spreads its layers over fourteen modules (``chain``, ``basics``, ``linear``, ``conv``, ``norm`` ...);
this package groups them differently (``graph``: the Chain machinery, ``leaves``: weighted / elementwise leaves on the
CUDA kernels, ``shape_ops``: views and small helpers, ``composites``: chains built from them).  A user switching over
keeps imports such as ``from refiners.fluxion.layers.chain import ChainError, Distribute`` working: every reference module
path is registered here as a module object exposing the same public names."""

from __future__ import annotations

import sys
from types import ModuleType

# reference module -> public names it defines (found in this package's namespace)
_LAYOUT: dict[str, tuple[str, ...]] = {
    "activations": ("Activation", "SiLU", "ReLU", "GeLUApproximation", "GeLU", "Sigmoid", "GLU"),
    "attentions": ("scaled_dot_product_attention", "scaled_dot_product_attention_non_optimized", "ScaledDotProductAttention",
                   "Attention", "SelfAttention", "SelfAttention2d"),
    "basics": ("Identity", "GetArg", "Flatten", "Unflatten", "Reshape", "Transpose", "Permute", "Slicing", "Squeeze",
               "Unsqueeze", "Sin", "Cos", "Multiply", "Parameter"),
    "chain": ("generate_unique_names", "structural_copy", "ChainError", "Chain", "UseContext", "SetContext", "Lambda",
              "Parallel", "Distribute", "Passthrough", "Sum", "Residual", "Concatenate", "Matmul", "ReturnException",
              "Return", "Breakpoint"),
    "conv": ("Conv2d", "ConvTranspose2d"),
    "converter": ("Converter",),
    "embedding": ("Embedding",),
    "linear": ("Linear", "MultiLinear"),
    "maxpool": ("MaxPool1d", "MaxPool2d"),
    "module": ("Module", "ContextModule", "WeightedModule", "ModuleTree"),
    "norm": ("LayerNorm", "GroupNorm", "LayerNorm2d", "InstanceNorm2d"),
    "padding": ("ReflectionPad2d",),
    "pixelshuffle": ("PixelUnshuffle",),
    "sampling": ("Interpolate", "Downsample", "Upsample"),
}


def register(package: ModuleType, sources: tuple[ModuleType, ...]) -> None:
    for name, public in _LAYOUT.items():
        full = f"{package.__name__}.{name}"
        if full in sys.modules:
            continue
        mod = ModuleType(full, f"Reference-layout view of {package.__name__} (see _paths.py).")
        for attr in public:
            owner = next(src for src in sources if hasattr(src, attr))
            setattr(mod, attr, getattr(owner, attr))
        mod.__all__ = list(public)  # type: ignore[attr-defined]
        sys.modules[full] = mod
        setattr(package, name, mod)
