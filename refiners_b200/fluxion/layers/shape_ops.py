"""Shape/glue leaves (no arithmetic worth a kernel): views, slicing, scaling, converter.

API follows /root/reference/src/refiners/fluxion/layers/basics.py:8-445,
converter.py:6-48, padding.py, pixelshuffle.py, maxpool.py.  These stay ATen views/ops on
both devices - they move no significant bytes on the denoising path.
"""

from __future__ import annotations

import torch
from torch import Size, Tensor, nn

from refiners_b200.fluxion.layers.base import ContextModule, Module, WeightedModule


class Identity(Module):
    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        return x


class GetArg(Module):
    def __init__(self, index: int) -> None:
        super().__init__()
        self.index = index

    def forward(self, *args: Tensor) -> Tensor:
        return args[self.index]


class Flatten(Module):
    def __init__(self, start_dim: int = 0, end_dim: int = -1) -> None:
        super().__init__()
        self.start_dim = start_dim
        self.end_dim = end_dim

    def forward(self, x: Tensor) -> Tensor:
        return x.flatten(self.start_dim, self.end_dim)


class Unflatten(Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, x: Tensor, sizes: Size) -> Tensor:
        return x.unflatten(self.dim, sizes)


class Reshape(Module):
    """Reshape everything but the batch dimension."""

    def __init__(self, *shape: int) -> None:
        super().__init__()
        self.shape = shape

    def forward(self, x: Tensor) -> Tensor:
        return x.reshape(x.shape[0], *self.shape)


class Transpose(Module):
    def __init__(self, dim0: int, dim1: int) -> None:
        super().__init__()
        self.dim0 = dim0
        self.dim1 = dim1

    def forward(self, x: Tensor) -> Tensor:
        return x.transpose(self.dim0, self.dim1)


class Permute(Module):
    def __init__(self, *dims: int) -> None:
        super().__init__()
        self.dims = dims

    def forward(self, x: Tensor) -> Tensor:
        return x.permute(*self.dims)


class Slicing(Module):
    """``x[..., start:end:step, ...]`` along ``dim`` with python-style clamping; the result is a
    copy (index_select), as in the reference (basics.py:266-303)."""

    def __init__(self, dim: int = 0, start: int = 0, end: int | None = None, step: int = 1) -> None:
        super().__init__()
        self.dim = dim
        self.start = start
        self.end = end
        self.step = step

    def forward(self, x: Tensor) -> Tensor:
        n = x.shape[self.dim]
        lo = self.start + n if self.start < 0 else self.start
        lo = min(max(lo, 0), n)
        hi = self.end or n
        hi = hi + n if hi < 0 else hi
        hi = min(max(hi, 0), n)
        if lo >= hi:
            shape = list(x.shape)
            shape[self.dim] = 0
            return torch.empty(*shape, device=x.device)
        index = torch.arange(lo, hi, self.step, device=x.device)
        return x.index_select(self.dim, index)


class Squeeze(Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, x: Tensor) -> Tensor:
        return x.squeeze(self.dim)


class Unsqueeze(Module):
    def __init__(self, dim: int) -> None:
        super().__init__()
        self.dim = dim

    def forward(self, x: Tensor) -> Tensor:
        return x.unsqueeze(self.dim)


class Sin(Module):
    def forward(self, x: Tensor) -> Tensor:
        return torch.sin(x)


class Cos(Module):
    def forward(self, x: Tensor) -> Tensor:
        return torch.cos(x)


class Multiply(Module):
    def __init__(self, scale: float = 1.0, bias: float = 0.0) -> None:
        super().__init__()
        self.scale = scale
        self.bias = bias

    def forward(self, x: Tensor) -> Tensor:
        return self.scale * x + self.bias


class Parameter(WeightedModule):
    """A learnable tensor broadcast over the batch of its input."""

    def __init__(
        self,
        *dims: int,
        requires_grad: bool = True,
        device: torch.device | str | None = None,
        dtype: torch.dtype | None = None,
    ) -> None:
        super().__init__()
        self.dims = dims
        self.weight = nn.Parameter(torch.randn(*dims, device=device, dtype=dtype), requires_grad=requires_grad)

    def forward(self, x: Tensor) -> Tensor:
        return self.weight.expand(x.shape[0], *self.dims)

    @property
    def requires_grad(self) -> bool:
        return self.weight.requires_grad

    @requires_grad.setter
    def requires_grad(self, value: bool) -> None:
        self.weight.requires_grad = value


class Converter(ContextModule):
    """Cast inputs to the parent's device and/or dtype (reference converter.py:6-48)."""

    def __init__(self, set_device: bool = True, set_dtype: bool = True) -> None:
        super().__init__()
        self.set_device = set_device
        self.set_dtype = set_dtype

    def forward(self, *inputs: Tensor) -> tuple[Tensor, ...]:
        parent = self.ensure_parent
        device = dtype = None
        if self.set_device:
            device = parent.device
            assert device is not None, "parent has no device"
        if self.set_dtype:
            dtype = parent.dtype
            assert dtype is not None, "parent has no dtype"
        return tuple(x.to(device=device, dtype=dtype) for x in inputs)

    def __repr__(self) -> str:
        return f"{type(self).__name__}(set_device={self.set_device}, set_dtype={self.set_dtype})"


class ReflectionPad2d(nn.ReflectionPad2d, Module):
    def __init__(self, padding: int) -> None:
        nn.ReflectionPad2d.__init__(self, padding)


class PixelUnshuffle(nn.PixelUnshuffle, Module):
    def __init__(self, downscale_factor: int) -> None:
        nn.PixelUnshuffle.__init__(self, downscale_factor)


class MaxPool1d(nn.MaxPool1d, Module):
    def __init__(
        self,
        kernel_size: int,
        stride: int | None = None,
        padding: int = 0,
        dilation: int = 1,
        return_indices: bool = False,
        ceil_mode: bool = False,
    ) -> None:
        nn.MaxPool1d.__init__(self, kernel_size, stride, padding, dilation, return_indices, ceil_mode)


class MaxPool2d(nn.MaxPool2d, Module):
    def __init__(
        self,
        kernel_size: int | tuple[int, int],
        stride: int | tuple[int, int] | None = None,
        padding: int | tuple[int, int] = (0, 0),
        dilation: int | tuple[int, int] = (1, 1),
        return_indices: bool = False,
        ceil_mode: bool = False,
    ) -> None:
        nn.MaxPool2d.__init__(self, kernel_size, stride, padding, dilation, return_indices, ceil_mode)
