"""Chains assembled from leaves: attention blocks, resampling, small MLPs.

Tree shapes (and therefore state-dict keys) follow the reference:
attentions.py:205-489, sampling.py:13-161, linear.py:61-128 under
/root/reference/src/refiners/fluxion/layers/.
"""

from __future__ import annotations

from torch import Size, Tensor
from torch.nn import functional as F

import torch

from refiners_b200 import backend as B
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.layers.base import Module
from refiners_b200.fluxion.layers.graph import Chain, Distribute, Lambda, Parallel, SetContext, UseContext
from refiners_b200.fluxion.layers.leaves import Conv2d, Linear, ReLU, ScaledDotProductAttention
from refiners_b200.fluxion.layers.shape_ops import Identity
from refiners_b200.fluxion.utils import interpolate

Device = torch.device
DType = torch.dtype


class Attention(Chain):
    """Distribute(Wq, Wk, Wv) -> SDPA -> Wo.  Inputs: (query, key, value) as [B, S, C]."""

    def __init__(
        self,
        embedding_dim: int,
        num_heads: int = 1,
        key_embedding_dim: int | None = None,
        value_embedding_dim: int | None = None,
        inner_dim: int | None = None,
        use_bias: bool = True,
        is_causal: bool = False,
        is_optimized: bool = True,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        assert embedding_dim % num_heads == 0, (
            f"embedding_dim {embedding_dim} must be divisible by num_heads {num_heads}"
        )
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.heads_dim = embedding_dim // num_heads
        self.key_embedding_dim = key_embedding_dim or embedding_dim
        self.value_embedding_dim = value_embedding_dim or embedding_dim
        self.inner_dim = inner_dim or embedding_dim
        self.use_bias = use_bias
        self.is_causal = is_causal
        self.is_optimized = is_optimized
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            Distribute(
                Linear(self.embedding_dim, self.inner_dim, bias=use_bias, **kw),
                Linear(self.key_embedding_dim, self.inner_dim, bias=use_bias, **kw),
                Linear(self.value_embedding_dim, self.inner_dim, bias=use_bias, **kw),
            ),
            ScaledDotProductAttention(num_heads=num_heads, is_causal=is_causal, is_optimized=is_optimized),
            Linear(self.inner_dim, self.embedding_dim, bias=True, **kw),
        )


class SelfAttention(Attention):
    """Attention whose three inputs are the same tensor (a leading Parallel of Identities)."""

    def __init__(
        self,
        embedding_dim: int,
        inner_dim: int | None = None,
        num_heads: int = 1,
        use_bias: bool = True,
        is_causal: bool = False,
        is_optimized: bool = True,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        super().__init__(
            embedding_dim=embedding_dim,
            inner_dim=inner_dim,
            num_heads=num_heads,
            use_bias=use_bias,
            is_causal=is_causal,
            is_optimized=is_optimized,
            device=device,
            dtype=dtype,
        )
        self.insert(0, Parallel(Identity(), Identity(), Identity()))


class SelfAttention2d(SelfAttention):
    """Self-attention over the pixels of an NCHW map."""

    def __init__(
        self,
        channels: int,
        num_heads: int = 1,
        use_bias: bool = True,
        is_causal: bool = False,
        is_optimized: bool = True,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        assert channels % num_heads == 0, f"channels {channels} must be divisible by num_heads {num_heads}"
        self.channels = channels
        super().__init__(
            embedding_dim=channels,
            num_heads=num_heads,
            use_bias=use_bias,
            is_causal=is_causal,
            is_optimized=is_optimized,
            device=device,
            dtype=dtype,
        )
        self.insert(0, Lambda(self._tensor_2d_to_sequence))
        self.append(Lambda(self._sequence_to_tensor_2d))

    def init_context(self) -> Contexts:
        return {"reshape": {"height": None, "width": None}}

    def _tensor_2d_to_sequence(self, x: Tensor) -> Tensor:
        height, width = x.shape[-2:]
        self.set_context("reshape", {"height": height, "width": width})
        return x.reshape(x.shape[0], x.shape[1], height * width).transpose(1, 2)

    def _sequence_to_tensor_2d(self, x: Tensor) -> Tensor:
        height, width = self.use_context("reshape").values()
        return x.transpose(1, 2).reshape(x.shape[0], x.shape[2], height, width)


class MultiLinear(Chain):
    def __init__(
        self,
        input_dim: int,
        output_dim: int,
        inner_dim: int,
        num_layers: int,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        stack: list[Module] = []
        width = input_dim
        for _ in range(num_layers - 1):
            stack += [Linear(width, inner_dim, device=device, dtype=dtype), ReLU()]
            width = inner_dim
        stack.append(Linear(inner_dim, output_dim, device=device, dtype=dtype))
        super().__init__(stack)


class Interpolate(Module):
    def __init__(self, mode: str = "nearest", antialias: bool = False) -> None:
        super().__init__()
        self.mode = mode
        self.antialias = antialias

    def forward(self, x: Tensor, shape: Size) -> Tensor:
        if self.mode == "nearest" and len(shape) == 2 and B.resize_nearest_supported(x):
            return B.resize_nearest(x, int(shape[0]), int(shape[1]))
        return interpolate(x, size=shape, mode=self.mode, antialias=self.antialias)


class Downsample(Chain):
    """Strided 3x3 conv; records the incoming spatial size in ``sampling.shapes``."""

    def __init__(
        self,
        channels: int,
        scale_factor: int,
        padding: int = 0,
        register_shape: bool = True,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.channels = channels
        self.in_channels = channels
        self.out_channels = channels
        self.scale_factor = scale_factor
        self.padding = padding
        super().__init__(
            Conv2d(channels, channels, kernel_size=3, stride=scale_factor, padding=padding, device=device, dtype=dtype)
        )
        if padding == 0:
            self.insert(0, Lambda(lambda x: F.pad(x, (0, 1, 0, 1))))
        if register_shape:
            self.insert(0, SetContext(context="sampling", key="shapes", callback=self.register_shape))

    def register_shape(self, shapes: list[Size], x: Tensor) -> None:
        shapes.append(x.shape[2:])


class Upsample(Chain):
    """Nearest resize (to a static factor or to the size a Downsample recorded) then 3x3 conv."""

    def __init__(
        self,
        channels: int,
        upsample_factor: int | None = None,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.channels = channels
        self.upsample_factor = upsample_factor
        size_source: Module = (
            Lambda(self._get_static_shape)
            if upsample_factor is not None
            else UseContext(context="sampling", key="shapes").compose(lambda shapes: shapes.pop())
        )
        super().__init__(
            Parallel(Identity(), size_source),
            Interpolate(),
            Conv2d(channels, channels, kernel_size=3, padding=1, device=device, dtype=dtype),
        )

    def _get_static_shape(self, x: Tensor) -> Size:
        assert self.upsample_factor is not None
        return Size([s * self.upsample_factor for s in x.shape[2:]])
