"""Compute leaves: Linear, Conv2d, norms, activations, scaled-dot-product attention.

Constructor signatures, class names and parameter names follow the reference
(linear.py:9-58, conv.py:6-61, norm.py:14-127, activations.py:31-160, attentions.py:60-202
under /root/reference/src/refiners/fluxion/layers/) because they are the drop-in boundary:
state-dict keys, ``isinstance(m, torch.nn.Linear)`` checks and ``repr`` depend on them.

Execution: a CUDA input always goes through the hand-written sm_100a kernels behind the
C ABI (refiners_b200.backend) - there is no ATen fallback on the GPU and a missing
``librefiners_b200.so`` raises.  A CPU input runs the torch.nn parent forward; that host path
is what BASELINE config 1 (SD1UNet fp32 on CPU, plumbing only) exercises.
"""

from __future__ import annotations

import math
from enum import Enum
import torch
from torch import Tensor, nn
from torch.nn import functional as F

from refiners_b200 import backend as B
from refiners_b200.fluxion.layers.base import Module, WeightedModule

Device = torch.device
DType = torch.dtype


# ------------------------------------------------------------------------------------ linear
class Linear(nn.Linear, WeightedModule):
    def __init__(
        self,
        in_features: int,
        out_features: int,
        bias: bool = True,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.in_features = in_features
        self.out_features = out_features
        nn.Linear.__init__(self, in_features, out_features, bias=bias, device=device, dtype=dtype)

    def forward(self, x: Tensor) -> Tensor:  # type: ignore[override]
        if x.is_cuda:
            return B.linear(x, self.weight, self.bias)
        return F.linear(x, self.weight, self.bias)


# -------------------------------------------------------------------------------------- conv
class Conv2d(nn.Conv2d, WeightedModule):
    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        kernel_size: int | tuple[int, int],
        stride: int | tuple[int, int] = (1, 1),
        padding: int | tuple[int, int] | str = (0, 0),
        groups: int = 1,
        use_bias: bool = True,
        dilation: int | tuple[int, int] = (1, 1),
        padding_mode: str = "zeros",
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        nn.Conv2d.__init__(
            self,
            in_channels,
            out_channels,
            kernel_size,  # type: ignore[arg-type]
            stride=stride,  # type: ignore[arg-type]
            padding=padding,  # type: ignore[arg-type]
            dilation=dilation,  # type: ignore[arg-type]
            groups=groups,
            bias=use_bias,
            padding_mode=padding_mode,
            device=device,
            dtype=dtype,
        )
        self.use_bias = use_bias

    def forward(self, x: Tensor) -> Tensor:  # type: ignore[override]
        if x.is_cuda:
            return B.conv2d_module(x, self)
        return nn.Conv2d.forward(self, x)


class ConvTranspose2d(nn.ConvTranspose2d, WeightedModule):
    """Not on the denoising path (used by out-of-scope decoders); host/ATen only."""

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        kernel_size: int | tuple[int, int],
        stride: int | tuple[int, int] = 1,
        padding: int | tuple[int, int] = 0,
        output_padding: int | tuple[int, int] = 0,
        groups: int = 1,
        use_bias: bool = True,
        dilation: int | tuple[int, int] = 1,
        padding_mode: str = "zeros",
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        nn.ConvTranspose2d.__init__(
            self,
            in_channels,
            out_channels,
            kernel_size,  # type: ignore[arg-type]
            stride=stride,  # type: ignore[arg-type]
            padding=padding,  # type: ignore[arg-type]
            output_padding=output_padding,  # type: ignore[arg-type]
            groups=groups,
            bias=use_bias,
            dilation=dilation,  # type: ignore[arg-type]
            padding_mode=padding_mode,
            device=device,
            dtype=dtype,
        )
        self.use_bias = use_bias


# ------------------------------------------------------------------------------------- norms
class LayerNorm(nn.LayerNorm, WeightedModule):
    def __init__(
        self,
        normalized_shape: int | list[int],
        eps: float = 0.00001,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        nn.LayerNorm.__init__(self, normalized_shape, eps=eps, elementwise_affine=True, device=device, dtype=dtype)

    def forward(self, x: Tensor) -> Tensor:  # type: ignore[override]
        if x.is_cuda and len(self.normalized_shape) == 1:
            return B.layer_norm(x, self.weight, self.bias, self.eps)
        return nn.LayerNorm.forward(self, x)


class GroupNorm(nn.GroupNorm, WeightedModule):
    def __init__(
        self,
        channels: int,
        num_groups: int,
        eps: float = 1e-5,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        nn.GroupNorm.__init__(self, num_groups, channels, eps=eps, affine=True, device=device, dtype=dtype)
        self.channels = channels
        self.num_groups = num_groups
        self.eps = eps

    def forward(self, x: Tensor) -> Tensor:  # type: ignore[override]
        if x.is_cuda:
            return B.group_norm(x, self.num_groups, self.weight, self.bias, self.eps, silu=False)
        return nn.GroupNorm.forward(self, x)


class LayerNorm2d(WeightedModule):
    """Per-pixel normalisation over channels of an NCHW map (reference norm.py:95-127)."""

    def __init__(
        self,
        channels: int,
        eps: float = 1e-6,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        super().__init__()
        self.weight = nn.Parameter(torch.ones(channels, device=device, dtype=dtype))
        self.bias = nn.Parameter(torch.zeros(channels, device=device, dtype=dtype))
        self.eps = eps

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            return B.layer_norm_2d(x, self.weight, self.bias, self.eps)
        mean = x.mean(1, keepdim=True)
        centred = x - mean
        var = centred.pow(2).mean(1, keepdim=True)
        return self.weight[:, None, None] * (centred / torch.sqrt(var + self.eps)) + self.bias[:, None, None]


class InstanceNorm2d(nn.InstanceNorm2d, Module):
    def __init__(
        self,
        num_features: int,
        eps: float = 1e-05,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        nn.InstanceNorm2d.__init__(self, num_features, eps=eps, device=device, dtype=dtype)


# ------------------------------------------------------------------------------- activations
class Activation(Module):
    def __init__(self) -> None:
        super().__init__()


class SiLU(Activation):
    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            return B.unary(x, "silu")
        return F.silu(x)


class ReLU(Activation):
    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            return B.unary(x, "relu")
        return F.relu(x)


class Sigmoid(Activation):
    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            return B.unary(x, "sigmoid")
        return torch.sigmoid(x)


class GeLUApproximation(Enum):
    NONE = "none"
    TANH = "tanh"
    SIGMOID = "sigmoid"


class GeLU(Activation):
    def __init__(self, approximation: GeLUApproximation = GeLUApproximation.NONE) -> None:
        super().__init__()
        self.approximation = approximation

    def forward(self, x: Tensor) -> Tensor:
        kind = self.approximation
        if x.is_cuda:
            return B.unary(x, {"none": "gelu", "tanh": "gelu_tanh", "sigmoid": "gelu_sigmoid"}[kind.value])
        if kind is GeLUApproximation.NONE:
            return F.gelu(x, approximate="none")
        if kind is GeLUApproximation.TANH:
            return F.gelu(x, approximate="tanh")
        return x * torch.sigmoid(1.702 * x)


class GLU(Activation):
    """``a * act(g)`` with ``(a, g) = x.chunk(2, -1)`` (reference activations.py:136-160)."""

    def __init__(self, activation: Activation) -> None:
        super().__init__()
        self.activation = activation

    def __repr__(self) -> str:
        return f"{type(self).__name__}(activation={self.activation})"

    def forward(self, x: Tensor) -> Tensor:
        assert x.shape[-1] % 2 == 0, "Non-batch input dimension must be divisible by 2"
        if x.is_cuda and type(self.activation) is GeLU and self.activation.approximation is GeLUApproximation.NONE:
            return B.geglu(x)
        value, gate = x.chunk(2, dim=-1)
        return value * self.activation(gate)


# --------------------------------------------------------------------------------- attention
def scaled_dot_product_attention(query: Tensor, key: Tensor, value: Tensor, is_causal: bool = False) -> Tensor:
    """Host (CPU) path, [B, H, S, d] operands - same call as the reference (attentions.py:15-34)."""
    return F.scaled_dot_product_attention(query, key, value, is_causal=is_causal)


def scaled_dot_product_attention_non_optimized(
    query: Tensor, key: Tensor, value: Tensor, is_causal: bool = False
) -> Tensor:
    if is_causal:
        raise NotImplementedError(
            "Causal attention for `scaled_dot_product_attention_non_optimized` is not yet implemented"
        )
    logits = (query @ key.transpose(-1, -2)) / math.sqrt(query.shape[-1])
    return torch.softmax(logits, dim=-1) @ value


class ScaledDotProductAttention(Module):
    """Multi-head attention over [B, S, C] operands; heads are split and merged inside.

    On CUDA the head split/merge is folded into the kernel's addressing (the kernel reads
    [B, S, H, d] views in place), so no transposed copies are made.
    """

    def __init__(
        self,
        num_heads: int = 1,
        is_causal: bool = False,
        is_optimized: bool = True,
        slice_size: int | None = None,
    ) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.is_causal = is_causal
        self.is_optimized = is_optimized
        self.slice_size = slice_size
        self.dot_product = (
            scaled_dot_product_attention if is_optimized else scaled_dot_product_attention_non_optimized
        )

    def forward(self, query: Tensor, key: Tensor, value: Tensor) -> Tensor:
        if self.slice_size:
            return self._sliced_attention(query, key, value, slice_size=self.slice_size)
        return self._process_attention(query, key, value)

    def _sliced_attention(self, query: Tensor, key: Tensor, value: Tensor, slice_size: int) -> Tensor:
        if query.is_cuda:
            # the flash kernel never materialises the Sq x Sk matrix, so query slicing (a memory
            # work-around in the reference, attentions.py:135-155) is a no-op here
            return self._process_attention(query, key, value)
        out = torch.zeros_like(query)
        for lo in range(0, query.shape[1], slice_size):
            hi = min(lo + slice_size, query.shape[1])
            out[:, lo:hi, :] = self._process_attention(query[:, lo:hi, :], key, value)
        return out

    def _process_attention(self, query: Tensor, key: Tensor, value: Tensor) -> Tensor:
        if query.is_cuda:
            self._check(query)
            return B.sdpa(query, key, value, self.num_heads, self.is_causal)
        return self._merge_multi_head(
            self.dot_product(
                query=self._split_to_multi_head(query),
                key=self._split_to_multi_head(key),
                value=self._split_to_multi_head(value),
                is_causal=self.is_causal,
            )
        )

    def _check(self, x: Tensor) -> None:
        assert x.ndim == 3, f"Expected input tensor with shape (batch_size sequence_length embedding_dim), got {x.shape}"
        assert x.shape[-1] % self.num_heads == 0, (
            f"Expected embedding_dim (x.shape[-1]={x.shape[-1]}) to be divisible by num_heads ({self.num_heads})"
        )

    def _split_to_multi_head(self, x: Tensor) -> Tensor:
        self._check(x)
        b, s, c = x.shape
        return x.reshape(b, s, self.num_heads, c // self.num_heads).transpose(1, 2)

    def _merge_multi_head(self, x: Tensor) -> Tensor:
        b, h, s, d = x.shape
        return x.transpose(1, 2).reshape(b, s, h * d)


# --------------------------------------------------------------------------------- embedding
class Embedding(nn.Embedding, WeightedModule):
    def __init__(
        self,
        num_embeddings: int,
        embedding_dim: int,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        nn.Embedding.__init__(self, num_embeddings, embedding_dim, device=device, dtype=dtype)
