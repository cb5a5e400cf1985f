"""DINOv2's Vision Transformer on the fluxion mirror.

Same module tree and state-dict keys as the reference's ``foundationals/dinov2/vit.py`` (`ViT` :289-413,
`TransformerLayer` :193-253, `Registers` :268-286, `InterpolateEmbedding` :55-100; compared node by node in
``tests/test_reference_structure.py``).  Every transformer layer is LayerNorm -> SelfAttention (head dim 64 for all
published sizes, i.e. the tcgen05 attention kernel) -> LayerScale, then LayerNorm -> MLP (Linear+GeLU epilogue, or
SwiGLU for the giant model) -> LayerScale, each with a residual - the leaves the diffusion path already runs on.
Batch-shards by image like SAM (no cross-image operation).
"""

from __future__ import annotations

from math import isqrt

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.utils import interpolate

Device = torch.device
DType = torch.dtype
_CTX = "dinov2_vit"


class ClassToken(fl.Chain):
    """The learnable [CLS] embedding, broadcast over the batch."""

    def __init__(self, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.embedding_dim = embedding_dim
        super().__init__(fl.Parameter(1, embedding_dim, device=device, dtype=dtype))


class PositionalEmbedding(fl.Chain):
    """Learnable positions for [CLS] + a square grid of patches (at the pre-training resolution)."""

    def __init__(
        self, sequence_length: int, embedding_dim: int, patch_size: int, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.sequence_length, self.embedding_dim, self.patch_size = sequence_length, embedding_dim, patch_size
        super().__init__(fl.Parameter(sequence_length, embedding_dim, device=device, dtype=dtype))


class InterpolateEmbedding(fl.Module):
    """Resample the patch-position grid to the grid of the actual input image; [CLS] keeps its embedding.
    The resize runs in fp32 whatever the model dtype (vit.py:86-91 in the reference)."""

    def __init__(self, mode: str, antialias: bool, patch_size: int) -> None:
        super().__init__()
        self.mode, self.antialias, self.patch_size = mode, antialias, patch_size

    def forward(self, x: Tensor, input: Tensor) -> Tensor:
        cls, grid = x[:, :1, :], x[:, 1:, :]
        batch, count, dim = grid.shape
        side = isqrt(count)
        assert side * side == count, "The sequence length must be a square number."
        target = torch.Size((input.shape[2] // self.patch_size, input.shape[3] // self.patch_size))
        grid = grid.reshape(batch, side, side, dim).permute(0, 3, 1, 2)
        grid = interpolate(x=grid.to(dtype=torch.float32), mode=self.mode, antialias=self.antialias, size=target)
        grid = grid.to(dtype=cls.dtype).permute(0, 2, 3, 1).reshape(batch, -1, dim)
        return torch.cat((cls, grid), dim=1)


class LayerScale(fl.WeightedModule):
    """Per-channel learnable gain on a residual branch."""

    def __init__(
        self, embedding_dim: int, init_value: float = 1.0, dtype: DType | None = None, device: Device | str | None = None,
    ) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        gain = torch.full(size=(embedding_dim,), fill_value=init_value, dtype=dtype, device=device)
        self.register_parameter(name="weight", param=torch.nn.Parameter(gain))

    def forward(self, x: Tensor) -> Tensor:
        return x * self.weight


class FeedForward(fl.Chain):
    """Linear -> activation -> Linear; a gated activation (GLU) consumes twice the hidden width."""

    def __init__(
        self, embedding_dim: int, feedforward_dim: int, activation: fl.Activation, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.feedforward_dim = embedding_dim, feedforward_dim
        kw = dict(device=device, dtype=dtype)
        expanded = 2 * feedforward_dim if isinstance(activation, fl.GLU) else feedforward_dim
        super().__init__(fl.Linear(embedding_dim, expanded, **kw), activation, fl.Linear(feedforward_dim, embedding_dim, **kw))


class PatchEncoder(fl.Chain):
    """Image -> patch tokens ``[B, (H/p)(W/p), D]`` (conv with kernel = stride = p); the image itself is kept in the
    context for the positional-embedding resize."""

    def __init__(
        self, in_channels: int, out_channels: int, patch_size: int, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.in_channels, self.out_channels, self.patch_size = in_channels, out_channels, patch_size
        super().__init__(
            fl.SetContext(context=_CTX, key="input"),
            fl.Conv2d(in_channels, out_channels, kernel_size=patch_size, stride=patch_size, device=device, dtype=dtype),
            fl.Reshape(out_channels, -1),
            fl.Transpose(1, 2),
        )


class TransformerLayer(fl.Chain):
    """Pre-norm attention and MLP branches, each scaled by a LayerScale and added back."""

    def __init__(
        self, embedding_dim: int, num_heads: int, norm_eps: float, mlp_ratio: int, activation: fl.Activation,
        feedforward_dim: int | None = None, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.num_heads, self.norm_eps, self.mlp_ratio = embedding_dim, num_heads, norm_eps, mlp_ratio
        self.feedforward_dim = feedforward_dim if feedforward_dim is not None else embedding_dim * mlp_ratio
        kw = dict(device=device, dtype=dtype)

        def branch(body: fl.Module) -> fl.Residual:
            return fl.Residual(fl.LayerNorm(embedding_dim, eps=norm_eps, **kw), body, LayerScale(embedding_dim, **kw))

        super().__init__(
            branch(fl.SelfAttention(embedding_dim=embedding_dim, num_heads=num_heads, **kw)),
            branch(FeedForward(embedding_dim, self.feedforward_dim, activation, **kw)),
        )


class Transformer(fl.Chain):
    """The stack of TransformerLayers."""


class PositionalEncoder(fl.Residual):
    """tokens + (resampled) positional embedding."""


class Registers(fl.Concatenate):
    """Register tokens spliced between [CLS] and the patch tokens (arXiv:2309.16588)."""

    def __init__(
        self, num_registers: int, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.num_registers, self.embedding_dim = num_registers, embedding_dim
        super().__init__(
            fl.Slicing(dim=1, end=1),
            fl.Parameter(num_registers, embedding_dim, device=device, dtype=dtype),
            fl.Slicing(dim=1, start=1),
            dim=1,
        )


class ViT(fl.Chain):
    """``[B, 3, H, W]`` (H, W multiples of the patch size) -> ``[B, 1 + registers + HW/p^2, D]`` tokens."""

    def __init__(
        self,
        embedding_dim: int = 768,
        patch_size: int = 16,
        image_size: int = 224,
        num_layers: int = 12,
        num_heads: int = 12,
        norm_eps: float = 1e-6,
        mlp_ratio: int = 4,
        num_registers: int = 0,
        activation: fl.Activation = fl.GeLU(),
        feedforward_dim: int | None = None,
        interpolate_antialias: bool = False,
        interpolate_mode: str = "bicubic",
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.patch_size, self.image_size = embedding_dim, patch_size, image_size
        self.num_layers, self.num_heads, self.norm_eps, self.mlp_ratio = num_layers, num_heads, norm_eps, mlp_ratio
        self.num_registers, self.feedforward_dim = num_registers, feedforward_dim
        kw = dict(device=device, dtype=dtype)
        grid = image_size // patch_size
        tokens = fl.Concatenate(ClassToken(embedding_dim, **kw), PatchEncoder(3, embedding_dim, patch_size, **kw), dim=1)
        positions = PositionalEncoder(
            PositionalEmbedding(grid * grid + 1, embedding_dim, patch_size, **kw),
            fl.Chain(
                fl.Parallel(fl.Identity(), fl.UseContext(context=_CTX, key="input")),
                InterpolateEmbedding(mode=interpolate_mode, antialias=interpolate_antialias, patch_size=patch_size),
            ),
        )
        layers = Transformer(
            TransformerLayer(
                embedding_dim=embedding_dim, feedforward_dim=feedforward_dim, activation=activation, num_heads=num_heads,
                mlp_ratio=mlp_ratio, norm_eps=norm_eps, **kw,
            )
            for _ in range(num_layers)
        )
        super().__init__(tokens, positions, layers, fl.LayerNorm(embedding_dim, eps=norm_eps, **kw))
        if num_registers > 0:
            self.insert_before_type(Transformer, Registers(num_registers, embedding_dim, **kw))

    def init_context(self) -> Contexts:
        return {_CTX: {"input": None}}
