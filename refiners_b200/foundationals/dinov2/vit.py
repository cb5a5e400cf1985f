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


def _note(module: fl.Module, **hyper: object) -> dict[str, object]:
    """Record constructor arguments on the module (they are what ``repr`` echoes and what adapters read)."""
    for name, value in hyper.items():
        setattr(module, name, value)
    return hyper


def resample_positions(positions: Tensor, grid_hw: tuple[int, int], mode: str, antialias: bool) -> Tensor:
    """``[B, 1 + s*s, D]`` learned positions -> ``[B, 1 + h*w, D]``: the square grid of patch positions is resized as an
    image (in fp32 whatever the model dtype: the reference's vit.py:86-91 does the same), [CLS] keeps its row."""
    batch, rows, dim = positions.shape
    side = isqrt(rows - 1)
    assert side * side == rows - 1, "The sequence length must be a square number."
    as_image = positions[:, 1:, :].unflatten(1, (side, side)).movedim(-1, 1)
    resized = interpolate(x=as_image.to(dtype=torch.float32), mode=mode, antialias=antialias, size=torch.Size(grid_hw))
    patches = resized.to(dtype=positions.dtype).movedim(1, -1).reshape(batch, -1, dim)
    return torch.cat((positions[:, :1, :], patches), dim=1)


class ClassToken(fl.Chain):
    """The learnable [CLS] embedding, broadcast over the batch."""

    def __init__(self, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        _note(self, embedding_dim=embedding_dim)
        super().__init__(fl.Parameter(1, embedding_dim, device=device, dtype=dtype))


class PositionalEmbedding(fl.Chain):
    """Learnable positions for [CLS] + a square grid of patches (at the pre-training resolution)."""

    def __init__(
        self, sequence_length: int, embedding_dim: int, patch_size: int, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        _note(self, sequence_length=sequence_length, embedding_dim=embedding_dim, patch_size=patch_size)
        super().__init__(fl.Parameter(sequence_length, embedding_dim, device=device, dtype=dtype))


class InterpolateEmbedding(fl.Module):
    """Resample the positions to the patch grid of the actual input image (`resample_positions`)."""

    def __init__(self, mode: str, antialias: bool, patch_size: int) -> None:
        super().__init__()
        _note(self, mode=mode, antialias=antialias, patch_size=patch_size)

    def forward(self, x: Tensor, input: Tensor) -> Tensor:
        height, width = input.shape[-2:]
        return resample_positions(x, (height // self.patch_size, width // self.patch_size), self.mode, self.antialias)


class LayerScale(fl.WeightedModule):
    """Per-channel learnable gain on a residual branch."""

    def __init__(
        self, embedding_dim: int, init_value: float = 1.0, dtype: DType | None = None, device: Device | str | None = None,
    ) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.weight = torch.nn.Parameter(torch.full((embedding_dim,), init_value, dtype=dtype, device=device))

    def forward(self, x: Tensor) -> Tensor:
        return self.weight * x


class FeedForward(fl.Chain):
    """Linear -> activation -> Linear; a gated activation (GLU) consumes twice the hidden width."""

    def __init__(
        self, embedding_dim: int, feedforward_dim: int, activation: fl.Activation, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        _note(self, embedding_dim=embedding_dim, feedforward_dim=feedforward_dim)
        gated = isinstance(activation, fl.GLU)
        super().__init__(
            fl.Linear(embedding_dim, feedforward_dim * (2 if gated else 1), device=device, dtype=dtype),
            activation,
            fl.Linear(feedforward_dim, embedding_dim, device=device, dtype=dtype),
        )


class PatchEncoder(fl.Chain):
    """Image -> patch tokens ``[B, (H/p)(W/p), D]`` (conv with kernel = stride = p); the image itself is kept in the
    context for the positional-embedding resize."""

    def __init__(
        self, in_channels: int, out_channels: int, patch_size: int, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        _note(self, in_channels=in_channels, out_channels=out_channels, patch_size=patch_size)
        to_patches = fl.Conv2d(in_channels, out_channels, kernel_size=patch_size, stride=patch_size, device=device, dtype=dtype)
        super().__init__(fl.SetContext(context=_CTX, key="input"), to_patches, fl.Reshape(out_channels, -1), fl.Transpose(1, 2))


class TransformerLayer(fl.Chain):
    """Pre-norm attention and MLP branches, each scaled by a LayerScale and added back."""

    def __init__(
        self, embedding_dim: int, num_heads: int, norm_eps: float, mlp_ratio: int, activation: fl.Activation,
        feedforward_dim: int | None = None, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        hidden = embedding_dim * mlp_ratio if feedforward_dim is None else feedforward_dim
        _note(self, embedding_dim=embedding_dim, num_heads=num_heads, norm_eps=norm_eps, mlp_ratio=mlp_ratio, feedforward_dim=hidden)
        on = dict(device=device, dtype=dtype)
        mixers = (
            fl.SelfAttention(embedding_dim=embedding_dim, num_heads=num_heads, **on),
            FeedForward(embedding_dim, hidden, activation, **on),
        )
        super().__init__(
            fl.Residual(fl.LayerNorm(embedding_dim, eps=norm_eps, **on), mixer, LayerScale(embedding_dim, **on)) for mixer in mixers
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
        _note(self, num_registers=num_registers, embedding_dim=embedding_dim)
        learned = fl.Parameter(num_registers, embedding_dim, device=device, dtype=dtype)
        super().__init__(fl.Slicing(dim=1, end=1), learned, fl.Slicing(dim=1, start=1), dim=1)


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
        _note(
            self, embedding_dim=embedding_dim, patch_size=patch_size, image_size=image_size, num_layers=num_layers, num_heads=num_heads,
            norm_eps=norm_eps, mlp_ratio=mlp_ratio, num_registers=num_registers, feedforward_dim=feedforward_dim,
        )
        on = dict(device=device, dtype=dtype)
        learned_positions = PositionalEmbedding((image_size // patch_size) ** 2 + 1, embedding_dim, patch_size, **on)
        stages: list[fl.Module] = [
            fl.Concatenate(ClassToken(embedding_dim, **on), PatchEncoder(3, embedding_dim, patch_size, **on), dim=1),
            PositionalEncoder(
                learned_positions,
                fl.Chain(
                    fl.Parallel(fl.Identity(), fl.UseContext(context=_CTX, key="input")),
                    InterpolateEmbedding(mode=interpolate_mode, antialias=interpolate_antialias, patch_size=patch_size),
                ),
            ),
        ]
        if num_registers > 0:
            stages.append(Registers(num_registers, embedding_dim, **on))
        stages.append(
            Transformer(
                TransformerLayer(embedding_dim, num_heads, norm_eps, mlp_ratio, activation, feedforward_dim=feedforward_dim, **on)
                for _ in range(num_layers)
            )
        )
        stages.append(fl.LayerNorm(embedding_dim, eps=norm_eps, **on))
        super().__init__(*stages)

    def init_context(self) -> Contexts:
        return {_CTX: {"input": None}}
