"""CLIP vision towers: the image side of IP-Adapter.

``CLIPImageEncoderH`` turns a 224 x 224 image into the 1024-wide embedding that `ImageProjection` expands to four
prompt tokens (or, truncated before the pooling, into the 257 x 1280 patch features the `PerceiverResampler` consumes).
Module tree / state-dict keys follow the reference's ``foundationals/clip/image_encoder.py`` (`ViTEmbeddings` :95-137,
`TransformerLayer` :55-92, `CLIPImageEncoder` :140-204) and ``clip/common.py`` (`PositionalEncoder`, `FeedForward`).
Runs once per image prompt; every layer is a LayerNorm / Linear(+GeLU) / SelfAttention the engine already has kernels
for (head dim 80 for the H tower: the zero-padded tcgen05 attention path).
"""

from __future__ import annotations

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.clip.common import FeedForward, PositionalEncoder

Device = torch.device
DType = torch.dtype


class ClassToken(fl.Chain):
    def __init__(self, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.embedding_dim = embedding_dim
        super().__init__(fl.Parameter(1, embedding_dim, device=device, dtype=dtype))


class PatchEncoder(fl.Chain):
    """conv(kernel = stride = patch) then NCHW -> NHWC."""

    def __init__(
        self, in_channels: int, out_channels: int, patch_size: int = 16, use_bias: bool = True,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.in_channels, self.out_channels, self.patch_size, self.use_bias = in_channels, out_channels, patch_size, use_bias
        super().__init__(
            fl.Conv2d(in_channels, out_channels, kernel_size=(patch_size, patch_size), stride=(patch_size, patch_size),
                      use_bias=use_bias, device=device, dtype=dtype),
            fl.Permute(0, 2, 3, 1),
        )


class TransformerLayer(fl.Chain):
    """Pre-norm self-attention and GeLU MLP, each with a residual."""

    def __init__(
        self, embedding_dim: int = 768, feedforward_dim: int = 3072, num_attention_heads: int = 12, layer_norm_eps: float = 1e-5,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.feedforward_dim = embedding_dim, feedforward_dim
        self.num_attention_heads, self.layer_norm_eps = num_attention_heads, layer_norm_eps
        kw = dict(device=device, dtype=dtype)

        def norm() -> fl.LayerNorm:
            return fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **kw)

        super().__init__(
            fl.Residual(norm(), fl.SelfAttention(embedding_dim=embedding_dim, num_heads=num_attention_heads, **kw)),
            fl.Residual(norm(), FeedForward(embedding_dim, feedforward_dim, **kw)),
        )


class ViTEmbeddings(fl.Chain):
    """[CLS | patch tokens] + positions."""

    def __init__(
        self, image_size: int = 224, embedding_dim: int = 768, patch_size: int = 32, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.image_size, self.embedding_dim, self.patch_size = image_size, embedding_dim, patch_size
        kw = dict(device=device, dtype=dtype)
        grid = (image_size // patch_size) ** 2
        patches = fl.Chain(PatchEncoder(3, embedding_dim, patch_size=patch_size, use_bias=False, **kw), fl.Reshape(grid, embedding_dim))
        super().__init__(
            fl.Concatenate(ClassToken(embedding_dim, **kw), patches, dim=1),
            fl.Residual(PositionalEncoder(max_sequence_length=grid + 1, embedding_dim=embedding_dim, **kw)),
        )


class CLIPImageEncoder(fl.Chain):
    """``[B, 3, S, S]`` -> ``[B, output_dim]``: embeddings, pre-LayerNorm, transformer, [CLS] pooling, post-LayerNorm,
    bias-free projection."""

    def __init__(
        self,
        image_size: int = 224,
        embedding_dim: int = 768,
        output_dim: int = 512,
        patch_size: int = 32,
        num_layers: int = 12,
        num_attention_heads: int = 12,
        feedforward_dim: int = 3072,
        layer_norm_eps: float = 1e-5,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.image_size, self.embedding_dim, self.output_dim, self.patch_size = image_size, embedding_dim, output_dim, patch_size
        self.num_layers, self.num_attention_heads, self.feedforward_dim = num_layers, num_attention_heads, feedforward_dim
        kw = dict(device=device, dtype=dtype)
        cls_token_pooling = lambda x: x[:, 0, :]  # noqa: E731  (the name shows in repr())
        super().__init__(
            ViTEmbeddings(image_size=image_size, embedding_dim=embedding_dim, patch_size=patch_size, **kw),
            fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **kw),
            fl.Chain(
                TransformerLayer(embedding_dim=embedding_dim, feedforward_dim=feedforward_dim,
                                 num_attention_heads=num_attention_heads, layer_norm_eps=layer_norm_eps, **kw)
                for _ in range(num_layers)
            ),
            fl.Lambda(func=cls_token_pooling),
            fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **kw),
            fl.Linear(embedding_dim, output_dim, bias=False, **kw),
        )


class CLIPImageEncoderH(CLIPImageEncoder):
    """ViT-H/14: 1280 wide, 32 layers, 16 heads (d = 80), 1024-d output."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(embedding_dim=1280, output_dim=1024, patch_size=14, num_layers=32, num_attention_heads=16,
                         feedforward_dim=5120, device=device, dtype=dtype)


class CLIPImageEncoderG(CLIPImageEncoder):
    """ViT-bigG/14: 1664 wide, 48 layers, 16 heads (d = 104), 1280-d output."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(embedding_dim=1664, output_dim=1280, patch_size=14, num_layers=48, num_attention_heads=16,
                         feedforward_dim=8192, device=device, dtype=dtype)
