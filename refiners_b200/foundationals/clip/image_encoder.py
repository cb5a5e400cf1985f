"""CLIP vision towers: the image side of IP-Adapter.

``CLIPImageEncoderH`` turns a 224 x 224 image into the 1024-wide embedding that `ImageProjection` expands to four
prompt tokens (or, truncated before the pooling, into the 257 x 1280 patch features the `PerceiverResampler` consumes).
Module tree / state-dict keys follow the reference's ``foundationals/clip/image_encoder.py`` (`ViTEmbeddings` :95-137,
`TransformerLayer` :55-92, `CLIPImageEncoder` :140-204) and ``clip/common.py`` (`PositionalEncoder`, `FeedForward`).
Runs once per image prompt; every layer is a LayerNorm / Linear(+GeLU) / SelfAttention the engine already has kernels
for (head dim 80 for the H tower: the zero-padded tcgen05 attention path).
"""

from __future__ import annotations

from typing import Any

import torch

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.clip.common import FeedForward, PositionalEncoder

Device = torch.device
DType = torch.dtype


# published vision towers: (width, layers, heads, feed-forward width, output width); all see 224 x 224 images in 14 x 14 patches
_TOWERS: dict[str, tuple[int, int, int, int, int]] = {
    "H": (1280, 32, 16, 5120, 1024),  # OpenCLIP ViT-H/14  (head dim 80): IP-Adapter for SD 1.5 and the SDXL "vit-h" variants
    "G": (1664, 48, 16, 8192, 1280),  # OpenCLIP ViT-bigG/14 (head dim 104): the first SDXL IP-Adapter
}


def _note(module: fl.Module, **hyper: Any) -> None:
    """Constructor arguments kept on the module: ``repr`` echoes them, adapters read them."""
    for name, value in hyper.items():
        setattr(module, name, value)


class ClassToken(fl.Chain):
    def __init__(self, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        _note(self, embedding_dim=embedding_dim)
        super().__init__(fl.Parameter(1, embedding_dim, device=device, dtype=dtype))


class PatchEncoder(fl.Chain):
    """conv(kernel = stride = patch) then NCHW -> NHWC."""

    def __init__(
        self, in_channels: int, out_channels: int, patch_size: int = 16, use_bias: bool = True,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        _note(self, in_channels=in_channels, out_channels=out_channels, patch_size=patch_size, use_bias=use_bias)
        square = (patch_size, patch_size)
        cut = fl.Conv2d(in_channels, out_channels, kernel_size=square, stride=square, use_bias=use_bias, device=device, dtype=dtype)
        super().__init__(cut, fl.Permute(0, 2, 3, 1))


class TransformerLayer(fl.Chain):
    """Pre-norm self-attention and GeLU MLP, each with a residual."""

    def __init__(
        self, embedding_dim: int = 768, feedforward_dim: int = 3072, num_attention_heads: int = 12, layer_norm_eps: float = 1e-5,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        _note(self, embedding_dim=embedding_dim, feedforward_dim=feedforward_dim, num_attention_heads=num_attention_heads,
              layer_norm_eps=layer_norm_eps)
        on: dict[str, Any] = {"device": device, "dtype": dtype}
        mixers = (
            fl.SelfAttention(embedding_dim=embedding_dim, num_heads=num_attention_heads, **on),
            FeedForward(embedding_dim, feedforward_dim, **on),
        )
        super().__init__(fl.Residual(fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **on), mixer) for mixer in mixers)


class ViTEmbeddings(fl.Chain):
    """[CLS | patch tokens] + positions."""

    def __init__(
        self, image_size: int = 224, embedding_dim: int = 768, patch_size: int = 32, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        _note(self, image_size=image_size, embedding_dim=embedding_dim, patch_size=patch_size)
        on: dict[str, Any] = {"device": device, "dtype": dtype}
        patches_per_side = image_size // patch_size
        count = patches_per_side * patches_per_side
        patch_tokens = fl.Chain(
            PatchEncoder(3, embedding_dim, patch_size=patch_size, use_bias=False, **on), fl.Reshape(count, embedding_dim),
        )
        super().__init__(
            fl.Concatenate(ClassToken(embedding_dim, **on), patch_tokens, dim=1),
            fl.Residual(PositionalEncoder(max_sequence_length=1 + count, embedding_dim=embedding_dim, **on)),
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
        _note(self, image_size=image_size, embedding_dim=embedding_dim, output_dim=output_dim, patch_size=patch_size,
              num_layers=num_layers, num_attention_heads=num_attention_heads, feedforward_dim=feedforward_dim)
        on: dict[str, Any] = {"device": device, "dtype": dtype}

        def norm() -> fl.LayerNorm:
            return fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **on)

        trunk = fl.Chain(
            TransformerLayer(embedding_dim, feedforward_dim, num_attention_heads, layer_norm_eps, **on) for _ in range(num_layers)
        )
        super().__init__(
            ViTEmbeddings(image_size=image_size, embedding_dim=embedding_dim, patch_size=patch_size, **on),
            norm(),
            trunk,
            fl.Lambda(func=lambda x: x.select(1, 0)),  # the [CLS] row of every image; repr() prints "<lambda>(x)"
            norm(),
            fl.Linear(embedding_dim, output_dim, bias=False, **on),
        )


def _published(tag: str, device: Device | str | None, dtype: DType | None) -> dict[str, Any]:
    width, layers, heads, hidden, out = _TOWERS[tag]
    return dict(embedding_dim=width, output_dim=out, patch_size=14, num_layers=layers, num_attention_heads=heads,
                feedforward_dim=hidden, device=device, dtype=dtype)


class CLIPImageEncoderH(CLIPImageEncoder):
    """ViT-H/14: 1280 wide, 32 layers, 16 heads (d = 80), 1024-d output."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(**_published("H", device, dtype))


class CLIPImageEncoderG(CLIPImageEncoder):
    """ViT-bigG/14: 1664 wide, 48 layers, 16 heads (d = 104), 1280-d output."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(**_published("G", device, dtype))
