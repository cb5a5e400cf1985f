"""SDXL UNet graph.

Tree shape (hence state-dict keys) follows
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/unet.py:
`TextTimeEmbedding` :20-59, `TimestepEncoder` :62-90, `SDXLCrossAttention` :93-112,
`DownBlocks` :115-170, `UpBlocks` :173-235, `MiddleBlock` :238-246, `OutputBlock` :249-255,
`SDXLUNet` :258-351.  Sub-modules are constructed in the reference's order so that a seeded
random init reproduces the reference's weights.
"""

from __future__ import annotations

from typing import Iterable, cast

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.context import Contexts
from refiners_b200.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock2d
from refiners_b200.foundationals.latent_diffusion.range_adapter import (
    RangeAdapter2d,
    RangeEncoder,
    compute_sinusoidal_embedding,
)
from refiners_b200.foundationals.latent_diffusion.unet_blocks import (
    ResidualAccumulator,
    ResidualBlock,
    ResidualConcatenator,
)

Device = torch.device
DType = torch.dtype


class TextTimeEmbedding(fl.Chain):
    """cat(pooled text embedding, sinusoid(time_ids)) -> 2-layer MLP -> [B, 1280]."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.timestep_embedding_dim, self.time_ids_embedding_dim, self.text_time_embedding_dim = 1280, 256, 2816
        super().__init__(
            fl.Concatenate(
                fl.UseContext(context="diffusion", key="pooled_text_embedding"),
                fl.Chain(
                    fl.UseContext(context="diffusion", key="time_ids"),
                    fl.Unsqueeze(dim=-1),
                    fl.Lambda(func=self.compute_sinusoidal_embedding),
                    fl.Reshape(-1),
                ),
                dim=1,
            ),
            fl.Converter(set_device=False, set_dtype=True),
            fl.Linear(self.text_time_embedding_dim, self.timestep_embedding_dim, device=device, dtype=dtype),
            fl.SiLU(),
            fl.Linear(self.timestep_embedding_dim, self.timestep_embedding_dim, device=device, dtype=dtype),
        )

    def compute_sinusoidal_embedding(self, x: Tensor) -> Tensor:
        return compute_sinusoidal_embedding(x=x, embedding_dim=self.time_ids_embedding_dim)


class TimestepEncoder(fl.Passthrough):
    """Writes ``range_adapter.<context_key>`` = MLP(sinusoid(timestep)) + TextTimeEmbedding."""

    def __init__(
        self, context_key: str = "timestep_embedding", device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.timestep_embedding_dim = 1280
        super().__init__(
            fl.Sum(
                fl.Chain(
                    fl.UseContext(context="diffusion", key="timestep"),
                    RangeEncoder(320, self.timestep_embedding_dim, device=device, dtype=dtype),
                ),
                TextTimeEmbedding(device=device, dtype=dtype),
            ),
            fl.SetContext(context="range_adapter", key=context_key),
        )

    @property
    def context_key(self) -> str:
        sink = self.ensure_find(fl.SetContext)
        assert sink.context == "range_adapter"
        return sink.key

    @context_key.setter
    def context_key(self, value: str) -> None:
        sink = self.ensure_find(fl.SetContext)
        assert sink.context == "range_adapter"
        sink.key = value


class SDXLCrossAttention(CrossAttentionBlock2d):
    def __init__(
        self, channels: int, num_attention_layers: int = 1, num_attention_heads: int = 10,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        super().__init__(
            channels=channels,
            context_embedding_dim=2048,
            context_key="clip_text_embedding",
            num_attention_layers=num_attention_layers,
            num_attention_heads=num_attention_heads,
            use_bias=False,
            use_linear_projection=True,
            device=device,
            dtype=dtype,
        )


def _stage(cin: int, cout: int, layers: int, heads: int, kw: dict, *tail: fl.Module) -> fl.Chain:
    """Chain(ResidualBlock[, SDXLCrossAttention][, tail...]) - one UNet level entry."""
    parts: list[fl.Module] = [ResidualBlock(in_channels=cin, out_channels=cout, **kw)]
    if layers:
        parts.append(SDXLCrossAttention(channels=cout, num_attention_layers=layers, num_attention_heads=heads, **kw))
    return fl.Chain(*parts, *tail)


class DownBlocks(fl.Chain):
    def __init__(self, in_channels: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.in_channels = in_channels
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            fl.Chain(fl.Conv2d(in_channels, 320, kernel_size=3, padding=1, **kw)),
            _stage(320, 320, 0, 0, kw),
            _stage(320, 320, 0, 0, kw),
            fl.Chain(fl.Downsample(channels=320, scale_factor=2, padding=1, **kw)),
            _stage(320, 640, 2, 10, kw),
            _stage(640, 640, 2, 10, kw),
            fl.Chain(fl.Downsample(channels=640, scale_factor=2, padding=1, **kw)),
            _stage(640, 1280, 10, 20, kw),
            _stage(1280, 1280, 10, 20, kw),
        )


class UpBlocks(fl.Chain):
    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        kw = dict(device=device, dtype=dtype)

        def stage_up(cin: int, cout: int, layers: int, heads: int) -> fl.Chain:
            # the Upsample conv is created after the attention stack, as in the reference
            res = ResidualBlock(in_channels=cin, out_channels=cout, **kw)
            attn = SDXLCrossAttention(channels=cout, num_attention_layers=layers, num_attention_heads=heads, **kw)
            return fl.Chain(res, attn, fl.Upsample(channels=cout, **kw))

        super().__init__(
            _stage(2560, 1280, 10, 20, kw),
            _stage(2560, 1280, 10, 20, kw),
            stage_up(1920, 1280, 10, 20),
            _stage(1920, 640, 2, 10, kw),
            _stage(1280, 640, 2, 10, kw),
            stage_up(960, 640, 2, 10),
            _stage(960, 320, 0, 0, kw),
            _stage(640, 320, 0, 0, kw),
            _stage(640, 320, 0, 0, kw),
        )


class MiddleBlock(fl.Chain):
    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            ResidualBlock(in_channels=1280, out_channels=1280, **kw),
            SDXLCrossAttention(channels=1280, num_attention_layers=10, num_attention_heads=20, **kw),
            ResidualBlock(in_channels=1280, out_channels=1280, **kw),
        )


class OutputBlock(fl.Chain):
    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            fl.GroupNorm(channels=320, num_groups=32, **kw),
            fl.SiLU(),
            fl.Conv2d(320, 4, kernel_size=3, stride=1, padding=1, **kw),
        )


class SDXLUNet(fl.Chain):
    """Stable Diffusion XL denoiser: latents [B, in_channels, H, W] -> predicted noise [B, 4, H, W].

    Conditioning arrives through contexts (``set_timestep``, ``set_clip_text_embedding``,
    ``set_pooled_text_embedding``, ``set_time_ids``).
    """

    def __init__(self, in_channels: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.in_channels = in_channels
        super().__init__(
            TimestepEncoder(device=device, dtype=dtype),
            DownBlocks(in_channels=in_channels, device=device, dtype=dtype),
            MiddleBlock(device=device, dtype=dtype),
            fl.Residual(fl.UseContext(context="unet", key="residuals").compose(lambda x: x[-1])),
            UpBlocks(device=device, dtype=dtype),
            OutputBlock(device=device, dtype=dtype),
        )
        for block in self.layers(ResidualBlock):
            body = block.layer("Chain", fl.Chain)
            RangeAdapter2d(
                target=body.layer("Conv2d_1", fl.Conv2d),
                channels=block.out_channels,
                embedding_dim=1280,
                context_key="timestep_embedding",
                device=device,
                dtype=dtype,
            ).inject(body)
        for n, level in enumerate(cast(Iterable[fl.Chain], self.DownBlocks)):
            level.append(ResidualAccumulator(n=n))
        for n, level in enumerate(cast(Iterable[fl.Chain], self.UpBlocks)):
            level.insert(0, ResidualConcatenator(n=-n - 2))

    def init_context(self) -> Contexts:
        return {
            "unet": {"residuals": [0.0] * 10},
            "diffusion": {"timestep": None, "time_ids": None, "pooled_text_embedding": None},
            "range_adapter": {"timestep_embedding": None},
            "sampling": {"shapes": []},
        }

    def set_clip_text_embedding(self, clip_text_embedding: Tensor) -> None:
        self.set_context("cross_attention_block", {"clip_text_embedding": clip_text_embedding})

    def set_timestep(self, timestep: Tensor) -> None:
        self.set_context("diffusion", {"timestep": timestep})

    def set_time_ids(self, time_ids: Tensor) -> None:
        self.set_context("diffusion", {"time_ids": time_ids})

    def set_pooled_text_embedding(self, pooled_text_embedding: Tensor) -> None:
        self.set_context("diffusion", {"pooled_text_embedding": pooled_text_embedding})
