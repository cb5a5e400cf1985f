"""Stable Diffusion 1.5 UNet graph.

Tree shape follows
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_1/unet.py
(`TimestepEncoder` :16-28, `CLIPLCrossAttention` :31-46, `DownBlocks` :49-100,
`UpBlocks` :103-160, `MiddleBlock` :163-169 [sic], `SD1UNet` :165-249).
"""

from __future__ import annotations

from typing import Iterable, cast

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.context import Contexts
from refiners_b200.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock2d
from refiners_b200.foundationals.latent_diffusion.range_adapter import RangeAdapter2d, RangeEncoder
from refiners_b200.foundationals.latent_diffusion.unet_blocks import (
    ResidualAccumulator,
    ResidualBlock,
    ResidualConcatenator,
)

Device = torch.device
DType = torch.dtype


class TimestepEncoder(fl.Passthrough):
    def __init__(
        self,
        context_key: str = "timestep_embedding",
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        super().__init__(
            fl.UseContext("diffusion", "timestep"),
            RangeEncoder(320, 1280, device=device, dtype=dtype),
            fl.SetContext("range_adapter", context_key),
        )


class CLIPLCrossAttention(CrossAttentionBlock2d):
    def __init__(self, channels: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(
            channels=channels,
            context_embedding_dim=768,
            context_key="clip_text_embedding",
            num_attention_heads=8,
            use_bias=False,
            device=device,
            dtype=dtype,
        )


def _level(cin: int, cout: int, attention: bool, kw: dict, upsample: bool = False) -> fl.Chain:
    parts: list[fl.Module] = [ResidualBlock(in_channels=cin, out_channels=cout, **kw)]
    if attention:
        parts.append(CLIPLCrossAttention(channels=cout, **kw))
    if upsample:
        parts.append(fl.Upsample(channels=cout, **kw))
    return fl.Chain(*parts)


class DownBlocks(fl.Chain):
    def __init__(self, in_channels: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.in_channels = in_channels
        kw = dict(device=device, dtype=dtype)

        def down(c: int) -> fl.Chain:
            return fl.Chain(fl.Downsample(channels=c, scale_factor=2, padding=1, **kw))

        super().__init__(
            fl.Chain(fl.Conv2d(in_channels, 320, kernel_size=3, padding=1, **kw)),
            _level(320, 320, True, kw),
            _level(320, 320, True, kw),
            down(320),
            _level(320, 640, True, kw),
            _level(640, 640, True, kw),
            down(640),
            _level(640, 1280, True, kw),
            _level(1280, 1280, True, kw),
            down(1280),
            _level(1280, 1280, False, kw),
            _level(1280, 1280, False, kw),
        )


class UpBlocks(fl.Chain):
    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            _level(2560, 1280, False, kw),
            _level(2560, 1280, False, kw),
            _level(2560, 1280, False, kw, upsample=True),
            _level(2560, 1280, True, kw),
            _level(2560, 1280, True, kw),
            _level(1920, 1280, True, kw, upsample=True),
            _level(1920, 640, True, kw),
            _level(1280, 640, True, kw),
            _level(960, 640, True, kw, upsample=True),
            _level(960, 320, True, kw),
            _level(640, 320, True, kw),
            _level(640, 320, True, kw),
        )


class MiddleBlock(fl.Chain):
    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            ResidualBlock(in_channels=1280, out_channels=1280, **kw),
            CLIPLCrossAttention(channels=1280, **kw),
            ResidualBlock(in_channels=1280, out_channels=1280, **kw),
        )


class SD1UNet(fl.Chain):
    """Stable Diffusion 1.5 denoiser; conditioning through ``set_timestep`` /
    ``set_clip_text_embedding``."""

    def __init__(self, in_channels: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.in_channels = in_channels
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            TimestepEncoder(**kw),
            DownBlocks(in_channels=in_channels, **kw),
            fl.Sum(
                fl.UseContext(context="unet", key="residuals").compose(lambda x: x[-1]),
                MiddleBlock(**kw),
            ),
            UpBlocks(**kw),
            fl.Chain(
                fl.GroupNorm(channels=320, num_groups=32, **kw),
                fl.SiLU(),
                fl.Conv2d(320, 4, kernel_size=3, stride=1, padding=1, **kw),
            ),
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
            level.append(ResidualAccumulator(n))
        for n, level in enumerate(cast(Iterable[fl.Chain], self.UpBlocks)):
            level.insert(0, ResidualConcatenator(-n - 2))

    def init_context(self) -> Contexts:
        return {
            "unet": {"residuals": [0.0] * 13},
            "diffusion": {"timestep": None},
            "range_adapter": {"timestep_embedding": None},
            "sampling": {"shapes": []},
        }

    def set_clip_text_embedding(self, clip_text_embedding: Tensor) -> None:
        self.set_context("cross_attention_block", {"clip_text_embedding": clip_text_embedding})

    def set_timestep(self, timestep: Tensor) -> None:
        self.set_context("diffusion", {"timestep": timestep})
