"""T2I-Adapter (arXiv:2302.08453): a small convolutional encoder turns a condition image (depth, canny, pose ...) into four
feature maps, computed ONCE per image, which are added to the UNet's encoder activations at four resolutions on every step.

Contract (class names, constructor arguments, trees / state-dict keys, context names) from
/root/reference/src/refiners/foundationals/latent_diffusion/t2i_adapter.py: `Downsample2d` :17-19, `ResidualBlock` :22-36,
`ResidualBlocks` :39-63, `StatefulResidualBlocks` :66-91, `ConditionEncoder` :94-127, `ConditionEncoderXL` :130-161,
`T2IFeatures` :164-174, `T2IAdapter` :177-220.  Where the features enter each UNet family is decided by
stable_diffusion_1/t2i_adapter.py and stable_diffusion_xl/t2i_adapter.py.

On CUDA the encoder is convs (3x3 + ReLU + 1x1 with the residual in the second conv's epilogue), `rb200_avg_pool2d` for the
2x downsampling, and per step each `T2IFeatures` is ONE launch: ``x + scale * feature`` (`rb200_add` with alpha = scale).
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any, Generic, TypeVar

import torch
from torch import Tensor
from torch.nn import AvgPool2d as _AvgPool2d

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.layers.base import Module

if TYPE_CHECKING:
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet

T = TypeVar("T", bound="SD1UNet | SDXLUNet")
TT2IAdapter = TypeVar("TT2IAdapter", bound="T2IAdapter[Any]")
Device = torch.device
DType = torch.dtype

CONTEXT = "t2iadapter"


class Downsample2d(_AvgPool2d, Module):
    """Average pooling with window = stride = ``scale_factor``."""

    def __init__(self, scale_factor: int) -> None:
        _AvgPool2d.__init__(self, kernel_size=scale_factor, stride=scale_factor)

    def forward(self, x: Tensor) -> Tensor:
        k = self.kernel_size if isinstance(self.kernel_size, int) else self.kernel_size[0]
        if x.is_cuda and x.ndim == 4 and x.shape[1] % (16 // x.element_size()) == 0:
            return B.avg_pool2d(x, k)
        return _AvgPool2d.forward(self, x)


class ResidualBlock(fl.Residual):
    def __init__(self, channels: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        where = {"device": device, "dtype": dtype}
        super().__init__(
            fl.Conv2d(in_channels=channels, out_channels=channels, kernel_size=3, padding=1, **where),
            fl.ReLU(),
            fl.Conv2d(in_channels=channels, out_channels=channels, kernel_size=1, **where),
        )


class ResidualBlocks(fl.Chain):
    """(optional 2x average pooling) -> (1x1 conv when the width changes) -> ``num_residual_blocks`` residual blocks."""

    def __init__(
        self, in_channels: int, out_channels: int, num_residual_blocks: int = 2, downsample: bool = False,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        where = {"device": device, "dtype": dtype}
        resize = Downsample2d(scale_factor=2) if downsample else fl.Identity()
        widen = fl.Identity() if in_channels == out_channels else fl.Conv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=1, **where)
        super().__init__(resize, widen, fl.Chain(ResidualBlock(channels=out_channels, **where) for _ in range(num_residual_blocks)))


class StatefulResidualBlocks(fl.Chain):
    """``ResidualBlocks`` whose result is also appended to the ``t2iadapter.features`` context list."""

    def __init__(
        self, in_channels: int, out_channels: int, num_residual_blocks: int = 2, downsample: bool = False,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        super().__init__(
            ResidualBlocks(in_channels=in_channels, out_channels=out_channels, num_residual_blocks=num_residual_blocks,
                           downsample=downsample, device=device, dtype=dtype),
            fl.SetContext(context=CONTEXT, key="features", callback=self.push),
        )

    def push(self, features: list[Tensor], x: Tensor) -> None:
        features.append(x)


def _encoder_layers(
    in_channels: int, channels: tuple[int, int, int, int], num_residual_blocks: int, downscale_factor: int,
    halving: tuple[bool, bool, bool, bool], where: dict[str, Any],
) -> list[fl.Module]:
    """Space-to-depth by ``downscale_factor``, a 3x3 stem onto the first width, four recorded stages (stage ``n`` halves the
    resolution first if ``halving[n]``, and widens from the previous stage's width), and the read-out of what was recorded."""
    unshuffle = fl.PixelUnshuffle(downscale_factor=downscale_factor)
    stem = fl.Conv2d(in_channels=in_channels * downscale_factor**2, out_channels=channels[0], kernel_size=3, padding=1, **where)
    incoming = (channels[0], *channels[:-1])
    stages = [
        StatefulResidualBlocks(narrow, wide, num_residual_blocks, downsample=halve, **where)
        for narrow, wide, halve in zip(incoming, channels, halving)
    ]
    return [unshuffle, stem, *stages, fl.UseContext(context=CONTEXT, key="features")]


class ConditionEncoder(fl.Chain):
    """SD 1.5 layout: image / 8, then every stage after the first halves again (1/8, 1/16, 1/32, 1/64)."""

    def __init__(
        self, in_channels: int = 3, channels: tuple[int, int, int, int] = (320, 640, 1280, 1280), num_residual_blocks: int = 2,
        downscale_factor: int = 8, scale: float = 1.0, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.scale = scale
        where = {"device": device, "dtype": dtype}
        super().__init__(*_encoder_layers(in_channels, channels, num_residual_blocks, downscale_factor, (False, True, True, True), where))

    def init_context(self) -> Contexts:
        return {CONTEXT: {"features": []}}


class ConditionEncoderXL(ConditionEncoder, fl.Chain):
    """SDXL layout: image / 16, and only the third stage halves (1/16, 1/16, 1/32, 1/32): SDXL's UNet has three levels."""

    def __init__(
        self, in_channels: int = 3, channels: tuple[int, int, int, int] = (320, 640, 1280, 1280), num_residual_blocks: int = 2,
        downscale_factor: int = 16, scale: float = 1.0, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.scale = scale
        where = {"device": device, "dtype": dtype}
        fl.Chain.__init__(self, *_encoder_layers(in_channels, channels, num_residual_blocks, downscale_factor, (False, False, True, False), where))


class T2IFeatures(fl.Residual):
    """``x + scale * condition_features_<name>[index]``."""

    def __init__(self, name: str, index: int, scale: float = 1.0) -> None:
        self.name = name
        self.index = index
        self.scale = scale
        super().__init__(
            fl.UseContext(context=CONTEXT, key=f"condition_features_{self.name}").compose(func=lambda features: self.scale * features[self.index])
        )

    def forward(self, *inputs: Any) -> Any:
        x = inputs[0]
        if len(inputs) == 1 and isinstance(x, Tensor) and x.is_cuda and not self._forward_hooks and not self[0]._forward_hooks:
            feature = self.use_context(CONTEXT)[f"condition_features_{self.name}"][self.index]
            if isinstance(feature, Tensor) and feature.shape == x.shape and feature.dtype == x.dtype:
                return B.add(x, feature, alpha=float(self.scale))  # one launch instead of a scalar multiply and an add
        return super().forward(*inputs)


class T2IAdapter(Generic[T], fl.Chain, Adapter[T]):
    """Base of the per-family adapters.  A family states WHERE its four feature maps enter the UNet:

      entry_blocks        indices into ``DownBlocks``: a `T2IFeatures` goes in front of that block's skip-connection tap
      into_middle_block   whether the last feature map is appended to ``MiddleBlock`` instead (SDXL: three encoder levels)

    and everything else - placing / removing the `T2IFeatures`, the scale, the condition context - is shared here."""

    entry_blocks: tuple[int, ...] = ()
    into_middle_block: bool = False

    _condition_encoder: list[ConditionEncoder]  # list-wrapped: the encoder is not a sub-module of the adapted UNet
    _features: list[T2IFeatures] = []

    def __init__(
        self, target: T, name: str, condition_encoder: ConditionEncoder, weights: dict[str, Tensor] | None = None, scale: float | None = None,
    ) -> None:
        self.name = name
        if scale is not None or not self._features:
            count = len(self.entry_blocks) + int(self.into_middle_block)
            self._features = [T2IFeatures(name=name, index=i, scale=1.0 if scale is None else scale) for i in range(count)]
        self.residual_indices = self.entry_blocks
        if weights is not None:
            condition_encoder.load_state_dict(weights)
        self._condition_encoder = [condition_encoder]
        with self.setup_adapter(target):
            super().__init__(target)

    # -- placement ---------------------------------------------------------------------------------------
    def _entry_points(self) -> list[tuple[fl.Chain, T2IFeatures, bool]]:
        """(block, feature layer, goes in front of the block's ResidualAccumulator?) for every feature map."""
        from refiners_b200.foundationals.latent_diffusion.unet_blocks import ResidualAccumulator  # noqa: F401

        points = [(self.target.layer(("DownBlocks", n), fl.Chain), layer, True) for n, layer in zip(self.entry_blocks, self._features)]
        if self.into_middle_block:
            points.append((self.target.layer("MiddleBlock", fl.Chain), self._features[-1], False))
        return points

    def inject(self: TT2IAdapter, parent: fl.Chain | None = None) -> TT2IAdapter:
        from refiners_b200.foundationals.latent_diffusion.unet_blocks import ResidualAccumulator

        for block, layer, before_tap in self._entry_points():
            for present in block.layers(layer_type=T2IFeatures):
                assert present.name != self.name, f"T2I-Adapter named {self.name} is already injected"
            if before_tap:
                block.insert_before_type(ResidualAccumulator, layer)
            else:
                block.append(layer)
        return super().inject(parent)

    def eject(self) -> None:
        for block, layer, _ in self._entry_points():
            block.remove(layer)
        super().eject()

    # -- condition ---------------------------------------------------------------------------------------
    @property
    def condition_encoder(self) -> ConditionEncoder:
        return self._condition_encoder[0]

    def compute_condition_features(self, condition: Tensor) -> tuple[Tensor, ...]:
        return self.condition_encoder(condition)

    def set_condition_features(self, features: tuple[Tensor, ...]) -> None:
        self.set_context(CONTEXT, {f"condition_features_{self.name}": features})

    def init_context(self) -> Contexts:
        return {CONTEXT: {f"condition_features_{self.name}": None}}

    @property
    def scale(self) -> float:
        return self._features[0].scale

    @scale.setter
    def scale(self, value: float) -> None:
        for feature in self._features:
            feature.scale = value

    def structural_copy(self: TT2IAdapter) -> TT2IAdapter:
        raise RuntimeError("T2I-Adapter cannot be copied, eject it first.")
