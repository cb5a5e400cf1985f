"""SD 1.5 ControlNet: a trainable copy of the UNet encoder that turns a condition image into 13
residual corrections, added to the UNet's skip connections (slots 0..11) and to its middle block
(slot 12).

Mirrors ``foundationals/latent_diffusion/stable_diffusion_1/controlnet.py:16-230`` of the reference:
same module tree (hence the same state-dict keys, e.g.
``Controlnet.DownBlocks.Chain_3.Passthrough.Conv2d.weight``), same context contract
(``controlnet.condition_<name>``, ``range_adapter.timestep_embedding_<name>``, ``unet.residuals``),
same inject / eject behaviour (the control copy becomes child 0 of the UNet).  The kernels are the
UNet's: every conv / GroupNorm / attention of the copy runs through ``librefiners_b200.so``.
"""

from __future__ import annotations

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.context import Contexts
from refiners_b200.foundationals.latent_diffusion.range_adapter import RangeAdapter2d
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import (
    DownBlocks,
    MiddleBlock,
    SD1UNet,
    TimestepEncoder,
)
from refiners_b200.foundationals.latent_diffusion.unet_blocks import ResidualBlock

Device = torch.device
DType = torch.dtype

NUM_RESIDUALS = 13  # 12 encoder skips + the middle block


def _forwarded(name: str) -> property:
    """A property of the adapter that lives on its control copy (``scale``, ``scale_decay``)."""

    def read(self: "SD1ControlnetAdapter") -> float:
        return getattr(self.controlnet, name)

    def write(self: "SD1ControlnetAdapter", value: float) -> None:
        setattr(self.controlnet, name, value)

    return property(read, write)


class ConditionEncoder(fl.Chain):
    """``[B, 3, H, W]`` condition image -> ``[B, 320, H/8, W/8]`` (the UNet's first feature map)."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.out_channels = (16, 32, 96, 256)

        def conv(cin: int, cout: int, stride: int = 1) -> fl.Conv2d:
            return fl.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, device=device, dtype=dtype)

        widths = self.out_channels
        stem = fl.Chain(conv(3, widths[0]), fl.SiLU())
        halvings = [
            fl.Chain(conv(narrow, narrow), fl.SiLU(), conv(narrow, wide, stride=2), fl.SiLU())
            for narrow, wide in zip(widths, widths[1:])
        ]
        super().__init__(stem, *halvings, conv(widths[-1], 320))


class Controlnet(fl.Passthrough):
    """Half UNet (timestep encoder, down blocks, middle block) whose 13 zero-convolution taps accumulate
    ``tap(x) * scale * scale_decay ** (12 - n)`` into ``unet.residuals[n]``.  A Passthrough: the UNet
    input flows on unchanged to the UNet proper."""

    scale_decays: list[float]

    def __init__(
        self,
        name: str,
        scale: float = 1.0,
        scale_decay: float = 1.0,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.name = name
        self.scale = scale
        self._scale_decay = scale_decay
        self.compute_scale_decays()
        on = dict(device=device, dtype=dtype)
        self._timestep_key = f"timestep_embedding_{name}"
        super().__init__(
            TimestepEncoder(context_key=self._timestep_key, **on),
            fl.Slicing(dim=1, end=4),  # inpainting UNets feed 9 channels; the control copy sees the latents only
            DownBlocks(in_channels=4, **on),
            MiddleBlock(**on),
        )
        self._join_condition(on)
        self._condition_on_timestep(on)
        self._tap_every_level(on)

    # -- construction steps ------------------------------------------------------------------------------
    def _join_condition(self, on: dict) -> None:
        """The encoded condition is added right after the input convolution (recomputed every step, like the reference)."""
        encoded = fl.Residual(fl.UseContext("controlnet", f"condition_{self.name}"), ConditionEncoder(**on))
        self.layer(("DownBlocks", 0), fl.Chain).append(encoded)

    def _condition_on_timestep(self, on: dict) -> None:
        for block in self.layers(ResidualBlock):
            inner = block.layer("Chain", fl.Chain)
            RangeAdapter2d(
                target=inner.layer("Conv2d_1", fl.Conv2d), channels=block.out_channels, embedding_dim=1280,
                context_key=self._timestep_key, **on,
            ).inject(inner)

    def _tap_every_level(self, on: dict) -> None:
        """A zero convolution + accumulator at the end of each of the 12 encoder entries and of the middle block."""
        levels: list[tuple[fl.Chain, int]] = []
        for n, entry in enumerate(self.layer("DownBlocks", DownBlocks)):
            assert isinstance(entry, fl.Chain)
            width = getattr(entry[0], "out_channels", None)
            assert isinstance(width, int), f"first layer of DownBlocks entry {n} does not expose out_channels: {entry[0]}"
            levels.append((entry, width))
        levels.append((self.layer("MiddleBlock", MiddleBlock), 1280))
        assert len(levels) == NUM_RESIDUALS
        for slot, (chain, width) in enumerate(levels):
            zero_conv = fl.Conv2d(width, width, kernel_size=1, **on)
            chain.append(fl.Passthrough(zero_conv, fl.Lambda(self._accumulate_into(slot))))

    def _accumulate_into(self, n: int):
        def _store_residual(x: Tensor) -> Tensor:  # the name shows in repr(): Lambda(_store_residual(x))
            residuals = self.use_context("unet")["residuals"]
            residuals[n] = residuals[n] + x * self.scale * self.scale_decays[n]
            return x

        return _store_residual

    # -- the per-level weights -----------------------------------------------------------------------------
    @property
    def scale_decay(self) -> float:
        return self._scale_decay

    @scale_decay.setter
    def scale_decay(self, value: float) -> None:
        self._scale_decay = value
        self.compute_scale_decays()

    def compute_scale_decays(self) -> None:
        # 1.0 on the middle block, decaying towards the shallow skips ("prompt is more important" mode at 0.825)
        deepest = NUM_RESIDUALS - 1
        self.scale_decays = [self._scale_decay ** float(deepest - level) for level in range(NUM_RESIDUALS)]


class SD1ControlnetAdapter(fl.Chain, Adapter[SD1UNet]):
    scale = _forwarded("scale")
    scale_decay = _forwarded("scale_decay")

    def __init__(
        self,
        target: SD1UNet,
        name: str,
        scale: float = 1.0,
        scale_decay: float = 1.0,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        self.name = name
        copy = Controlnet(name=name, scale=scale, scale_decay=scale_decay, device=target.device, dtype=target.dtype)
        if weights is not None:
            copy.load_state_dict(weights)
        self._controlnet: list[Controlnet] = [copy]  # in a list: not a registered sub-module
        with self.setup_adapter(target):
            super().__init__(target)

    @property
    def controlnet(self) -> Controlnet:
        return self._controlnet[0]

    def inject(self, parent: fl.Chain | None = None) -> "SD1ControlnetAdapter":
        for child in self.target:  # control copies sit at the top level of the UNet, in front of everything else
            if isinstance(child, Controlnet):
                assert child is not self.controlnet, f"{self.controlnet} is already injected"
                assert child.name != self.name, f"Controlnet named {self.name} is already injected"
        self.target.insert(0, self.controlnet)
        return super().inject(parent)

    def eject(self) -> None:
        self.target.remove(self.controlnet)
        super().eject()

    def init_context(self) -> Contexts:
        return {"controlnet": {f"condition_{self.name}": None}}

    def set_controlnet_condition(self, condition: Tensor) -> None:
        self.set_context("controlnet", {f"condition_{self.name}": condition})

    def structural_copy(self) -> "SD1ControlnetAdapter":
        raise RuntimeError("Controlnet cannot be copied, eject it first.")
