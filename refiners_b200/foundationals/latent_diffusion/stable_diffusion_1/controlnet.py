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


class ConditionEncoder(fl.Chain):
    """``[B, 3, H, W]`` condition image -> ``[B, 320, H/8, W/8]`` (the UNet's first feature map)."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.out_channels = (16, 32, 96, 256)
        kw = dict(device=device, dtype=dtype)
        c = self.out_channels
        stages = [
            fl.Chain(
                fl.Conv2d(c[i], c[i], kernel_size=3, padding=1, **kw),
                fl.SiLU(),
                fl.Conv2d(c[i], c[i + 1], kernel_size=3, stride=2, padding=1, **kw),
                fl.SiLU(),
            )
            for i in range(len(c) - 1)
        ]
        super().__init__(
            fl.Chain(fl.Conv2d(3, c[0], kernel_size=3, stride=1, padding=1, **kw), fl.SiLU()),
            *stages,
            fl.Conv2d(c[-1], 320, kernel_size=3, padding=1, **kw),
        )


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
        kw = dict(device=device, dtype=dtype)
        temb_key = f"timestep_embedding_{name}"
        super().__init__(
            TimestepEncoder(context_key=temb_key, **kw),
            fl.Slicing(dim=1, end=4),  # inpainting UNets feed 9 channels; the control copy sees the latents only
            DownBlocks(in_channels=4, **kw),
            MiddleBlock(**kw),
        )
        # the encoded condition joins right after the input convolution (recomputed every step, like the reference)
        self.layer(("DownBlocks", 0), fl.Chain).append(
            fl.Residual(fl.UseContext("controlnet", f"condition_{name}"), ConditionEncoder(**kw))
        )
        for block in self.layers(ResidualBlock):
            body = block.layer("Chain", fl.Chain)
            RangeAdapter2d(
                target=body.layer("Conv2d_1", fl.Conv2d),
                channels=block.out_channels,
                embedding_dim=1280,
                context_key=temb_key,
                **kw,
            ).inject(body)
        for n, entry in enumerate(self.layer("DownBlocks", DownBlocks)):
            assert isinstance(entry, fl.Chain)
            width = getattr(entry[0], "out_channels", None)
            assert isinstance(width, int), f"first layer of DownBlocks entry {n} does not expose out_channels: {entry[0]}"
            entry.append(self._tap(width, n, kw))
        self.layer("MiddleBlock", MiddleBlock).append(self._tap(1280, NUM_RESIDUALS - 1, kw))

    def _tap(self, channels: int, n: int, kw: dict) -> fl.Passthrough:
        return fl.Passthrough(fl.Conv2d(channels, channels, kernel_size=1, **kw), fl.Lambda(self._accumulate_into(n)))

    def _accumulate_into(self, n: int):
        def _store_residual(x: Tensor) -> Tensor:  # the name shows in repr(): Lambda(_store_residual(x))
            residuals = self.use_context("unet")["residuals"]
            residuals[n] = residuals[n] + x * self.scale * self.scale_decays[n]
            return x

        return _store_residual

    @property
    def scale_decay(self) -> float:
        return self._scale_decay

    @scale_decay.setter
    def scale_decay(self, value: float) -> None:
        self._scale_decay = value
        self.compute_scale_decays()

    def compute_scale_decays(self) -> None:
        # 1.0 on the middle block, decaying towards the shallow skips ("prompt is more important" mode at 0.825)
        self.scale_decays = [self._scale_decay ** float(NUM_RESIDUALS - 1 - i) for i in range(NUM_RESIDUALS)]


class SD1ControlnetAdapter(fl.Chain, Adapter[SD1UNet]):
    def __init__(
        self,
        target: SD1UNet,
        name: str,
        scale: float = 1.0,
        scale_decay: float = 1.0,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        self.name = name
        controlnet = Controlnet(name=name, scale=scale, scale_decay=scale_decay, device=target.device, dtype=target.dtype)
        if weights is not None:
            controlnet.load_state_dict(weights)
        self._controlnet: list[Controlnet] = [controlnet]  # in a list: not a registered sub-module
        with self.setup_adapter(target):
            super().__init__(target)

    @property
    def controlnet(self) -> Controlnet:
        return self._controlnet[0]

    def inject(self, parent: fl.Chain | None = None) -> "SD1ControlnetAdapter":
        present = [layer for layer in self.target if isinstance(layer, Controlnet)]
        assert self.controlnet not in present, f"{self.controlnet} is already injected"
        assert all(cn.name != self.name for cn in present), f"Controlnet named {self.name} is already injected"
        self.target.insert(0, self.controlnet)
        return super().inject(parent)

    def eject(self) -> None:
        self.target.remove(self.controlnet)
        super().eject()

    def init_context(self) -> Contexts:
        return {"controlnet": {f"condition_{self.name}": None}}

    @property
    def scale(self) -> float:
        return self.controlnet.scale

    @scale.setter
    def scale(self, value: float) -> None:
        self.controlnet.scale = value

    @property
    def scale_decay(self) -> float:
        return self.controlnet.scale_decay

    @scale_decay.setter
    def scale_decay(self, value: float) -> None:
        self.controlnet.scale_decay = value

    def set_controlnet_condition(self, condition: Tensor) -> None:
        self.set_context("controlnet", {f"condition_{self.name}": condition})

    def structural_copy(self) -> "SD1ControlnetAdapter":
        raise RuntimeError("Controlnet cannot be copied, eject it first.")
