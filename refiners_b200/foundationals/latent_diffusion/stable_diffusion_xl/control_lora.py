"""ControlLora - SDXL's ControlNet equivalent (BASELINE config 4).

Behaviour follows
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/control_lora.py:
`ConditionEncoder` :14-87, `ZeroConvolution` :90-132, `ControlLora` :144-248, `ControlLoraAdapter` :251-411.

A `ControlLora` is a *structural copy* of the UNet's TimestepEncoder + DownBlocks + MiddleBlock: the
Chain skeleton is duplicated, the weighted leaves are the UNet's own (shared storage); LoRAs are then
attached to those shared leaves inside the copy only, a `ConditionEncoder` embeds the condition image
(3 x 1024 x 1024 -> 320 x 128 x 128) and `ZeroConvolution`s accumulate into `unet.residuals[n]`, which
the main UNet's `ResidualAccumulator`s then add to.  It is inserted at index 0 of the UNet, so every
denoising step runs this half-UNet first (+45 % FLOPs).

Because leaves are shared between two trees with different adapters, packed-weight caches are keyed
by the leaf's weight tensor object and LoRA packs by the adapter's own factor tensors
(refiners_b200/backend `_PackCache`).
"""

from __future__ import annotations

import torch
from torch import Tensor

from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.adapters.lora import Lora, LoraAdapter
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.layers import Chain, Conv2d, Multiply, Passthrough, Residual, SiLU, UseContext, WeightedModule
from refiners_b200.foundationals.latent_diffusion.range_adapter import RangeAdapter2d
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet
from refiners_b200.foundationals.latent_diffusion.unet_blocks import ResidualAccumulator, ResidualBlock

Device = torch.device
DType = torch.dtype


class ConditionEncoder(Chain):
    """conv-SiLU stem, three (conv, SiLU, stride-2 conv, SiLU) stages, output conv."""

    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 320,
        intermediate_channels: tuple[int, ...] = (16, 32, 96, 256),
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        kw = dict(device=device, dtype=dtype)
        c = intermediate_channels
        super().__init__(
            Chain(Conv2d(in_channels, c[0], kernel_size=3, stride=1, padding=1, **kw), SiLU()),
            *(
                Chain(
                    Conv2d(c[i], c[i], kernel_size=3, padding=1, **kw),
                    SiLU(),
                    Conv2d(c[i], c[i + 1], kernel_size=3, stride=2, padding=1, **kw),
                    SiLU(),
                )
                for i in range(len(c) - 1)
            ),
            Conv2d(c[-1], out_channels, kernel_size=3, padding=1, **kw),
        )


class ZeroConvolution(Passthrough):
    """``residuals[n] += scale * conv1x1(x)``; hands x through."""

    def __init__(
        self,
        in_channels: int,
        out_channels: int,
        residual_index: int,
        scale: float = 1.0,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self._scale = scale
        super().__init__(
            Conv2d(in_channels, out_channels, kernel_size=1, device=device, dtype=dtype),
            Multiply(scale=scale),
            ResidualAccumulator(n=residual_index),
        )

    @property
    def scale(self) -> float:
        return self._scale

    @scale.setter
    def scale(self, value: float) -> None:
        self._scale = value
        self.ensure_find(Multiply).scale = value


class ControlLora(Passthrough):
    def __init__(self, name: str, unet: SDXLUNet, scale: float = 1.0, condition_channels: int = 3) -> None:
        self.name = name
        timestep_encoder = unet.layer("TimestepEncoder", Chain).structural_copy()
        downblocks = unet.layer("DownBlocks", Chain).structural_copy()
        middle_block = unet.layer("MiddleBlock", Chain).structural_copy()
        super().__init__(timestep_encoder, downblocks, middle_block)

        # its own timestep-embedding slot, so the copy and the UNet do not overwrite each other
        key = f"timestep_embedding_control_lora_{name}"
        timestep_encoder.context_key = key  # type: ignore[attr-defined]
        for range_adapter in self.layers(RangeAdapter2d):
            range_adapter.context_key = key

        first = downblocks.layer(0, Chain)
        stem_channels = first.layer(0, Conv2d).out_channels
        first.append(
            Residual(
                UseContext(f"control_lora_{name}", "condition"),
                ConditionEncoder(in_channels=condition_channels, out_channels=stem_channels, device=unet.device, dtype=unet.dtype),
            )
        )
        for accumulator in list(self.layers(ResidualAccumulator)):
            block = self.ensure_find_parent(accumulator)
            head = block[0]
            assert hasattr(head, "out_channels"), f"{head} has no out_channels attribute"
            channels = head.out_channels
            assert isinstance(channels, int)
            block.replace(
                accumulator,
                ZeroConvolution(channels, channels, residual_index=accumulator.n, scale=scale, device=unet.device, dtype=unet.dtype),
            )
        mid_channels = middle_block.layer(0, ResidualBlock).out_channels
        middle_block.append(
            ZeroConvolution(mid_channels, mid_channels, residual_index=len(downblocks), scale=scale, device=unet.device, dtype=unet.dtype)
        )

    @property
    def scale(self) -> float:
        return self.ensure_find(ZeroConvolution).scale

    @scale.setter
    def scale(self, value: float) -> None:
        for zero_conv in self.layers(ZeroConvolution):
            zero_conv.scale = value


class ControlLoraAdapter(Chain, Adapter[SDXLUNet]):
    def __init__(
        self,
        name: str,
        target: SDXLUNet,
        scale: float = 1.0,
        condition_channels: int = 3,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        with self.setup_adapter(target):
            self.name = name
            self._control_lora = [ControlLora(name=name, unet=target, scale=scale, condition_channels=condition_channels)]
            super().__init__(target)
        if weights:
            self.load_weights(weights)

    @property
    def control_lora(self) -> ControlLora:
        return self._control_lora[0]

    def init_context(self) -> Contexts:
        return {f"control_lora_{self.name}": {"condition": None}}

    def inject(self, parent: Chain | None = None) -> "ControlLoraAdapter":
        self.target.insert(index=0, module=self.control_lora)
        return super().inject(parent)

    def eject(self) -> None:
        self.target.remove(self.control_lora)
        return super().eject()

    def structural_copy(self) -> "ControlLoraAdapter":
        raise RuntimeError("ControlLoraAdapter cannot be copied, eject it first.")

    @property
    def scale(self) -> float:
        return self.control_lora.scale

    @scale.setter
    def scale(self, value: float) -> None:
        self.control_lora.scale = value

    def set_condition(self, condition: Tensor) -> None:
        self.set_context(f"control_lora_{self.name}", {"condition": condition})

    # -- weights -------------------------------------------------------------------------------
    def load_weights(self, state_dict: dict[str, Tensor]) -> None:
        self.load_lora_layers(self.name, state_dict, self.control_lora)
        self.load_zero_convolution_layers(state_dict, self.control_lora)
        self.load_condition_encoder(state_dict, self.control_lora)

    @staticmethod
    def load_lora_layers(name: str, state_dict: dict[str, Tensor], control_lora: ControlLora) -> None:
        """Keys ``ControlLora.<path to leaf>.{down,up}`` -> one LoraAdapter per addressed leaf, injected
        inside the control copy only."""
        weights = {
            f"{key.removeprefix('ControlLora.')}.weight": value.to(dtype=control_lora.dtype, device=control_lora.device)
            for key, value in state_dict.items()
            if "ControlLora" in key
        }
        adapters: list[LoraAdapter] = []
        for key, lora in Lora.from_dict(name, state_dict=weights).items():
            leaf = control_lora.layer(key.split("."), WeightedModule)
            assert lora.is_compatible(leaf)
            adapters.append(LoraAdapter(leaf, lora))
        for adapter in adapters:
            adapter.inject(control_lora)

    @staticmethod
    def load_zero_convolution_layers(state_dict: dict[str, Tensor], control_lora: ControlLora) -> None:
        for i, zero_conv in enumerate(control_lora.layers(ZeroConvolution)):
            tag = f"ZeroConvolution_{i + 1:02d}"
            zero_conv.load_state_dict({k.removeprefix(f"{tag}."): v for k, v in state_dict.items() if tag in k})

    @staticmethod
    def load_condition_encoder(state_dict: dict[str, Tensor], control_lora: ControlLora) -> None:
        encoder = control_lora.ensure_find(ConditionEncoder)
        encoder.load_state_dict({k.removeprefix("ConditionEncoder."): v for k, v in state_dict.items() if "ConditionEncoder" in k})
