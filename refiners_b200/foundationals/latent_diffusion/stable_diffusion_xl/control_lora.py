"""ControlLora - SDXL's ControlNet equivalent (BASELINE config 4).

Contract (class names, constructor signatures, resulting tree = ``repr`` and state-dict keys, context names, checkpoint
key conventions) from
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/control_lora.py:
`ConditionEncoder` :14-87, `ZeroConvolution` :90-132, `ControlLora` :144-248, `ControlLoraAdapter` :251-411; the
reference's tests/adapters/test_control_lora.py runs against this module (tests/test_reference_own_tests.py).

What a ControlLora is.  A *structural copy* of the UNet's encoder half (TimestepEncoder + DownBlocks + MiddleBlock):
the Chain skeleton is duplicated, the weighted leaves are the UNet's own tensors.  The copy is then specialised - it gets
its own timestep-embedding slot, a ``ConditionEncoder`` that injects the condition image (3 x 1024 x 1024 ->
320 x 128 x 128) at the end of its first block, LoRAs on its (shared) leaves, and, where the UNet stores skip
connections, ``ZeroConvolution``s that instead ADD a scaled 1x1 projection into those same slots.  Placed in front of
the UNet's children it therefore runs first on every step and pre-loads ``unet.residuals`` with corrections which the
UNet's own ``ResidualAccumulator``s then add to (+45 % FLOPs).

Because leaves are shared between two trees that adapt them differently, packed-weight caches are keyed by the weight
tensor object and merged-LoRA weights by (weight, factors, scales): refiners_b200.backend._PackCache.  The encoded
condition is step-invariant and is hoisted out of the per-step CUDA graph (engine.graph).
"""

from __future__ import annotations

from typing import Iterator

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
    """Condition image -> feature map at 1/8 resolution: a 3x3 stem, one (3x3, SiLU, stride-2 3x3, SiLU) stage per
    further entry of ``intermediate_channels``, and a 3x3 projection to ``out_channels``."""

    def __init__(
        self,
        in_channels: int = 3,
        out_channels: int = 320,
        intermediate_channels: tuple[int, ...] = (16, 32, 96, 256),
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        def conv(cin: int, cout: int, stride: int = 1) -> Conv2d:
            return Conv2d(cin, cout, kernel_size=3, stride=stride, padding=1, device=device, dtype=dtype)

        widths = intermediate_channels
        stages = [Chain(conv(narrow, narrow), SiLU(), conv(narrow, wide, stride=2), SiLU()) for narrow, wide in zip(widths, widths[1:])]
        super().__init__(Chain(conv(in_channels, widths[0]), SiLU()), *stages, conv(widths[-1], out_channels))


class ZeroConvolution(Passthrough):
    """``residuals[residual_index] += scale * conv1x1(x)``; the activation itself passes through unchanged.
    (Zero-initialised in a trained checkpoint's early life, hence the name.)"""

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
        encoder_half = [unet.layer(part, Chain).structural_copy() for part in ("TimestepEncoder", "DownBlocks", "MiddleBlock")]
        super().__init__(*encoder_half)
        timestep_encoder, down_blocks, middle_block = encoder_half
        where = {"device": unet.device, "dtype": unet.dtype}

        # (1) a private timestep-embedding slot: the copy's TimestepEncoder writes it, the copy's RangeAdapter2ds read it,
        #     and neither disturbs the UNet's own
        slot = f"timestep_embedding_control_lora_{name}"
        timestep_encoder.context_key = slot  # type: ignore[attr-defined]
        for reader in self.layers(RangeAdapter2d):
            reader.context_key = slot

        # (2) the encoded condition is added at the END of the first block (after that block's residual tap)
        first = down_blocks.layer(0, Chain)
        first.append(
            Residual(
                UseContext(f"control_lora_{name}", "condition"),
                ConditionEncoder(in_channels=condition_channels, out_channels=first.layer(0, Conv2d).out_channels, **where),
            )
        )

        # (3) where the UNet's encoder stores a skip connection, the copy contributes a correction instead
        for tap in [*self.layers(ResidualAccumulator)]:
            block = self.ensure_find_parent(tap)
            head = block[0]
            assert hasattr(head, "out_channels"), f"{head} has no out_channels attribute"
            width = head.out_channels
            assert isinstance(width, int)
            block.replace(tap, ZeroConvolution(width, width, residual_index=tap.n, scale=scale, **where))
        #     ... and one more after the middle block, into the slot the UNet adds right after its own middle block
        width = middle_block.layer(0, ResidualBlock).out_channels
        middle_block.append(ZeroConvolution(width, width, residual_index=len(down_blocks), scale=scale, **where))

    def _taps(self) -> Iterator[ZeroConvolution]:
        return self.layers(ZeroConvolution)

    @property
    def scale(self) -> float:
        return next(iter(self._taps())).scale

    @scale.setter
    def scale(self, value: float) -> None:
        for tap in self._taps():
            tap.scale = value


def _section(state_dict: dict[str, Tensor], tag: str) -> dict[str, Tensor]:
    """Entries of a ControlLora checkpoint whose key mentions ``tag``, with the ``<tag>.`` prefix removed."""
    return {key.removeprefix(f"{tag}."): value for key, value in state_dict.items() if tag in key}


class ControlLoraAdapter(Chain, Adapter[SDXLUNet]):
    """Puts a ``ControlLora`` in front of an SDXL UNet's children; the condition image arrives through the context
    ``control_lora_<name>.condition``."""

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
            # list-wrapped: the control copy becomes a child of the TARGET on inject, not of this adapter
            self._control_lora = [ControlLora(name=name, unet=target, scale=scale, condition_channels=condition_channels)]
            super().__init__(target)
        if weights:
            self.load_weights(weights)

    @property
    def control_lora(self) -> ControlLora:
        return self._control_lora[0]

    def init_context(self) -> Contexts:
        return {f"control_lora_{self.name}": {"condition": None}}

    def set_condition(self, condition: Tensor) -> None:
        self.set_context(f"control_lora_{self.name}", {"condition": condition})

    @property
    def scale(self) -> float:
        return self.control_lora.scale

    @scale.setter
    def scale(self, value: float) -> None:
        self.control_lora.scale = value

    # -- placement ------------------------------------------------------------------------------------
    def inject(self, parent: Chain | None = None) -> "ControlLoraAdapter":
        self.target.insert(index=0, module=self.control_lora)  # runs before everything else of the UNet
        return super().inject(parent)

    def eject(self) -> None:
        self.target.remove(self.control_lora)
        super().eject()

    def structural_copy(self) -> "ControlLoraAdapter":
        raise RuntimeError("ControlLoraAdapter cannot be copied, eject it first.")

    # -- checkpoints ----------------------------------------------------------------------------------------
    def load_weights(self, state_dict: dict[str, Tensor]) -> None:
        """A ControlLora checkpoint has three sections: ``ControlLora.<leaf path>.{down,up}`` LoRA factors,
        ``ZeroConvolution_NN.*`` and ``ConditionEncoder.*``."""
        control = self.control_lora
        self.load_lora_layers(self.name, state_dict, control)
        self.load_zero_convolution_layers(state_dict, control)
        self.load_condition_encoder(state_dict, control)

    @staticmethod
    def load_lora_layers(name: str, state_dict: dict[str, Tensor], control_lora: ControlLora) -> None:
        """One ``LoraAdapter`` per addressed leaf, injected inside the control copy only (the leaf itself stays shared
        with the UNet, which keeps running it unadapted)."""
        factors = {
            f"{path}.weight": tensor.to(dtype=control_lora.dtype, device=control_lora.device)
            for path, tensor in _section(state_dict, "ControlLora").items()
        }
        pending: list[LoraAdapter] = []
        for path, lora in Lora.from_dict(name, state_dict=factors).items():
            leaf = control_lora.layer(path.split("."), WeightedModule)
            assert lora.is_compatible(leaf)
            pending.append(LoraAdapter(leaf, lora))
        for adapter in pending:  # all leaves are resolved before the first injection changes any path
            adapter.inject(control_lora)

    @staticmethod
    def load_zero_convolution_layers(state_dict: dict[str, Tensor], control_lora: ControlLora) -> None:
        for number, tap in enumerate(control_lora.layers(ZeroConvolution), start=1):
            tap.load_state_dict(_section(state_dict, f"ZeroConvolution_{number:02d}"))

    @staticmethod
    def load_condition_encoder(state_dict: dict[str, Tensor], control_lora: ControlLora) -> None:
        control_lora.ensure_find(ConditionEncoder).load_state_dict(_section(state_dict, "ConditionEncoder"))
