"""T2I-Adapter on an SDXL UNet: three feature maps in front of the skip-connection taps of encoder blocks 3, 5 and 8, the
fourth at the end of the middle block, which has no tap of its own (contract:
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/t2i_adapter.py:8-49)."""

from __future__ import annotations

from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import MiddleBlock, SDXLUNet
from refiners_b200.foundationals.latent_diffusion.t2i_adapter import ConditionEncoderXL, T2IAdapter, T2IFeatures
from refiners_b200.foundationals.latent_diffusion.unet_blocks import ResidualAccumulator


class SDXLT2IAdapter(T2IAdapter[SDXLUNet]):
    def __init__(
        self, target: SDXLUNet, name: str, condition_encoder: ConditionEncoderXL | None = None, scale: float = 1.0,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        self.residual_indices = (3, 5, 8)
        self._features = [T2IFeatures(name=name, index=i, scale=scale) for i in range(4)]
        super().__init__(
            target=target, name=name, weights=weights,
            condition_encoder=condition_encoder or ConditionEncoderXL(device=target.device, dtype=target.dtype),
        )

    def _blocks(self) -> list[fl.Chain]:
        return [self.target.layer(("DownBlocks", n), fl.Chain) for n in self.residual_indices]

    def inject(self: "SDXLT2IAdapter", parent: fl.Chain | None = None) -> "SDXLT2IAdapter":
        for block, feature in zip(self._blocks(), self._features):  # three blocks, four features: the last goes below
            self._claim(block)
            block.insert_before_type(ResidualAccumulator, feature)
        middle = self.target.layer("MiddleBlock", MiddleBlock)
        self._claim(middle)
        middle.append(self._features[-1])
        return super().inject(parent)

    def eject(self: "SDXLT2IAdapter") -> None:
        for block, feature in zip(self._blocks(), self._features):
            block.remove(feature)
        self.target.layer("MiddleBlock", MiddleBlock).remove(self._features[-1])
        super().eject()
