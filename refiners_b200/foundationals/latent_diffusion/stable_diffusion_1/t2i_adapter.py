"""T2I-Adapter on an SD 1.5 UNet: one feature map in front of the skip-connection tap of the last block of each encoder
level (contract: /root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_1/t2i_adapter.py:8-41)."""

from __future__ import annotations

from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet
from refiners_b200.foundationals.latent_diffusion.t2i_adapter import ConditionEncoder, T2IAdapter, T2IFeatures
from refiners_b200.foundationals.latent_diffusion.unet_blocks import ResidualAccumulator


class SD1T2IAdapter(T2IAdapter[SD1UNet]):
    def __init__(
        self, target: SD1UNet, name: str, condition_encoder: ConditionEncoder | None = None, scale: float = 1.0,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        self.residual_indices = (2, 5, 8, 11)
        self._features = [T2IFeatures(name=name, index=i, scale=scale) for i in range(4)]
        super().__init__(
            target=target, name=name, weights=weights,
            condition_encoder=condition_encoder or ConditionEncoder(device=target.device, dtype=target.dtype),
        )

    def _blocks(self) -> list[fl.Chain]:
        return [self.target.layer(("DownBlocks", n), fl.Chain) for n in self.residual_indices]

    def inject(self: "SD1T2IAdapter", parent: fl.Chain | None = None) -> "SD1T2IAdapter":
        for block, feature in zip(self._blocks(), self._features, strict=True):
            self._claim(block)
            block.insert_before_type(ResidualAccumulator, feature)
        return super().inject(parent)

    def eject(self: "SD1T2IAdapter") -> None:
        for block, feature in zip(self._blocks(), self._features, strict=True):
            block.remove(feature)
        super().eject()
