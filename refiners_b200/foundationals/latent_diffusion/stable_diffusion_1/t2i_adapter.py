"""T2I-Adapter on an SD 1.5 UNet: one feature map in front of the skip-connection tap of the last block of each of the four
encoder levels (contract: /root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_1/t2i_adapter.py:8-41).
The placement logic itself lives in `T2IAdapter`."""

from __future__ import annotations

from torch import Tensor

from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet
from refiners_b200.foundationals.latent_diffusion.t2i_adapter import ConditionEncoder, T2IAdapter


class SD1T2IAdapter(T2IAdapter[SD1UNet]):
    entry_blocks = (2, 5, 8, 11)
    into_middle_block = False

    def __init__(
        self, target: SD1UNet, name: str, condition_encoder: ConditionEncoder | None = None, scale: float = 1.0,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        encoder = condition_encoder or ConditionEncoder(device=target.device, dtype=target.dtype)
        super().__init__(target=target, name=name, condition_encoder=encoder, weights=weights, scale=scale)
