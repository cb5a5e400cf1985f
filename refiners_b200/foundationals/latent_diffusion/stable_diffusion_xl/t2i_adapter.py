"""T2I-Adapter on an SDXL UNet: three feature maps in front of the skip-connection taps of encoder blocks 3, 5 and 8, the
fourth at the end of the middle block, which has no tap of its own (contract:
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/t2i_adapter.py:8-49).  The placement logic
itself lives in `T2IAdapter`."""

from __future__ import annotations

from torch import Tensor

from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet
from refiners_b200.foundationals.latent_diffusion.t2i_adapter import ConditionEncoderXL, T2IAdapter


class SDXLT2IAdapter(T2IAdapter[SDXLUNet]):
    entry_blocks = (3, 5, 8)
    into_middle_block = True

    def __init__(
        self, target: SDXLUNet, name: str, condition_encoder: ConditionEncoderXL | None = None, scale: float = 1.0,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        encoder = condition_encoder or ConditionEncoderXL(device=target.device, dtype=target.dtype)
        super().__init__(target=target, name=name, condition_encoder=encoder, weights=weights, scale=scale)
