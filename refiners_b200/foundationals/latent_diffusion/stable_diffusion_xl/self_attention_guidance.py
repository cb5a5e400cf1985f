"""Where the SAG probes go in an SDXL UNet (contract:
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/self_attention_guidance.py:11-41)."""

from __future__ import annotations

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.latent_diffusion.self_attention_guidance import SAGAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.self_attention_guidance import place_probes, remove_probes
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import MiddleBlock, SDXLUNet


class SDXLSAGAdapter(SAGAdapter[SDXLUNet]):
    def __init__(self, target: SDXLUNet, scale: float = 1.0, kernel_size: int = 9, sigma: float = 1.0) -> None:
        super().__init__(target=target, scale=scale, kernel_size=kernel_size, sigma=sigma)

    def inject(self: "SDXLSAGAdapter", parent: fl.Chain | None = None) -> "SDXLSAGAdapter":
        place_probes(self.target.ensure_find(MiddleBlock))
        return super().inject(parent)

    def eject(self) -> None:
        remove_probes(self.target.ensure_find(MiddleBlock))
        super().eject()
