"""SD 1.5 pipeline step (UNet + solver); see stable_diffusion_xl/model.py for scope notes.

Follows /root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_1/model.py:25-120.
"""

from __future__ import annotations

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder
from refiners_b200.foundationals.latent_diffusion.model import LatentDiffusionModel
from refiners_b200.foundationals.latent_diffusion.solvers import DDIM, Solver
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet


class SD1Autoencoder(LatentDiffusionAutoencoder):
    """The SD 1.5 VAE: latents = 0.18215 x encoder output (reference stable_diffusion_1/model.py:15-22)."""

    encoder_scale: float = 0.18215


class StableDiffusion_1(LatentDiffusionModel):
    unet: SD1UNet

    def __init__(
        self,
        unet: SD1UNet | None = None,
        lda: fl.Chain | None = None,
        clip_text_encoder: fl.Chain | None = None,
        solver: Solver | None = None,
        device: torch.device | str = "cpu",
        dtype: torch.dtype = torch.float32,
    ) -> None:
        super().__init__(
            unet=unet or SD1UNet(in_channels=4),
            lda=lda,
            clip_text_encoder=clip_text_encoder,
            solver=solver or DDIM(num_inference_steps=30),
            device=device,
            dtype=dtype,
        )

    def set_unet_context(self, *, timestep: Tensor, clip_text_embedding: Tensor, **_: Tensor) -> None:
        self.unet.set_timestep(timestep=timestep)
        self.unet.set_clip_text_embedding(clip_text_embedding=clip_text_embedding)

    def _sag_adapter_type(self) -> type:
        from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.self_attention_guidance import SD1SAGAdapter

        return SD1SAGAdapter
