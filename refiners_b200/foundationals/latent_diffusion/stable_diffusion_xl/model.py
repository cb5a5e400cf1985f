"""SDXL pipeline step (UNet + solver).

Follows /root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/model.py:22-162.
The text encoders and VAE are optional (run once per prompt/image, outside the per-step path).
"""

from __future__ import annotations

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder
from refiners_b200.foundationals.latent_diffusion.model import LatentDiffusionModel
from refiners_b200.foundationals.latent_diffusion.solvers import DDIM, Solver
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet


class SDXLAutoencoder(LatentDiffusionAutoencoder):
    """The SDXL VAE: latents = 0.13025 x encoder output (reference stable_diffusion_xl/model.py:12-19)."""

    encoder_scale: float = 0.13025


class StableDiffusion_XL(LatentDiffusionModel):
    unet: SDXLUNet

    def __init__(
        self, unet: SDXLUNet | None = None, lda: fl.Chain | None = None, clip_text_encoder: fl.Chain | None = None,
        solver: Solver | None = None, device: torch.device | str = "cpu", dtype: torch.dtype = torch.float32,
    ) -> None:
        super().__init__(
            unet=unet or SDXLUNet(in_channels=4),
            lda=lda,
            clip_text_encoder=clip_text_encoder,
            solver=solver or DDIM(num_inference_steps=30),
            device=device,
            dtype=dtype,
        )

    def __call__(  # type: ignore[override]
        self,
        x: Tensor,
        step: int,
        *,
        clip_text_embedding: Tensor,
        pooled_text_embedding: Tensor,
        time_ids: Tensor,
        condition_scale: float = 5.0,
    ) -> Tensor:
        return super().__call__(
            x=x,
            step=step,
            clip_text_embedding=clip_text_embedding,
            pooled_text_embedding=pooled_text_embedding,
            time_ids=time_ids,
            condition_scale=condition_scale,
        )

    @property
    def default_time_ids(self) -> Tensor:
        time_ids = torch.tensor([1024, 1024, 0, 0, 1024, 1024], device=self.device)
        return time_ids.repeat(2 if self.classifier_free_guidance else 1, 1)

    def set_unet_context(  # type: ignore[override]
        self,
        *,
        timestep: Tensor,
        clip_text_embedding: Tensor,
        pooled_text_embedding: Tensor,
        time_ids: Tensor,
        **_: Tensor,
    ) -> None:
        self.unet.set_timestep(timestep=timestep)
        self.unet.set_clip_text_embedding(clip_text_embedding=clip_text_embedding)
        self.unet.set_pooled_text_embedding(pooled_text_embedding=pooled_text_embedding)
        self.unet.set_time_ids(time_ids=time_ids)

    def _sag_adapter_type(self) -> type:
        from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.self_attention_guidance import SDXLSAGAdapter

        return SDXLSAGAdapter

    def _unconditional_half(  # type: ignore[override]
        self, *, clip_text_embedding: Tensor, pooled_text_embedding: Tensor, time_ids: Tensor, **_: Tensor
    ) -> dict[str, Tensor]:
        return {
            "clip_text_embedding": clip_text_embedding.chunk(2)[0],
            "pooled_text_embedding": pooled_text_embedding.chunk(2)[0],
            "time_ids": time_ids.chunk(2)[0],
        }

    def forward(  # type: ignore[override]
        self,
        x: Tensor,
        step: int,
        *,
        clip_text_embedding: Tensor,
        pooled_text_embedding: Tensor,
        time_ids: Tensor,
        condition_scale: float = 5.0,
        **kwargs: Tensor,
    ) -> Tensor:
        return super().forward(
            x=x,
            step=step,
            clip_text_embedding=clip_text_embedding,
            pooled_text_embedding=pooled_text_embedding,
            time_ids=time_ids,
            condition_scale=condition_scale,
            **kwargs,
        )
