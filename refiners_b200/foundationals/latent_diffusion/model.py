"""One denoising step: classifier-free-guidance batching, UNet call, solver update.

Follows /root/reference/src/refiners/foundationals/latent_diffusion/model.py:15-169
(`LatentDiffusionModel.forward` :128-159).  Scope note: the VAE (`lda`) and the CLIP text
encoder run once per image/prompt, outside the per-step hot path, and are not rebuilt here
(SURVEY.md section 8f); they are optional constructor arguments and any fluxion Chain works.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, TypeVar

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.latent_diffusion.solvers import Solver

TModel = TypeVar("TModel", bound="LatentDiffusionModel")
Device = torch.device
DType = torch.dtype


class LatentDiffusionModel(fl.Module, ABC):
    def __init__(
        self, unet: fl.Chain, lda: fl.Chain | None, clip_text_encoder: fl.Chain | None, solver: Solver,
        classifier_free_guidance: bool = True, device: Device | str = "cpu", dtype: DType = torch.float32,
    ) -> None:
        super().__init__()
        self.device: Device = device if isinstance(device, Device) else Device(device)
        self.dtype = dtype
        self.unet = unet.to(device=self.device, dtype=self.dtype)
        self.lda = None if lda is None else lda.to(device=self.device, dtype=self.dtype)
        self.clip_text_encoder = (
            None if clip_text_encoder is None else clip_text_encoder.to(device=self.device, dtype=self.dtype)
        )
        self.solver = solver.to(device=self.device, dtype=self.dtype)
        self.classifier_free_guidance = classifier_free_guidance
        self._graphed_unet: list[object] = []  # list-wrapped: not a torch sub-module

    def enable_cuda_graph(self, enabled: bool = True) -> None:
        """Replay the UNet forward from a captured CUDA graph (refiners_b200.engine.graph).
        The first call after enabling - or after any structural edit - runs the Python walker
        once under capture; contexts keep being set through the normal API."""
        for runner in self._graphed_unet:
            runner.close()  # type: ignore[attr-defined]
        self._graphed_unet = []
        if enabled:
            from refiners_b200.engine.graph import GraphedChain

            self._graphed_unet = [GraphedChain(self.unet)]

    def _run_unet(self, latents: Tensor) -> Tensor:
        if self._graphed_unet and latents.is_cuda:
            return self._graphed_unet[0](latents)  # type: ignore[operator]
        return self.unet(latents)

    def set_inference_steps(self, num_steps: int, first_step: int = 0) -> None:
        self.solver = self.solver.rebuild(num_inference_steps=num_steps, first_inference_step=first_step)

    @staticmethod
    def sample_noise(
        size: tuple[int, ...], device: Device | None = None, dtype: DType | None = None,
        offset_noise: float | None = None,
    ) -> Tensor:
        noise = torch.randn(size=size, device=device, dtype=dtype)
        if offset_noise is not None:
            noise += offset_noise * torch.randn(size=(size[0], size[1], 1, 1), device=device, dtype=dtype)
        return noise

    def init_latents(self, size: tuple[int, int], init_image: Any = None, noise: Tensor | None = None) -> Tensor:
        height, width = size
        lh, lw = height // 8, width // 8
        if noise is None:
            noise = self.sample_noise(size=(1, 4, lh, lw), device=self.device, dtype=self.dtype)
        assert list(noise.shape[2:]) == [lh, lw], f"noise shape is not compatible: {noise.shape}, with size: {size}"
        if init_image is None:
            latent = noise
        else:
            assert self.lda is not None, "image-to-image needs a latent autoencoder (out of the hot path)"
            encoded = self.lda.image_to_latents(init_image.resize(size=(width, height)))  # type: ignore[attr-defined]
            latent = self.solver.add_noise(x=encoded, noise=noise, step=self.solver.first_inference_step)
        return self.solver.scale_model_input(latent, step=-1)

    @property
    def steps(self) -> list[int]:
        return self.solver.inference_steps

    @abstractmethod
    def set_unet_context(self, *, timestep: Tensor, clip_text_embedding: Tensor, **_: Tensor) -> None: ...

    def has_self_attention_guidance(self) -> bool:
        return False

    def compute_self_attention_guidance(
        self, x: Tensor, noise: Tensor, step: int, *, clip_text_embedding: Tensor, **kwargs: Tensor,
    ) -> Tensor:
        raise NotImplementedError("self-attention guidance is out of the hot-path scope (SURVEY.md section 2 #18)")

    def forward(
        self, x: Tensor, step: int, *, clip_text_embedding: Tensor, condition_scale: float = 7.5, **kwargs: Tensor,
    ) -> Tensor:
        cfg = self.classifier_free_guidance
        if cfg:
            assert clip_text_embedding.shape[0] % 2 == 0, f"invalid batch size: {clip_text_embedding.shape[0]}"
        timestep = self.solver.timesteps[step].unsqueeze(dim=0)
        self.set_unet_context(timestep=timestep, clip_text_embedding=clip_text_embedding, **kwargs)
        latents = torch.cat((x, x)) if cfg else x
        latents = self.solver.scale_model_input(latents, step=step)
        if cfg:
            unconditional, conditional = self._run_unet(latents).chunk(2)
            predicted_noise = unconditional + condition_scale * (conditional - unconditional)
            x = x.narrow(dim=1, start=0, length=4)  # > 4 input channels (inpainting) keep 4 latent ones
            if self.has_self_attention_guidance():
                predicted_noise += self.compute_self_attention_guidance(
                    x=x, noise=unconditional, step=step, clip_text_embedding=clip_text_embedding, **kwargs
                )
        else:
            predicted_noise = self._run_unet(latents)
            x = x.narrow(dim=1, start=0, length=4)
        return self.solver(x, predicted_noise=predicted_noise, step=step)

    def structural_copy(self: TModel) -> TModel:
        return self.__class__(  # type: ignore[call-arg]
            unet=self.unet.structural_copy(),
            lda=None if self.lda is None else self.lda.structural_copy(),
            clip_text_encoder=None if self.clip_text_encoder is None else self.clip_text_encoder.structural_copy(),
            solver=self.solver,
            device=self.device,
            dtype=self.dtype,
        )
