"""One denoising step of a latent diffusion model: guidance batching, UNet call, guidance combine, solver update.

Contract (constructor, ``forward`` signature and semantics, ``init_latents`` / ``sample_noise`` / ``steps`` /
``set_inference_steps`` / ``structural_copy``) from
/root/reference/src/refiners/foundationals/latent_diffusion/model.py:15-169 (``forward`` :128-159).

The step is written here as three stages - *prepare* the model input, *predict* the noise, *advance* the latents - so
that the CUDA path can swap each stage for its fused form without changing what is computed:

  prepare   ``scale_model_input(cat((x, x)))``                    one launch (rb200_cfg_scale_input) for the Euler solver
  predict   the UNet, replayed from a captured CUDA graph when ``enable_cuda_graph()`` is on; text / image-prompt K, V
            projections and condition encoders are hoisted out of the per-step graph (engine.graph)
  advance   ``uncond + s * (cond - uncond)`` then the solver update    one launch (rb200_cfg_euler) for the Euler solver

The fused stages round every intermediate exactly where the operator-by-operator evaluation does, so both give the
same bits (tests/test_full_size_gpu.py compares them).  The VAE (``lda``) and the text encoder run once per image /
prompt, outside the per-step hot path; they are optional here and any fluxion Chain works.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, TypeVar

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.foundationals.latent_diffusion.solvers import Euler, ModelPredictionType, Solver

TModel = TypeVar("TModel", bound="LatentDiffusionModel")
Device = torch.device
DType = torch.dtype


class LatentDiffusionModel(fl.Module, ABC):
    def __init__(
        self, unet: fl.Chain, lda: fl.Chain | None, clip_text_encoder: fl.Chain | None, solver: Solver,
        classifier_free_guidance: bool = True, device: Device | str = "cpu", dtype: DType = torch.float32,
    ) -> None:
        super().__init__()
        self.device: Device = Device(device) if isinstance(device, str) else device
        self.dtype = dtype
        place = lambda part: None if part is None else part.to(device=self.device, dtype=self.dtype)  # noqa: E731
        self.unet = place(unet)
        self.lda = place(lda)
        self.clip_text_encoder = place(clip_text_encoder)
        self.solver = place(solver)
        self.classifier_free_guidance = classifier_free_guidance
        self._graphed_unet: list[Any] = []  # at most one engine.graph.GraphedChain; list-wrapped: not a torch sub-module

    # -- execution engine ---------------------------------------------------------------------------
    def enable_cuda_graph(self, enabled: bool = True) -> None:
        """Replay the UNet forward from a captured CUDA graph (refiners_b200.engine.graph.GraphedChain).  The first call
        after enabling - or after any edit of the tree, a scale or a weight - runs the Python walker once under capture;
        contexts keep being set through the normal API."""
        while self._graphed_unet:
            self._graphed_unet.pop().close()
        if enabled:
            from refiners_b200.engine.graph import GraphedChain

            self._graphed_unet.append(GraphedChain(self.unet))

    def _run_unet(self, latents: Tensor) -> Tensor:
        # self-attention guidance reads what the walker's probes record on every pass: it runs the walker itself
        graphed = self._graphed_unet and latents.is_cuda and not self.has_self_attention_guidance()
        runner = self._graphed_unet[0] if graphed else self.unet
        return runner(latents)

    def _fused_euler(self, x: Tensor) -> bool:
        """Whether the prepare / advance stages may run as the two fused launches: plain Euler in noise-prediction mode
        on CUDA latents of the solver's dtype with the four latent channels only, fusion switched on, no SAG."""
        solver = self.solver
        return (
            type(solver) is Euler
            and solver.params.model_prediction_type == ModelPredictionType.NOISE
            and x.is_cuda and x.dtype == solver.sigmas.dtype and solver.sigmas.device == x.device
            and x.ndim == 4 and x.shape[1] == 4
            and B.fusion_enabled()
            and not self.has_self_attention_guidance()
        )

    # -- the denoising step -----------------------------------------------------------------------------
    @abstractmethod
    def set_unet_context(self, *, timestep: Tensor, clip_text_embedding: Tensor, **_: Tensor) -> None: ...

    # -- self-attention guidance (arXiv:2210.00939) ----------------------------------------------------------
    def _sag_adapter_type(self) -> type[Any] | None:
        """The SAG adapter class of this model family (None: the family has none)."""
        return None

    def _find_sag_adapter(self) -> Any:
        kind = self._sag_adapter_type()
        if kind is None:
            return None
        return next((parent for parent in self.unet.get_parents() if isinstance(parent, kind)), None)

    def has_self_attention_guidance(self) -> bool:
        return self._find_sag_adapter() is not None

    def set_self_attention_guidance(self, enable: bool, scale: float = 1.0) -> None:
        """Inject (or re-scale) / eject the family's SAG adapter around the UNet."""
        adapter = self._find_sag_adapter()
        if not enable:
            if adapter is not None:
                adapter.eject()
        elif adapter is not None:
            adapter.scale = scale
        else:
            kind = self._sag_adapter_type()
            assert kind is not None, f"{type(self).__name__} has no self-attention guidance adapter"
            kind(target=self.unet, scale=scale).inject()

    def _unconditional_half(self, *, clip_text_embedding: Tensor, **kwargs: Tensor) -> dict[str, Tensor]:
        """The conditioning of the guidance pass: the unconditional half of whatever is batched (uncond | cond)."""
        return {"clip_text_embedding": clip_text_embedding.chunk(2)[0], **kwargs}

    def compute_self_attention_guidance(
        self, x: Tensor, noise: Tensor, step: int, *, clip_text_embedding: Tensor, **kwargs: Tensor,
    ) -> Tensor:
        """``scale * (noise - unet(degraded latents))`` where the degraded latents are ``x`` blurred where the middle
        block attends (the probes recorded that during the pass that produced ``noise``), and the extra UNet pass is
        unconditional: half the batch, half of every batched context (reference stable_diffusion_1/model.py:175-213,
        stable_diffusion_xl/model.py:194-250)."""
        adapter = self._find_sag_adapter()
        assert adapter is not None
        degraded = adapter.compute_degraded_latents(
            solver=self.solver, latents=x, noise=noise, step=step, classifier_free_guidance=True
        )
        self.set_unet_context(
            timestep=self.solver.timesteps[step].unsqueeze(dim=0),
            **self._unconditional_half(clip_text_embedding=clip_text_embedding, **kwargs),
        )
        if "ip_adapter" not in self.unet.provider.contexts:
            return adapter.scale * (noise - self.unet(degraded))
        # an injected IP-Adapter keeps its (uncond | cond) image embedding in a context of its own: halve it for this
        # pass and put the full one back
        image_prompt = self.unet.use_context("ip_adapter")
        full = image_prompt["clip_image_embedding"]
        image_prompt["clip_image_embedding"] = full.chunk(2)[0]
        try:
            return adapter.scale * (noise - self.unet(degraded))
        finally:
            image_prompt["clip_image_embedding"] = full

    def forward(
        self, x: Tensor, step: int, *, clip_text_embedding: Tensor, condition_scale: float = 7.5, **kwargs: Tensor,
    ) -> Tensor:
        guided = self.classifier_free_guidance
        if guided:
            assert clip_text_embedding.shape[0] % 2 == 0, f"invalid batch size: {clip_text_embedding.shape[0]}"
        self.set_unet_context(
            timestep=self.solver.timesteps[step].unsqueeze(dim=0), clip_text_embedding=clip_text_embedding, **kwargs
        )
        if self._fused_euler(x):
            model_input = B.cfg_scale_input(x, self.solver.sigmas, step, twice=guided)
            return B.cfg_euler(x, self._run_unet(model_input), self.solver.sigmas, step, condition_scale, guided)

        # prepare: with guidance the batch is (unconditional | conditional) copies of the same latents
        model_input = self.solver.scale_model_input(torch.cat((x, x)) if guided else x, step=step)
        prediction = self._run_unet(model_input)
        # advance: only the four latent channels move (an inpainting UNet is fed more)
        latents = x.narrow(dim=1, start=0, length=4)
        if guided:
            unconditional, conditional = prediction.chunk(2)
            prediction = unconditional + condition_scale * (conditional - unconditional)
            if self.has_self_attention_guidance():
                prediction += self.compute_self_attention_guidance(
                    x=latents, noise=unconditional, step=step, clip_text_embedding=clip_text_embedding, **kwargs
                )
        return self.solver(latents, predicted_noise=prediction, step=step)

    # -- prompts ------------------------------------------------------------------------------------------
    def compute_clip_text_embedding(self, text: str | list[str], negative_text: str | list[str] = "") -> Any:
        """What ``clip_text_encoder`` makes of the prompt(s); with classifier-free guidance the negative prompt's result
        comes first in the batch (unconditional | conditional).  A tower that returns several tensors (SDXL: token and
        pooled embeddings) has each of them batched that way.  Reference: stable_diffusion_1/model.py:114-133,
        stable_diffusion_xl/model.py:87-111."""
        assert self.clip_text_encoder is not None, "this model was built without a text encoder"
        prompts = [text] if isinstance(text, str) else text
        if not self.classifier_free_guidance:
            return self.clip_text_encoder(prompts)
        negatives = [negative_text] if isinstance(negative_text, str) else negative_text
        assert len(prompts) == len(negatives), "The length of the text list and negative_text should be the same"
        wanted, unwanted = self.clip_text_encoder(prompts), self.clip_text_encoder(negatives)
        if isinstance(wanted, tuple):
            return tuple(torch.cat((no, yes), dim=0) for no, yes in zip(unwanted, wanted))
        return torch.cat((unwanted, wanted))

    # -- bookkeeping around the loop ------------------------------------------------------------------------
    @property
    def steps(self) -> list[int]:
        return self.solver.inference_steps

    def set_inference_steps(self, num_steps: int, first_step: int = 0) -> None:
        self.solver = self.solver.rebuild(num_inference_steps=num_steps, first_inference_step=first_step)

    @staticmethod
    def sample_noise(
        size: tuple[int, ...], device: Device | None = None, dtype: DType | None = None,
        offset_noise: float | None = None,
    ) -> Tensor:
        """Standard normal noise; ``offset_noise`` adds a per-(sample, channel) constant drawn after it."""
        noise = torch.randn(size=size, device=device, dtype=dtype)
        if offset_noise is None:
            return noise
        per_channel = torch.randn(size=(size[0], size[1], 1, 1), device=device, dtype=dtype)
        return noise.add_(offset_noise * per_channel)

    def init_latents(self, size: tuple[int, int], init_image: Any = None, noise: Tensor | None = None) -> Tensor:
        """Starting latents for an image of ``size`` = (height, width) pixels: pure noise, or - image to image - the
        encoded ``init_image`` noised up to the solver's first inference step; scaled for the solver's first input."""
        height, width = size
        latent_hw = [height // 8, width // 8]
        if noise is None:
            noise = self.sample_noise(size=(1, 4, *latent_hw), device=self.device, dtype=self.dtype)
        assert list(noise.shape[2:]) == latent_hw, f"noise shape is not compatible: {noise.shape}, with size: {size}"
        start = noise
        if init_image is not None:
            assert self.lda is not None, "image-to-image needs a latent autoencoder (out of the hot path)"
            encoded = self.lda.image_to_latents(init_image.resize(size=(width, height)))  # type: ignore[attr-defined]
            start = self.solver.add_noise(x=encoded, noise=noise, step=self.solver.first_inference_step)
        return self.solver.scale_model_input(start, step=-1)

    def structural_copy(self: TModel) -> TModel:
        twin = lambda part: None if part is None else part.structural_copy()  # noqa: E731
        return type(self)(  # type: ignore[call-arg]
            unet=self.unet.structural_copy(), lda=twin(self.lda), clip_text_encoder=twin(self.clip_text_encoder),
            solver=self.solver, device=self.device, dtype=self.dtype,
        )
