"""Self-attention guidance (SAG, arXiv:2210.00939): a second, unconditional UNet pass on latents that were blurred where the
middle block's self-attention concentrates, whose prediction is pushed away from.

Class names, constructor arguments, the context name / keys and the resulting trees are the contract of
/root/reference/src/refiners/foundationals/latent_diffusion/self_attention_guidance.py (`SelfAttentionMap` :22-47,
`SelfAttentionShape` :50-59, `SAGAdapter` :62-101).  The per-model adapters (stable_diffusion_1/self_attention_guidance.py,
stable_diffusion_xl/self_attention_guidance.py) decide WHERE the two probes go; `LatentDiffusionModel` runs the extra pass.

What the probes leave in the ``self_attention_map`` context on every UNet pass:

  middle_block_attn_shape   (height, width) of the middle block's feature map - appended; the mask pops it
  middle_block_attn_map     softmax(q k^T / sqrt(d)) of the middle block's FIRST self-attention, [B, heads, S, S]

On CUDA the map comes from ``rb200_attention_probs`` (csrc/attn_probs.cu) - the flash kernels never materialise it - while the
attention itself stays on the flash path: the probe sits between the q / k / v ``Distribute`` and the
``ScaledDotProductAttention``, where the fusion planner meets a module it does not know and runs that one layer leaf by leaf.
Mask, blur and re-noising are O(latents) glue on a handful of [B, 4, H, W] tensors.
"""

from __future__ import annotations

from typing import TYPE_CHECKING, Any, Generic, TypeVar

import torch
from torch import Size, Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.utils import gaussian_blur, interpolate
from refiners_b200.foundationals.latent_diffusion.solvers import Solver

if TYPE_CHECKING:
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet

T = TypeVar("T", bound="SD1UNet | SDXLUNet")
TSAGAdapter = TypeVar("TSAGAdapter", bound="SAGAdapter[Any]")

CONTEXT = "self_attention_map"
MAP, SHAPE = "middle_block_attn_map", "middle_block_attn_shape"


def attention_probabilities(query: Tensor, key: Tensor, num_heads: int) -> Tensor:
    """``softmax(q k^T / sqrt(d))`` per head: [B, S, heads * d] x 2 -> [B, heads, Sq, Sk]."""
    if query.is_cuda:
        return B.attention_probs(query, key, num_heads)
    assert query.ndim == 3, f"Expected tensor with shape (batch_size sequence_length embedding_dim), got {query.shape}"
    width = query.shape[-1]
    assert width % num_heads == 0, f"Embedding dim (x.shape[-1]={width}) must be divisible by num heads"
    per_head = [t.unflatten(-1, (num_heads, width // num_heads)).transpose(1, 2) for t in (query, key)]
    logits = per_head[0] @ per_head[1].transpose(-1, -2)
    return torch.softmax(logits / (width // num_heads) ** 0.5, dim=-1)


def attended_positions(probabilities: Tensor) -> Tensor:
    """[B, heads, S, S] -> bool [B, S]: key positions whose attention mass, averaged over the heads and summed over the
    queries, exceeds 1 (= more than their share: the masses of a row sum to 1 and there are as many rows as keys)."""
    return probabilities.mean(dim=1).sum(dim=1) > 1.0


class SelfAttentionMap(fl.Passthrough):
    """Sees (query, key, value) on their way to the attention and stores the attention probabilities."""

    def __init__(self, num_heads: int, context_key: str) -> None:
        self.num_heads = num_heads
        self.context_key = context_key
        super().__init__(fl.Lambda(func=self.compute_attention_scores), fl.SetContext(context=CONTEXT, key=context_key))

    def split_to_multi_head(self, x: Tensor) -> Tensor:
        """[B, S, heads * d] -> [B, heads, S, d] (a view)."""
        return x.unflatten(-1, (self.num_heads, x.shape[-1] // self.num_heads)).transpose(1, 2)

    def compute_attention_scores(self, query: Tensor, key: Tensor, value: Tensor) -> Tensor:
        return attention_probabilities(query, key, self.num_heads)


class SelfAttentionShape(fl.Passthrough):
    """Sees a feature map and appends its (height, width) to the context list."""

    def __init__(self, context_key: str) -> None:
        self.context_key = context_key
        super().__init__(fl.SetContext(context=CONTEXT, key=context_key, callback=self.register_shape))

    def register_shape(self, shapes: list[Size], x: Tensor) -> None:
        assert x.ndim == 4, f"Expected 4D tensor, got {x.ndim}D with shape {x.shape}"
        shapes.append(x.shape[-2:])


class SAGAdapter(Generic[T], fl.Chain, Adapter[T]):
    def __init__(self, target: T, scale: float = 1.0, kernel_size: int = 9, sigma: float = 1.0) -> None:
        self.scale = scale
        self.kernel_size = kernel_size
        self.sigma = sigma
        with self.setup_adapter(target):
            super().__init__(target)

    def inject(self: TSAGAdapter, parent: fl.Chain | None = None) -> TSAGAdapter:
        return super().inject(parent)

    def eject(self) -> None:
        super().eject()

    def init_context(self) -> Contexts:
        return {CONTEXT: {MAP: None, SHAPE: []}}

    def compute_sag_mask(self, latents: Tensor, classifier_free_guidance: bool = True) -> Tensor:
        """1.0 where a latent pixel is attended to more than its share (at the middle block's resolution, nearest-upsampled
        to the latents), on every channel; of a guided pass only the unconditional half of the batch counts."""
        recorded = self.use_context(CONTEXT)
        probabilities, map_size = recorded[MAP], recorded[SHAPE].pop()
        assert len(map_size) == 2
        if classifier_free_guidance:
            probabilities = probabilities[: probabilities.shape[0] // 2]
        batch, channels = latents.shape[:2]
        coarse = attended_positions(probabilities).to(probabilities.dtype).reshape(batch, 1, *map_size)
        return interpolate(coarse.repeat(1, channels, 1, 1), Size(latents.shape[-2:]))

    def compute_degraded_latents(
        self, solver: Solver, latents: Tensor, noise: Tensor, step: int, classifier_free_guidance: bool = True
    ) -> Tensor:
        """The clean latents the noise prediction implies, blurred inside the mask, noised back to ``step`` with that noise."""
        mask = self.compute_sag_mask(latents=latents, classifier_free_guidance=classifier_free_guidance)
        clean = solver.remove_noise(x=latents, noise=noise, step=step)
        blurred = gaussian_blur(clean, kernel_size=self.kernel_size, sigma=self.sigma)
        return solver.add_noise(blurred * mask + clean * (1 - mask), noise=noise, step=step)
