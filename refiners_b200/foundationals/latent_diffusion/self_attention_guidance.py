"""Self-attention guidance (SAG, arXiv:2210.00939): a second, unconditional UNet pass on latents that were blurred
where the middle block's self-attention concentrates, whose prediction is pushed away from.

Contract (class names, constructor arguments, context name / keys, resulting tree) from
/root/reference/src/refiners/foundationals/latent_diffusion/self_attention_guidance.py:
`SelfAttentionMap` :22-47, `SelfAttentionShape` :50-59, `SAGAdapter` :62-101.  The per-model adapters
(stable_diffusion_1/self_attention_guidance.py, stable_diffusion_xl/self_attention_guidance.py) decide WHERE the two
probes go; the models' ``compute_self_attention_guidance`` runs the extra pass (model.py:147-154 of the reference).

What the probes record, per UNet pass, in the ``self_attention_map`` context:

  middle_block_attn_shape   spatial size of the middle block's feature map (appended; popped by the mask)
  middle_block_attn_map     softmax(q k^T / sqrt(d)) of the middle block's FIRST self-attention, [B, heads, S, S]

On CUDA the map comes from ``rb200_attention_probs`` (csrc/attn_probs.cu) - the flash kernels never materialise it -
and the attention itself still runs on the flash path: the probe sits between the q/k/v ``Distribute`` and the
``ScaledDotProductAttention``, where the fusion planner sees an unknown module and falls back to per-leaf launches
for that one layer.  The mask, blur and re-noising are O(latents) glue on a handful of [B, 4, H, W] tensors.
"""

from __future__ import annotations

import math
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


class SelfAttentionMap(fl.Passthrough):
    """Sees (query, key, value) on their way to the attention and stores the attention probabilities."""

    def __init__(self, num_heads: int, context_key: str) -> None:
        self.num_heads = num_heads
        self.context_key = context_key
        super().__init__(
            fl.Lambda(func=self.compute_attention_scores),
            fl.SetContext(context=CONTEXT, key=context_key),
        )

    def split_to_multi_head(self, x: Tensor) -> Tensor:
        """[B, S, heads * d] -> [B, heads, S, d]."""
        assert x.ndim == 3, f"Expected tensor with shape (batch_size sequence_length embedding_dim), got {x.shape}"
        batch, length, width = x.shape
        assert width % self.num_heads == 0, f"Embedding dim (x.shape[-1]={width}) must be divisible by num heads"
        return x.reshape(batch, length, self.num_heads, width // self.num_heads).transpose(1, 2)

    def compute_attention_scores(self, query: Tensor, key: Tensor, value: Tensor) -> Tensor:
        if query.is_cuda:
            return B.attention_probs(query, key, self.num_heads)
        heads_q, heads_k = self.split_to_multi_head(query), self.split_to_multi_head(key)
        logits = heads_q @ heads_k.permute(0, 1, 3, 2)
        return torch.softmax(logits / math.sqrt(heads_q.shape[-1]), dim=-1)


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
        return {CONTEXT: {"middle_block_attn_map": None, "middle_block_attn_shape": []}}

    def compute_sag_mask(self, latents: Tensor, classifier_free_guidance: bool = True) -> Tensor:
        """1.0 where a latent pixel is attended to more than average: key positions whose attention mass, averaged
        over heads and summed over queries, exceeds 1 - at the middle block's resolution, nearest-upsampled."""
        recorded = self.use_context(CONTEXT)
        probabilities = recorded["middle_block_attn_map"]
        if classifier_free_guidance:
            probabilities = probabilities.chunk(2)[0]  # the unconditional half
        map_size = recorded["middle_block_attn_shape"].pop()
        assert len(map_size) == 2
        batch, channels, height, width = latents.shape
        attended = probabilities.mean(dim=1).sum(dim=1) > 1.0
        mask = attended.reshape(batch, *map_size).unsqueeze(1).repeat(1, channels, 1, 1).type(probabilities.dtype)
        return interpolate(mask, Size((height, width)))

    def compute_degraded_latents(
        self, solver: Solver, latents: Tensor, noise: Tensor, step: int, classifier_free_guidance: bool = True
    ) -> Tensor:
        """Predicted clean latents, blurred inside the mask, noised back to ``step`` with the same noise."""
        mask = self.compute_sag_mask(latents=latents, classifier_free_guidance=classifier_free_guidance)
        clean = solver.remove_noise(x=latents, noise=noise, step=step)
        blurred = gaussian_blur(clean, kernel_size=self.kernel_size, sigma=self.sigma)
        return solver.add_noise(blurred * mask + clean * (1 - mask), noise=noise, step=step)
