"""StyleAligned (arXiv:2312.02133): every image of a batch attends, in every self-attention, to the FIRST image of the batch as
well as to itself, after its queries and keys were moved to the first image's per-channel statistics - so a batch of prompts
comes out in one shared style.

Contract (class names, constructor arguments, resulting trees) from
/root/reference/src/refiners/foundationals/latent_diffusion/style_aligned.py: `ExtractReferenceFeatures` :13-44, `AdaIN` :47-92,
`ScaleReferenceFeatures` :95-133, `StyleAligned` :136-207, `SharedSelfAttentionAdapter` :210-265, `StyleAlignedAdapter` :268-329.

The batch is the classifier-free-guidance batch: two halves of ``batch_size`` images each, and "the first image" means the
first of EACH half.  Between the q / k / v projections and the attention of every `SelfAttention`, three `StyleAligned` chains
turn (q, k, v) into

    q' = adain(q, q_ref)                                       [B, S, C]
    k' = cat(adain(k, k_ref), scaled(k_ref))  along the tokens   [B, 2S, C]
    v' = cat(v, scaled(v_ref))                                  [B, 2S, C]

with ``x_ref`` the first image of x's half repeated over the half, ``adain`` = per-(image, channel) standardisation over the
tokens followed by the reference's deviation and mean, and ``scaled`` multiplying the reference by ``scale`` for every image
but the first of its half (which therefore sees itself twice, unscaled).

On CUDA each `StyleAligned` chain is TWO launches whatever its variant (`rb200_style_aligned`: per-(image, channel) token
statistics, then one pass that writes the renormalised targets and the scaled reference rows); the attention then runs on
the flash kernel with twice as many keys as queries.
"""

from __future__ import annotations

from functools import cached_property
from typing import Any, Generic, TypeVar

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet

T = TypeVar("T", bound="SD1UNet | SDXLUNet")


class ExtractReferenceFeatures(fl.Module):
    """[2b, S, C] -> [2b, S, C]: image 0 of each guidance half, repeated over its half."""

    def forward(self, features: Tensor) -> Tensor:
        half = features.shape[0] // 2
        first, second = features.chunk(2, dim=0)
        return torch.stack((first[0], second[0])).repeat_interleave(half, dim=0)


class AdaIN(fl.Module):
    """(targets, reference) -> (targets moved to the reference's per-channel token statistics, reference)."""

    def __init__(self, epsilon: float = 1e-8) -> None:
        super().__init__()
        self.epsilon = epsilon

    def forward(self, targets: Tensor, reference: Tensor) -> tuple[Tensor, Tensor]:
        def stats(x: Tensor) -> tuple[Tensor, Tensor]:
            return torch.mean(x, dim=-2, keepdim=True), torch.std(x, dim=-2, keepdim=True)

        centre, spread = stats(targets)
        ref_centre, ref_spread = stats(reference)
        return (targets - centre) / (spread + self.epsilon) * ref_spread + ref_centre, reference


class ScaleReferenceFeatures(fl.Module):
    """Multiply by ``scale`` every image but the first of each guidance half."""

    def __init__(self, scale: float = 1.0) -> None:
        super().__init__()
        self.scale = scale

    def forward(self, features: Tensor) -> Tensor:
        half = features.shape[0] // 2
        scaled = features.clone()
        scaled.reshape(2, half, *features.shape[1:])[:, 1:] *= self.scale
        return scaled


class StyleAligned(fl.Chain):
    def __init__(self, adain: bool, concatenate: bool, scale: float = 1.0) -> None:
        super().__init__(
            fl.Parallel(fl.Identity(), ExtractReferenceFeatures()),
            AdaIN(),
            fl.Distribute(fl.Identity(), ScaleReferenceFeatures(scale=scale)),
            fl.Concatenate(fl.GetArg(index=0), fl.GetArg(index=1), dim=-2),
        )
        if not adain:
            self.remove(self.ensure_find(AdaIN))
        if not concatenate:
            self.replace(old_module=self.ensure_find(fl.Concatenate), new_module=fl.GetArg(index=0))

    @property
    def scale(self) -> float:
        return self.ensure_find(ScaleReferenceFeatures).scale

    @scale.setter
    def scale(self, scale: float) -> None:
        self.ensure_find(ScaleReferenceFeatures).scale = scale

    def forward(self, *inputs: Any) -> Any:
        x = inputs[0]
        if len(inputs) == 1 and isinstance(x, Tensor) and x.is_cuda and x.ndim == 3 and x.shape[0] % 2 == 0 and self._stock():
            normalise = self.find(AdaIN)
            return B.style_aligned(
                x, adain=normalise is not None, concatenate=self.find(fl.Concatenate) is not None, scale=float(self.scale),
                epsilon=float(normalise.epsilon) if normalise is not None else 0.0,
            )
        return super().forward(*inputs)

    def _stock(self) -> bool:
        """Only the tree the constructor builds (and no hooks) takes the fused path; an edited chain runs module by module."""
        kinds = [type(m).__name__ for m in self]
        shape_ok = kinds in (["Parallel", "AdaIN", "Distribute", "Concatenate"], ["Parallel", "AdaIN", "Distribute", "GetArg"],
                             ["Parallel", "Distribute", "Concatenate"], ["Parallel", "Distribute", "GetArg"])
        return shape_ok and not any(m._forward_hooks or m._forward_pre_hooks for m in self.modules())


class SharedSelfAttentionAdapter(fl.Chain, Adapter[fl.SelfAttention]):
    """Puts a `Distribute` of three `StyleAligned` chains (for q, k, v) in front of the attention of one `SelfAttention`."""

    def __init__(self, target: fl.SelfAttention, scale: float = 1.0) -> None:
        with self.setup_adapter(target):
            super().__init__(target)
        self._style_aligned_layers = [
            StyleAligned(adain=True, concatenate=False, scale=scale),   # queries
            StyleAligned(adain=True, concatenate=True, scale=scale),    # keys
            StyleAligned(adain=False, concatenate=True, scale=scale),   # values
        ]

    @cached_property
    def style_aligned_layers(self) -> fl.Distribute:
        return fl.Distribute(*self._style_aligned_layers)

    def inject(self, parent: fl.Chain | None = None) -> "SharedSelfAttentionAdapter":
        self.target.insert_before_type(module_type=fl.ScaledDotProductAttention, new_module=self.style_aligned_layers)
        return super().inject(parent)

    def eject(self) -> None:
        self.target.remove(self.style_aligned_layers)
        super().eject()

    @property
    def scale(self) -> float:
        return self.style_aligned_layers.layer(0, StyleAligned).scale

    @scale.setter
    def scale(self, scale: float) -> None:
        for chain in self.style_aligned_layers:
            chain.scale = scale


class StyleAlignedAdapter(Generic[T], fl.Chain, Adapter[T]):
    """`SharedSelfAttentionAdapter`s on every `SelfAttention` of a UNet."""

    def __init__(self, target: T, scale: float = 1.0) -> None:
        with self.setup_adapter(target):
            super().__init__(target)
        self.shared_self_attention_adapters = tuple(
            SharedSelfAttentionAdapter(target=attention, scale=scale) for attention in self.target.layers(fl.SelfAttention)
        )

    def inject(self, parent: fl.Chain | None = None) -> "StyleAlignedAdapter[T]":
        for adapter in self.shared_self_attention_adapters:
            adapter.inject()
        return super().inject(parent)

    def eject(self) -> None:
        for adapter in self.shared_self_attention_adapters:
            adapter.eject()
        super().eject()

    @property
    def scale(self) -> float:
        return self.shared_self_attention_adapters[0].scale

    @scale.setter
    def scale(self, scale: float) -> None:
        for adapter in self.shared_self_attention_adapters:
            adapter.scale = scale
