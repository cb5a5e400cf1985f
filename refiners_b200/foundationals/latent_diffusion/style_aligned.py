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


def reference_of(batch_size: int, device: torch.device | None = None) -> Tensor:
    """Index of the image every image of a guidance batch borrows its style from: the first of its own half.
    (half, half) images -> [0] * half + [half] * half."""
    half = batch_size // 2
    return torch.arange(batch_size, device=device).div(half, rounding_mode="floor") * half


def token_statistics(x: Tensor) -> tuple[Tensor, Tensor]:
    """Per (image, channel) mean and unbiased deviation over the tokens of [B, S, C], both [B, 1, C]."""
    return x.mean(dim=-2, keepdim=True), x.std(dim=-2, keepdim=True)


class ExtractReferenceFeatures(fl.Module):
    """[2b, S, C] -> [2b, S, C]: every image replaced by the reference image of its half."""

    def forward(self, features: Tensor) -> Tensor:
        return features.index_select(0, reference_of(features.shape[0], features.device))


class AdaIN(fl.Module):
    """(targets, reference) -> (targets standardised over their tokens, then given the reference's deviation and mean; reference)."""

    def __init__(self, epsilon: float = 1e-8) -> None:
        super().__init__()
        self.epsilon = epsilon

    def forward(self, targets: Tensor, reference: Tensor) -> tuple[Tensor, Tensor]:
        mean, deviation = token_statistics(targets)
        reference_mean, reference_deviation = token_statistics(reference)
        standardised = (targets - mean) / (deviation + self.epsilon)
        return standardised * reference_deviation + reference_mean, reference


class ScaleReferenceFeatures(fl.Module):
    """Multiply by ``scale`` every image that is not its own reference (the first of each half keeps factor 1)."""

    def __init__(self, scale: float = 1.0) -> None:
        super().__init__()
        self.scale = scale

    def forward(self, features: Tensor) -> Tensor:
        batch = features.shape[0]
        own = reference_of(batch, features.device) == torch.arange(batch, device=features.device)
        keep = own.reshape(batch, *([1] * (features.ndim - 1)))
        return torch.where(keep, features, features * self.scale)


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
