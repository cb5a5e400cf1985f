"""PerceiverResampler: the image projection of IP-Adapter "plus" (fine-grained prompts).

A small set of learnable latent tokens cross-attends to the CLIP patch features and comes out as ``num_tokens``
text-space tokens per image (16 for SDXL).  Module tree / state-dict keys follow the reference's
``foundationals/latent_diffusion/image_prompt.py`` (`FeedForward` :48-75, `PerceiverScaledDotProductAttention` :81-114,
`PerceiverAttention` :117-163, `LatentsToken` :166-172, `PerceiverResampler` :183-234).

It runs once per image prompt, not per denoising step; its output is the step-invariant
``ip_adapter.clip_image_embedding`` context that the per-step kernels consume.
"""

from __future__ import annotations

import math

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.context import Contexts

Device = torch.device
DType = torch.dtype
_CTX = "perceiver_resampler"


class FeedForward(fl.Chain):
    """Linear -> GeLU -> Linear, no biases."""

    def __init__(
        self, embedding_dim: int, feedforward_dim: int, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.feedforward_dim = embedding_dim, feedforward_dim
        kw = dict(bias=False, device=device, dtype=dtype)
        super().__init__(fl.Linear(embedding_dim, feedforward_dim, **kw), fl.GeLU(), fl.Linear(feedforward_dim, embedding_dim, **kw))


class PerceiverScaledDotProductAttention(fl.Module):
    """Attention of the latent queries over [features | latents]; q and k are each scaled by d^-1/4 BEFORE the
    product (more stable in fp16 than scaling the logits) and the softmax runs in fp32."""

    def __init__(self, head_dim: int, num_heads: int) -> None:
        super().__init__()
        self.num_heads = num_heads
        self.scale = 1 / math.sqrt(math.sqrt(head_dim))

    def forward(self, key_value: Tensor, query: Tensor) -> Tensor:
        batch, length, _ = query.shape
        key, value = key_value.chunk(2, dim=-1)
        if query.is_cuda:  # (q d^-1/4) (k d^-1/4)^T = q k^T d^-1/2: the library's attention (fp32 softmax inside), strided k / v read in place
            return B.sdpa(query, key, value, self.num_heads)
        q, k, v = self.reshape_tensor(query), self.reshape_tensor(key), self.reshape_tensor(value)
        logits = (q * self.scale) @ (k * self.scale).transpose(-2, -1)
        weights = torch.softmax(input=logits.float(), dim=-1).type(logits.dtype)
        return (weights @ v).permute(0, 2, 1, 3).reshape(batch, length, -1)

    def reshape_tensor(self, x: Tensor) -> Tensor:
        batch, length, _ = x.shape
        return x.view(batch, length, self.num_heads, -1).transpose(1, 2).reshape(batch, self.num_heads, length, -1)


class PerceiverAttention(fl.Chain):
    """(features, latents) -> latents': both normed, keys/values from their concatenation, queries from the latents."""

    def __init__(
        self, embedding_dim: int, head_dim: int = 64, num_heads: int = 8, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.head_dim = embedding_dim, head_dim
        self.inner_dim = head_dim * num_heads
        kw = dict(device=device, dtype=dtype)

        def project(width: int) -> fl.Linear:
            return fl.Linear(embedding_dim, width, bias=False, **kw)

        super().__init__(
            fl.Distribute(fl.LayerNorm(embedding_dim, **kw), fl.LayerNorm(embedding_dim, **kw)),
            fl.Parallel(
                fl.Chain(fl.Lambda(func=self.to_kv), project(2 * self.inner_dim)),
                fl.Chain(fl.GetArg(index=1), project(self.inner_dim)),
            ),
            PerceiverScaledDotProductAttention(head_dim=head_dim, num_heads=num_heads),
            fl.Linear(self.inner_dim, embedding_dim, bias=False, **kw),
        )

    def to_kv(self, x: Tensor, latents: Tensor) -> Tensor:
        return torch.cat((x, latents), dim=-2)


class LatentsToken(fl.Chain):
    """The learnable latent queries, broadcast over the batch."""

    def __init__(self, num_tokens: int, latents_dim: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.num_tokens, self.latents_dim = num_tokens, latents_dim
        super().__init__(fl.Parameter(num_tokens, latents_dim, device=device, dtype=dtype))


class Transformer(fl.Chain):
    pass


class TransformerLayer(fl.Chain):
    pass


class PerceiverResampler(fl.Chain):
    """CLIP patch features ``[B, S, input_dim]`` -> ``[B, num_tokens, output_dim]``."""

    def __init__(
        self,
        latents_dim: int = 1024,
        num_attention_layers: int = 8,
        num_attention_heads: int = 16,
        head_dim: int = 64,
        num_tokens: int = 8,
        input_dim: int = 768,
        output_dim: int = 1024,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.latents_dim, self.num_attention_layers, self.head_dim = latents_dim, num_attention_layers, head_dim
        self.num_attention_heads, self.num_tokens = num_attention_heads, num_tokens
        self.input_dim, self.output_dim = input_dim, output_dim
        self.feedforward_dim = 4 * latents_dim
        kw = dict(device=device, dtype=dtype)

        def layer() -> TransformerLayer:
            return TransformerLayer(
                fl.Residual(
                    fl.Parallel(fl.UseContext(context=_CTX, key="x"), fl.Identity()),  # (features, latents)
                    PerceiverAttention(embedding_dim=latents_dim, head_dim=head_dim, num_heads=num_attention_heads, **kw),
                ),
                fl.Residual(fl.LayerNorm(latents_dim, **kw), FeedForward(latents_dim, self.feedforward_dim, **kw)),
            )

        super().__init__(
            fl.Linear(input_dim, latents_dim, **kw),
            fl.SetContext(context=_CTX, key="x"),
            LatentsToken(num_tokens, latents_dim, **kw),
            Transformer(layer() for _ in range(num_attention_layers)),
            fl.Linear(latents_dim, output_dim, **kw),
            fl.LayerNorm(output_dim, **kw),
        )

    def init_context(self) -> Contexts:
        return {_CTX: {"x": None}}
