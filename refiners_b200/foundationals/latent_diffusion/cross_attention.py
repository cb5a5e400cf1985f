"""Transformer blocks of the latent-diffusion UNets.

Tree shapes follow /root/reference/src/refiners/foundationals/latent_diffusion/cross_attention.py
(`CrossAttentionBlock` :25-73, `StatefulFlatten` :76-89, `CrossAttentionBlock2d` :92-175).
"""

from __future__ import annotations

import torch
from torch import Size, Tensor

from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.layers import (
    GLU,
    Attention,
    Chain,
    Conv2d,
    Flatten,
    GeLU,
    GroupNorm,
    Identity,
    LayerNorm,
    Linear,
    Parallel,
    Residual,
    SelfAttention,
    SetContext,
    Transpose,
    Unflatten,
    UseContext,
)

Device = torch.device
DType = torch.dtype


class CrossAttentionBlock(Chain):
    """self-attention, cross-attention on ``cross_attention_block.<context_key>``, GEGLU MLP -
    each pre-normed and residual."""

    def __init__(
        self,
        embedding_dim: int,
        context_embedding_dim: int,
        context_key: str,
        num_heads: int = 1,
        use_bias: bool = True,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.context_embedding_dim = context_embedding_dim
        self.context = "cross_attention_block"
        self.context_key = context_key
        self.num_heads = num_heads
        self.use_bias = use_bias
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            Residual(
                LayerNorm(embedding_dim, **kw),
                SelfAttention(embedding_dim=embedding_dim, num_heads=num_heads, use_bias=use_bias, **kw),
            ),
            Residual(
                LayerNorm(embedding_dim, **kw),
                Parallel(
                    Identity(),
                    UseContext(context=self.context, key=context_key),
                    UseContext(context=self.context, key=context_key),
                ),
                Attention(
                    embedding_dim=embedding_dim,
                    num_heads=num_heads,
                    key_embedding_dim=context_embedding_dim,
                    value_embedding_dim=context_embedding_dim,
                    use_bias=use_bias,
                    **kw,
                ),
            ),
            Residual(
                LayerNorm(embedding_dim, **kw),
                Linear(embedding_dim, 2 * 4 * embedding_dim, **kw),
                GLU(GeLU()),
                Linear(4 * embedding_dim, embedding_dim, **kw),
            ),
        )


class StatefulFlatten(Chain):
    """Flatten that first pushes the flattened sizes on a context stack (popped by the
    matching Unflatten)."""

    def __init__(self, context: str, key: str, start_dim: int = 0, end_dim: int = -1) -> None:
        self.start_dim = start_dim
        self.end_dim = end_dim
        super().__init__(
            SetContext(context=context, key=key, callback=self.push),
            Flatten(start_dim=start_dim, end_dim=end_dim),
        )

    def push(self, sizes: list[Size], x: Tensor) -> None:
        stop = self.end_dim + 1 if self.end_dim >= 0 else x.ndim + self.end_dim + 1
        sizes.append(x.shape[self.start_dim : stop])


class CrossAttentionBlock2d(Residual):
    """NCHW wrapper: GroupNorm, project in, N transformer blocks on [B, HW, C], project out."""

    def __init__(
        self,
        channels: int,
        context_embedding_dim: int,
        context_key: str,
        num_attention_heads: int = 1,
        num_attention_layers: int = 1,
        num_groups: int = 32,
        use_bias: bool = True,
        use_linear_projection: bool = False,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        assert channels % num_attention_heads == 0, "in_channels must be divisible by num_attention_heads"
        self.channels = channels
        self.in_channels = channels
        self.out_channels = channels
        self.context_embedding_dim = context_embedding_dim
        self.num_attention_heads = num_attention_heads
        self.num_attention_layers = num_attention_layers
        self.num_groups = num_groups
        self.use_bias = use_bias
        self.context_key = context_key
        self.use_linear_projection = use_linear_projection
        self.projection_type = "Linear" if use_linear_projection else "Conv2d"
        kw = dict(device=device, dtype=dtype)

        norm = GroupNorm(channels=channels, num_groups=num_groups, eps=1e-6, **kw)
        to_sequence = (StatefulFlatten(context="flatten", key="sizes", start_dim=2), Transpose(1, 2))

        def to_map() -> tuple[Transpose, Parallel, Unflatten]:
            return (
                Transpose(1, 2),
                Parallel(Identity(), UseContext(context="flatten", key="sizes").compose(lambda sizes: sizes.pop())),
                Unflatten(dim=2),
            )

        if use_linear_projection:
            in_block = Chain(norm, *to_sequence, Linear(channels, channels, **kw))
            out_block = Chain(Linear(channels, channels, **kw), *to_map())
        else:
            in_block = Chain(norm, Conv2d(channels, channels, kernel_size=1, **kw), *to_sequence)
            out_block = Chain(*to_map(), Conv2d(channels, channels, kernel_size=1, **kw))

        super().__init__(
            in_block,
            Chain(
                CrossAttentionBlock(
                    embedding_dim=channels,
                    context_embedding_dim=context_embedding_dim,
                    context_key=context_key,
                    num_heads=num_attention_heads,
                    use_bias=use_bias,
                    **kw,
                )
                for _ in range(num_attention_layers)
            ),
            out_block,
        )

    def init_context(self) -> Contexts:
        return {"flatten": {"sizes": []}}
