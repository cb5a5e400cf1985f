"""Transformer blocks of the latent-diffusion UNets.

The module trees (and therefore the state-dict keys) are those of the reference's
``foundationals/latent_diffusion/cross_attention.py`` (`CrossAttentionBlock` :25-73, `StatefulFlatten`
:76-89, `CrossAttentionBlock2d` :92-175); ``tests/test_reference_structure.py`` compares them node by
node.  What the engine makes of these trees on a GPU (fused q/k/v GEMM, residual and GEGLU epilogues,
tcgen05 attention) is decided by ``refiners_b200.engine.fusion`` from the leaf types, not here.
"""

from __future__ import annotations

import torch
from torch import Size, Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.context import Contexts

Device = torch.device
DType = torch.dtype

_CONTEXT = "cross_attention_block"


def _prenorm_residual(width: int, kw: dict, *body: fl.Module) -> fl.Residual:
    """x + body(LayerNorm(x)): the shape of all three sub-blocks of a transformer block."""
    return fl.Residual(fl.LayerNorm(width, **kw), *body)


class CrossAttentionBlock(fl.Chain):
    """One transformer block on ``[B, S, C]`` tokens: self-attention, attention over the conditioning
    tokens found at context ``cross_attention_block.<context_key>``, GEGLU MLP (4x expansion)."""

    def __init__(
        self, embedding_dim: int, context_embedding_dim: int, context_key: str, num_heads: int = 1, use_bias: bool = True,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.context_embedding_dim = embedding_dim, context_embedding_dim
        self.context, self.context_key = _CONTEXT, context_key
        self.num_heads, self.use_bias = num_heads, use_bias
        kw = dict(device=device, dtype=dtype)
        width, hidden = embedding_dim, 4 * embedding_dim

        def conditioning() -> fl.UseContext:
            return fl.UseContext(context=_CONTEXT, key=context_key)

        tokens_self = fl.SelfAttention(embedding_dim=width, num_heads=num_heads, use_bias=use_bias, **kw)
        tokens_text = fl.Attention(
            embedding_dim=width, num_heads=num_heads, key_embedding_dim=context_embedding_dim,
            value_embedding_dim=context_embedding_dim, use_bias=use_bias, **kw,
        )
        super().__init__(
            _prenorm_residual(width, kw, tokens_self),
            # query from the tokens, key and value from the conditioning sequence
            _prenorm_residual(width, kw, fl.Parallel(fl.Identity(), conditioning(), conditioning()), tokens_text),
            _prenorm_residual(width, kw, fl.Linear(width, 2 * hidden, **kw), fl.GLU(fl.GeLU()), fl.Linear(hidden, width, **kw)),
        )


class StatefulFlatten(fl.Chain):
    """Flatten that remembers: the sizes it collapses are pushed on the list at ``<context>.<key>``
    and popped again by the matching Unflatten on the way out."""

    def __init__(self, context: str, key: str, start_dim: int = 0, end_dim: int = -1) -> None:
        self.start_dim, self.end_dim = start_dim, end_dim
        super().__init__(
            fl.SetContext(context=context, key=key, callback=self.push),
            fl.Flatten(start_dim=start_dim, end_dim=end_dim),
        )

    def push(self, sizes: list[Size], x: Tensor) -> None:
        last = self.end_dim if self.end_dim >= 0 else x.ndim + self.end_dim
        sizes.append(x.shape[self.start_dim : last + 1])


class CrossAttentionBlock2d(fl.Residual):
    """The transformer of a UNet level: GroupNorm, channel projection in, ``num_attention_layers``
    `CrossAttentionBlock`s on the ``[B, H*W, C]`` pixel sequence, channel projection out, plus the skip.
    SD 1.5 projects with 1x1 convs on the map, SDXL with Linears on the sequence."""

    def __init__(
        self, channels: int, context_embedding_dim: int, context_key: str, num_attention_heads: int = 1,
        num_attention_layers: int = 1, num_groups: int = 32, use_bias: bool = True, use_linear_projection: bool = False,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        assert channels % num_attention_heads == 0, "in_channels must be divisible by num_attention_heads"
        # repr() echoes these in assignment order (the reference's order): keep it
        self.channels = self.in_channels = self.out_channels = channels
        self.context_embedding_dim = context_embedding_dim
        self.num_attention_heads, self.num_attention_layers = num_attention_heads, num_attention_layers
        self.num_groups, self.use_bias, self.context_key = num_groups, use_bias, context_key
        self.use_linear_projection = use_linear_projection
        self.projection_type = "Linear" if use_linear_projection else "Conv2d"
        kw = dict(device=device, dtype=dtype)

        def project() -> fl.Module:
            return fl.Linear(channels, channels, **kw) if use_linear_projection else fl.Conv2d(channels, channels, kernel_size=1, **kw)

        map_to_sequence = [StatefulFlatten(context="flatten", key="sizes", start_dim=2), fl.Transpose(1, 2)]
        sequence_to_map = [
            fl.Transpose(1, 2),
            fl.Parallel(fl.Identity(), fl.UseContext(context="flatten", key="sizes").compose(lambda sizes: sizes.pop())),
            fl.Unflatten(dim=2),
        ]
        norm = fl.GroupNorm(channels=channels, num_groups=num_groups, eps=1e-6, **kw)
        if use_linear_projection:  # project the sequence
            entry, leave = [norm, *map_to_sequence, project()], [project(), *sequence_to_map]
        else:                      # project the map
            entry, leave = [norm, project(), *map_to_sequence], [*sequence_to_map, project()]
        blocks = [
            CrossAttentionBlock(
                embedding_dim=channels, context_embedding_dim=context_embedding_dim, context_key=context_key,
                num_heads=num_attention_heads, use_bias=use_bias, **kw,
            )
            for _ in range(num_attention_layers)
        ]
        super().__init__(fl.Chain(*entry), fl.Chain(*blocks), fl.Chain(*leave))

    def init_context(self) -> Contexts:
        return {"flatten": {"sizes": []}}
