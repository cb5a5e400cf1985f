"""Functional restatement of DINOv2's ViT over reference state-dict keys.

Reference: /root/reference/src/refiners/foundationals/dinov2/vit.py
  ClassToken :12-29, PositionalEmbedding :32-52, InterpolateEmbedding :55-100, LayerScale :103-131,
  FeedForward :134-165, PatchEncoder :168-195, TransformerLayer :198-253, Registers :268-286, ViT :289-413.
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""

from __future__ import annotations

from math import isqrt
from typing import Mapping

import torch
from torch import Tensor
from torch.nn import functional as F

from oracle import ops

SD = Mapping[str, Tensor]


def _lin(sd: SD, prefix: str, x: Tensor) -> Tensor:
    return ops.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def positional_embedding(table: Tensor, image: Tensor, patch: int, mode: str, antialias: bool) -> Tensor:
    """InterpolateEmbedding (:55-100): [CLS] row kept, the M x M grid of patch positions resampled to the
    (H/p, W/p) grid of the input in fp32 (bicubic by default; antialias for the register models).  NB the
    reference names image dims 2 and 3 'W' and 'H'; what matters is that the target grid is (dim2/p, dim3/p)."""
    cls, grid = table[:1], table[1:]
    side = isqrt(grid.shape[0])
    assert side * side == grid.shape[0]
    target = (image.shape[2] // patch, image.shape[3] // patch)
    g = grid.reshape(1, side, side, -1).permute(0, 3, 1, 2).to(torch.float32)
    g = F.interpolate(g, size=target, mode=mode, antialias=antialias).to(table.dtype)
    return torch.cat([cls, g.permute(0, 2, 3, 1).reshape(-1, table.shape[1])], dim=0)


def transformer_layer(sd: SD, prefix: str, x: Tensor, heads: int, eps: float, swiglu: bool) -> Tensor:
    """x += ls1 * Attn(LN(x)); x += ls2 * MLP(LN(x)) (:198-253).  MLP = Linear, GeLU(erf) or GLU(SiLU), Linear."""
    p = prefix + ".Residual_1"
    t = ops.layer_norm(x, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"], eps)
    a = p + ".SelfAttention"
    t = ops.sdpa(_lin(sd, a + ".Distribute.Linear_1", t), _lin(sd, a + ".Distribute.Linear_2", t), _lin(sd, a + ".Distribute.Linear_3", t), heads)
    x = x + _lin(sd, a + ".Linear", t) * sd[p + ".LayerScale.weight"]
    p = prefix + ".Residual_2"
    t = ops.layer_norm(x, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"], eps)
    t = _lin(sd, p + ".FeedForward.Linear_1", t)
    if swiglu:
        value, gate = t.chunk(2, dim=-1)
        t = value * ops.silu(gate)
    else:
        t = ops.gelu(t)
    return x + _lin(sd, p + ".FeedForward.Linear_2", t) * sd[p + ".LayerScale.weight"]


def vit(
    sd: SD, image: Tensor, *, patch_size: int, num_layers: int, num_heads: int, norm_eps: float = 1e-6, num_registers: int = 0,
    swiglu: bool = False, interpolate_mode: str = "bicubic", interpolate_antialias: bool = False,
) -> Tensor:
    """ViT.forward (:289-413): tokens = [CLS | patches] + positions; registers spliced after [CLS]; layers; LayerNorm."""
    B = image.shape[0]
    w, b = sd["Concatenate.PatchEncoder.Conv2d.weight"], sd.get("Concatenate.PatchEncoder.Conv2d.bias")
    patches = ops.conv2d(image, w, b, stride=patch_size)
    patches = patches.reshape(B, w.shape[0], -1).transpose(1, 2)
    cls = sd["Concatenate.ClassToken.Parameter.weight"].expand(B, -1, -1)
    x = torch.cat([cls, patches], dim=1)
    pos = positional_embedding(sd["PositionalEncoder.PositionalEmbedding.Parameter.weight"], image, patch_size, interpolate_mode, interpolate_antialias)
    x = x + pos
    if num_registers:
        reg = sd["Registers.Parameter.weight"].expand(B, -1, -1)
        x = torch.cat([x[:, :1], reg, x[:, 1:]], dim=1)
    for i in range(num_layers):
        x = transformer_layer(sd, f"Transformer.TransformerLayer_{i + 1}", x, num_heads, norm_eps, swiglu)
    return ops.layer_norm(x, sd["LayerNorm.weight"], sd["LayerNorm.bias"], norm_eps)
