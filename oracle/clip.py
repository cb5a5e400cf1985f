"""Functional restatement of the CLIP vision tower over reference state-dict keys.

Reference: /root/reference/src/refiners/foundationals/clip/image_encoder.py
  ClassToken :10-27, PatchEncoder :30-64, ViTEmbeddings :86-141, TransformerLayer (FeedForward :67-83) :144-186,
  CLIPImageEncoder :189-276 ([CLS | patches] + positions -> pre-LayerNorm -> N pre-norm layers -> [CLS] -> post-LayerNorm ->
  bias-free projection).  TEST INFRASTRUCTURE - see oracle/__init__.py.
"""

from __future__ import annotations

from typing import Mapping

import torch
from torch import Tensor

from oracle import ops

SD = Mapping[str, Tensor]


def _lin(sd: SD, prefix: str, x: Tensor) -> Tensor:
    return ops.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _ln(sd: SD, prefix: str, x: Tensor, eps: float) -> Tensor:
    return ops.layer_norm(x, sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def transformer_layer(sd: SD, prefix: str, x: Tensor, heads: int, eps: float) -> Tensor:
    """x += Attn(LN(x)); x += W2 gelu(W1 LN(x))  (image_encoder.py:144-186)."""
    a = prefix + ".Residual_1.SelfAttention"
    t = _ln(sd, prefix + ".Residual_1.LayerNorm", x, eps)
    t = ops.sdpa(_lin(sd, a + ".Distribute.Linear_1", t), _lin(sd, a + ".Distribute.Linear_2", t), _lin(sd, a + ".Distribute.Linear_3", t), heads)
    x = x + _lin(sd, a + ".Linear", t)
    f = prefix + ".Residual_2.FeedForward"
    t = _ln(sd, prefix + ".Residual_2.LayerNorm", x, eps)
    return x + _lin(sd, f + ".Linear_2", ops.gelu(_lin(sd, f + ".Linear_1", t)))


def image_encoder(sd: SD, image: Tensor, *, patch_size: int, num_layers: int, num_heads: int, eps: float = 1e-5) -> Tensor:
    """CLIPImageEncoder.forward (:189-276): [B, 3, S, S] -> [B, output_dim]."""
    B = image.shape[0]
    w = sd["ViTEmbeddings.Concatenate.Chain.PatchEncoder.Conv2d.weight"]
    patches = ops.conv2d(image, w, None, stride=patch_size).reshape(B, w.shape[0], -1).transpose(1, 2)
    cls = sd["ViTEmbeddings.Concatenate.ClassToken.Parameter.weight"].expand(B, -1, -1)
    x = torch.cat([cls, patches], dim=1)
    x = x + sd["ViTEmbeddings.Residual.PositionalEncoder.Embedding.weight"][: x.shape[1]]
    x = _ln(sd, "LayerNorm_1", x, eps)
    for i in range(num_layers):
        name = "Chain.TransformerLayer" + ("" if num_layers == 1 else f"_{i + 1}")
        x = transformer_layer(sd, name, x, num_heads, eps)
    return _lin(sd, "Linear", _ln(sd, "LayerNorm_2", x[:, 0, :], eps))
