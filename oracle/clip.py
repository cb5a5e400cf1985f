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


# --------------------------------------------------------------------------------------- text towers
def text_layers(sd: SD, tokens: Tensor, *, first: int, last: int, heads: int, eps: float = 1e-5, quick_gelu: bool = False,
                x: Tensor | None = None, count: int | None = None) -> Tensor:
    """Transformer layers ``first .. last`` (1-based, inclusive) of a CLIPTextEncoder (clip/text_encoder.py:32-91, 94-188): pre-LN,
    CAUSAL self-attention, feed-forward with exact or quick (x * sigmoid(1.702 x)) GeLU.  With ``x`` None the token + position
    embeddings come first (:157-170)."""
    if x is None:
        x = sd["Sum.TokenEncoder.weight"][tokens] + sd["Sum.PositionalEncoder.Embedding.weight"][: tokens.shape[1]]
    total = count if count is not None else sum(k.endswith(".Residual_1.LayerNorm.weight") for k in sd)
    act = (lambda t: t * torch.sigmoid(1.702 * t)) if quick_gelu else ops.gelu
    for i in range(first, last + 1):
        p = "TransformerLayer" + ("" if total == 1 else f"_{i}")
        a = p + ".Residual_1.SelfAttention"
        t = _ln(sd, p + ".Residual_1.LayerNorm", x, eps)
        t = ops.sdpa(_lin(sd, a + ".Distribute.Linear_1", t), _lin(sd, a + ".Distribute.Linear_2", t), _lin(sd, a + ".Distribute.Linear_3", t),
                     heads, is_causal=True)
        x = x + _lin(sd, a + ".Linear", t)
        f = p + ".Residual_2.FeedForward"
        t = _ln(sd, p + ".Residual_2.LayerNorm", x, eps)
        x = x + _lin(sd, f + ".Linear_2", act(_lin(sd, f + ".Linear_1", t)))
    return x


def text_encoder(sd: SD, tokens: Tensor, *, num_layers: int, heads: int, quick_gelu: bool = False, eps: float = 1e-5) -> Tensor:
    """CLIPTextEncoder.forward on token ids: [B, 77] -> [B, 77, width]."""
    x = text_layers(sd, tokens, first=1, last=num_layers, heads=heads, eps=eps, quick_gelu=quick_gelu, count=num_layers)
    return _ln(sd, "LayerNorm", x, eps)


def split_double_text_encoder(sd: SD) -> tuple[dict, dict, Tensor]:
    """A DoubleTextEncoder state dict as (CLIPTextEncoderL keys, CLIPTextEncoderG keys, projection): the pooling adapter keeps
    the bigG tower's layers 1-31 under ``...TextEncoderWithPooling.CLIPTextEncoderG`` and its last layer + final LayerNorm, as a
    one-layer chain, under ``...TextEncoderWithPooling.Parallel.Chain.CLIPTextEncoderG``."""
    L, body, head = "Parallel.CLIPTextEncoderL.", "Parallel.TextEncoderWithPooling.CLIPTextEncoderG.", "Parallel.TextEncoderWithPooling.Parallel.Chain.CLIPTextEncoderG."
    sd_l = {k[len(L):]: v for k, v in sd.items() if k.startswith(L)}
    sd_g = {k[len(body):]: v for k, v in sd.items() if k.startswith(body)}
    sd_g |= {k[len(head):].replace("TransformerLayer.", "TransformerLayer_32."): v for k, v in sd.items() if k.startswith(head)}
    return sd_l, sd_g, sd["Parallel.TextEncoderWithPooling.Parallel.Chain.Linear.weight"]


def double_text_encoder(sd_l: SD, sd_g: SD, projection: Tensor, tokens_l: Tensor, tokens_g: Tensor, end_of_text: int = 49407) -> tuple[Tensor, Tensor]:
    """SDXL's DoubleTextEncoder (stable_diffusion_xl/text_encoder.py:13-101): both towers read before their last layer; the bigG
    tower's last layer + final LayerNorm + bias-free projection, read at the first end-of-text position, is the pooled embedding."""
    hidden_l = text_layers(sd_l, tokens_l, first=1, last=11, heads=12, quick_gelu=True, count=12)
    hidden_g = text_layers(sd_g, tokens_g, first=1, last=31, heads=20, count=32)
    top = _ln(sd_g, "LayerNorm", text_layers(sd_g, tokens_g, first=32, last=32, heads=20, x=hidden_g, count=32), 1e-5)
    projected = ops.linear(top, projection, None)
    at = (tokens_g == end_of_text).int().argmax(dim=1)
    pooled = projected[torch.arange(tokens_g.shape[0]), at]
    return torch.cat([hidden_l, hidden_g], dim=-1), pooled
