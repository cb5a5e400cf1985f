"""Functional restatement of the SAM ViT image encoder.

Reference: /root/reference/src/refiners/foundationals/segment_anything/image_encoder.py
  PatchEncoder :9-34, PositionalEncoder :37-55, RelativePositionAttention :58-143,
  FusedSelfAttention :146-190, FeedForward :193-203 [sic], WindowPartition/Merge :206-236,
  TransformerLayer :239-283, Neck :286-310, SAMViT :316-368.
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""

from __future__ import annotations

from typing import Mapping

import torch
from torch import Tensor

from oracle import ops

SD = Mapping[str, Tensor]


def relative_position_attention(qkv: Tensor, horizontal_embedding: Tensor, vertical_embedding: Tensor, heads: int) -> Tensor:
    """softmax(q k^T d^-1/2 + rel_v[:, :, :, :, None] + rel_h[:, :, :, None, :]) v on a [B, H, W, 3C]
    fused projection (image_encoder.py:87-143).  rel_v[b,h,w,kh] = q[b,h,w] . vertical_embedding[h - kh + H - 1],
    rel_h[b,h,w,kw] = q[b,h,w] . horizontal_embedding[w - kw + W - 1]; q is unscaled in those terms.
    The vertical term is added first (the reference fixes this order in a comment, :97-101)."""
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    d = C // heads
    t = qkv.reshape(B, H * W, 3, heads, d).permute(2, 0, 3, 1, 4).reshape(3, B * heads, H * W, d)
    q, k, v = t[0], t[1], t[2]
    ah, aw = torch.arange(H, device=qkv.device), torch.arange(W, device=qkv.device)
    ih = ah[:, None] - ah[None, :] + H - 1
    iw = aw[:, None] - aw[None, :] + W - 1
    q5 = q.reshape(B * heads, H, W, d)
    rel_v = torch.einsum("bhwc,hkc->bhwk", q5, vertical_embedding[ih])   # [B', H, W, Hk]
    rel_h = torch.einsum("bhwc,wkc->bhwk", q5, horizontal_embedding[iw])  # [B', H, W, Wk]
    logits = (q * d**-0.5) @ k.transpose(-1, -2)
    logits = (logits.reshape(-1, H, W, H, W) + rel_v[..., :, None]) + rel_h[..., None, :]
    logits = logits.reshape(B * heads, H * W, H * W)
    logits = logits - logits.max(dim=-1, keepdim=True).values
    p = torch.exp(logits)
    p = p / p.sum(dim=-1, keepdim=True)
    o = p @ v
    return o.reshape(B, heads, H, W, d).permute(0, 2, 3, 1, 4).reshape(B, H, W, C)


def fused_self_attention(sd: SD, prefix: str, x: Tensor, heads: int) -> Tensor:
    """Linear(C -> 3C) -> RelativePositionAttention -> Linear(C -> C)  (image_encoder.py:146-190)."""
    qkv = ops.linear(x, sd[prefix + ".Linear_1.weight"], sd.get(prefix + ".Linear_1.bias"))
    o = relative_position_attention(
        qkv,
        sd[prefix + ".RelativePositionAttention.horizontal_embedding"],
        sd[prefix + ".RelativePositionAttention.vertical_embedding"],
        heads,
    )
    return ops.linear(o, sd[prefix + ".Linear_2.weight"], sd[prefix + ".Linear_2.bias"])


def window_partition(x: Tensor, window: int) -> tuple[Tensor, tuple[int, int, int, int]]:
    """Zero-pad H, W up to multiples of ``window`` and cut into [B * nH * nW, window, window, C]
    (image_encoder.py:206-219)."""
    B, H, W, C = x.shape
    ph, pw = (window - H % window) % window, (window - W % window) % window
    if ph or pw:
        x = torch.nn.functional.pad(x, (0, 0, 0, pw, 0, ph))
    Hp, Wp = H + ph, W + pw
    x = x.reshape(B, Hp // window, window, Wp // window, window, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, window, window, C), (H, W, Hp, Wp)


def window_merge(x: Tensor, window: int, dims: tuple[int, int, int, int]) -> Tensor:
    """Inverse of window_partition, cropping the padding (image_encoder.py:222-236)."""
    H, W, Hp, Wp = dims
    B = x.shape[0] // ((Hp // window) * (Wp // window))
    x = x.reshape(B, Hp // window, Wp // window, window, window, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, -1)
    return x[:, :H, :W, :]


def transformer_layer(sd: SD, prefix: str, x: Tensor, heads: int, window: int | None, eps: float = 1e-6) -> Tensor:
    """image_encoder.py:239-283: x + attn(LN(x)) (windowed or global), then x + MLP(LN(x)) with
    exact GeLU."""
    r1, r2 = prefix + ".Residual_1", prefix + ".Residual_2"
    h = ops.layer_norm(x, sd[r1 + ".LayerNorm.weight"], sd[r1 + ".LayerNorm.bias"], eps)
    if window is not None:
        h, dims = window_partition(h, window)
        h = fused_self_attention(sd, r1 + ".FusedSelfAttention", h, heads)
        h = window_merge(h, window, dims)
    else:
        h = fused_self_attention(sd, r1 + ".FusedSelfAttention", h, heads)
    x = x + h
    h = ops.layer_norm(x, sd[r2 + ".LayerNorm.weight"], sd[r2 + ".LayerNorm.bias"], eps)
    f = r2 + ".FeedForward"
    h = ops.linear(ops.gelu(ops.linear(h, sd[f + ".Linear_1.weight"], sd[f + ".Linear_1.bias"])), sd[f + ".Linear_2.weight"], sd[f + ".Linear_2.bias"])
    return x + h


def patch_encoder(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """16x16 stride-16 conv then NCHW -> NHWC (image_encoder.py:9-34)."""
    w = sd[prefix + ".Conv2d.weight"]
    y = ops.conv2d(x, w, sd.get(prefix + ".Conv2d.bias"), stride=w.shape[-1], padding=0)
    return y.permute(0, 2, 3, 1)


def neck(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """NHWC -> NCHW, conv1x1 (no bias), LayerNorm2d, conv3x3 (no bias), LayerNorm2d (image_encoder.py:286-310)."""
    h = ops.conv2d(x.permute(0, 3, 1, 2), sd[prefix + ".Conv2d_1.weight"], None)
    h = ops.layer_norm_2d(h, sd[prefix + ".LayerNorm2d_1.weight"], sd[prefix + ".LayerNorm2d_1.bias"], 1e-6)
    h = ops.conv2d(h, sd[prefix + ".Conv2d_2.weight"], None, padding=1)
    return ops.layer_norm_2d(h, sd[prefix + ".LayerNorm2d_2.weight"], sd[prefix + ".LayerNorm2d_2.bias"], 1e-6)


def sam_vit(sd: SD, x: Tensor, num_layers: int, heads: int, global_indices: tuple[int, ...], window: int = 14) -> Tensor:
    """SAMViT forward (image_encoder.py:316-368): patches + learned positions, ``num_layers``
    transformer layers (global attention at ``global_indices``), neck."""
    h = patch_encoder(sd, "PatchEncoder", x)
    h = h + sd["PositionalEncoder.Parameter.weight"]
    for i in range(num_layers):
        name = "Transformer.TransformerLayer" + ("" if num_layers == 1 else f"_{i + 1}")
        h = transformer_layer(sd, name, h, heads, None if i in global_indices else window)
    return neck(sd, "Neck", h)
