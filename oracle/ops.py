"""Leaf operations of the hot path, restated as plain fp32 tensor math (CPU).

Deliberately written with elementary tensor ops (matmul, mean, exp ...) rather than the
torch.nn.functional fused calls the reference uses, so that it is an independent statement of
the arithmetic.  TEST INFRASTRUCTURE - see oracle/__init__.py.
"""

from __future__ import annotations

import math

import torch
from torch import Tensor

# The timed CPU baseline (bench.py `cpu_baseline` / `--impl reference`) must cost what the
# REFERENCE costs on the host, and the reference calls ATen's fused CPU kernels:
#   F.conv2d (conv.py:6), F.scaled_dot_product_attention (attentions.py:29-34),
#   F.group_norm / F.layer_norm (norm.py:14,52), F.linear (linear.py:9).
# With FAST = True the functions below make exactly those calls; the default (False) keeps the
# independent elementary-op restatement that the parity tests check the kernels against.
FAST = False


def linear(x: Tensor, weight: Tensor, bias: Tensor | None = None) -> Tensor:
    """y = x W^T + b.  Reference: fluxion/layers/linear.py:9-58 (torch.nn.Linear.forward)."""
    if FAST:
        return torch.nn.functional.linear(x, weight, bias)
    y = x @ weight.transpose(-1, -2)
    return y if bias is None else y + bias


def conv2d(x: Tensor, weight: Tensor, bias: Tensor | None, stride: int = 1, padding: int = 0) -> Tensor:
    """Zero-padded cross-correlation, NCHW.  Reference: fluxion/layers/conv.py:6-61.
    Restated as unfold (im2col) + matmul."""
    if FAST:
        return torch.nn.functional.conv2d(x, weight, bias, stride=stride, padding=padding)
    B, C, H, W = x.shape
    Co, Ci, R, S = weight.shape
    assert Ci == C
    Ho = (H + 2 * padding - R) // stride + 1
    Wo = (W + 2 * padding - S) // stride + 1
    cols = torch.nn.functional.unfold(x, (R, S), padding=padding, stride=stride)  # [B, C*R*S, Ho*Wo]
    y = weight.reshape(Co, -1) @ cols
    if bias is not None:
        y = y + bias[None, :, None]
    return y.reshape(B, Co, Ho, Wo)


def group_norm(x: Tensor, num_groups: int, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    """Per (sample, group) standardisation with biased variance, then per-channel affine.
    Reference: fluxion/layers/norm.py:52-92 (torch.nn.GroupNorm)."""
    if FAST:
        return torch.nn.functional.group_norm(x, num_groups, weight, bias, eps)
    B, C = x.shape[:2]
    g = x.reshape(B, num_groups, -1)
    mean = g.mean(dim=-1, keepdim=True)
    var = ((g - mean) ** 2).mean(dim=-1, keepdim=True)
    y = ((g - mean) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = (1, C) + (1,) * (x.ndim - 2)
    return y * weight.reshape(shape) + bias.reshape(shape)


def layer_norm(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    """Row standardisation over the last dim.  Reference: fluxion/layers/norm.py:14-49."""
    if FAST:
        return torch.nn.functional.layer_norm(x, (x.shape[-1],), weight, bias, eps)
    mean = x.mean(dim=-1, keepdim=True)
    var = ((x - mean) ** 2).mean(dim=-1, keepdim=True)
    return (x - mean) / torch.sqrt(var + eps) * weight + bias


def layer_norm_2d(x: Tensor, weight: Tensor, bias: Tensor, eps: float) -> Tensor:
    """Channel-wise LN of an NCHW map.  Reference: fluxion/layers/norm.py:119-127."""
    mean = x.mean(1, keepdim=True)
    var = (x - mean).pow(2).mean(1, keepdim=True)
    return weight[:, None, None] * ((x - mean) / torch.sqrt(var + eps)) + bias[:, None, None]


def silu(x: Tensor) -> Tensor:
    """x * sigmoid(x).  Reference: fluxion/layers/activations.py:31-41."""
    if FAST:
        return torch.nn.functional.silu(x)
    return x / (1.0 + torch.exp(-x))


def relu(x: Tensor) -> Tensor:
    """fluxion/layers/activations.py ReLU."""
    return torch.clamp_min(x, 0.0)


def gelu(x: Tensor) -> Tensor:
    """Exact (erf) GeLU.  Reference: fluxion/layers/activations.py:83-114 (approximation NONE)."""
    if FAST:
        return torch.nn.functional.gelu(x)
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def glu_gelu(x: Tensor) -> Tensor:
    """value * gelu(gate), (value, gate) = halves of the last dim.
    Reference: fluxion/layers/activations.py:136-160."""
    value, gate = x.chunk(2, dim=-1)
    return value * gelu(gate)


def sdpa(q: Tensor, k: Tensor, v: Tensor, num_heads: int, is_causal: bool = False) -> Tensor:
    """softmax(Q K^T / sqrt(d)) V per head on [B, S, C] operands.
    Reference: fluxion/layers/attentions.py:115-202 (split heads :177-192, SDPA :15-34, merge :194-202)."""
    B, Sq, C = q.shape
    Sk = k.shape[1]
    d = C // num_heads
    qh = q.reshape(B, Sq, num_heads, d).transpose(1, 2)
    kh = k.reshape(B, Sk, num_heads, d).transpose(1, 2)
    vh = v.reshape(B, Sk, num_heads, d).transpose(1, 2)
    if FAST:
        o = torch.nn.functional.scaled_dot_product_attention(qh, kh, vh, is_causal=is_causal)
        return o.transpose(1, 2).reshape(B, Sq, C)
    logits = (qh @ kh.transpose(-1, -2)) / math.sqrt(d)
    if is_causal:
        mask = torch.ones(Sq, Sk, dtype=torch.bool, device=q.device).tril()
        logits = logits.masked_fill(~mask, float("-inf"))
    logits = logits - logits.max(dim=-1, keepdim=True).values
    p = torch.exp(logits)
    p = p / p.sum(dim=-1, keepdim=True)
    return (p @ vh).transpose(1, 2).reshape(B, Sq, C)


def lora_linear(x: Tensor, weight: Tensor, bias: Tensor | None, loras: list[tuple[Tensor, Tensor, float]]) -> Tensor:
    """LoraAdapter(Linear): W x + b + sum_i scale_i * up_i(down_i(x)).
    Reference: fluxion/adapters/lora.py:14-99 (Lora = down, up, Multiply), :383-448 (Sum)."""
    y = linear(x, weight, bias)
    for down, up, scale in loras:
        y = y + scale * linear(linear(x, down), up)
    return y


def sinusoidal_embedding(x: Tensor, embedding_dim: int) -> Tensor:
    """cos | sin of x * 10000^(-i/half).  Reference: latent_diffusion/range_adapter.py:11-22."""
    half = embedding_dim // 2
    exponent = -math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=x.device) / half
    emb = x.unsqueeze(1).float() * torch.exp(exponent).unsqueeze(0)
    return torch.cat([torch.cos(emb), torch.sin(emb)], dim=-1)


def nearest_upsample(x: Tensor, size: tuple[int, int]) -> Tensor:
    """Nearest-neighbour resize (index = floor(dst * in / out)).
    Reference: fluxion/layers/sampling.py:13-38 via fluxion/utils.py interpolate."""
    H, W = x.shape[-2:]
    hi = (torch.arange(size[0], device=x.device) * H) // size[0]
    wi = (torch.arange(size[1], device=x.device) * W) // size[1]
    return x[..., hi[:, None], wi[None, :]]
