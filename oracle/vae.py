"""Functional restatement of the latent-diffusion VAE over reference state-dict keys.

Reference: /root/reference/src/refiners/foundationals/latent_diffusion/auto_encoder.py
  Resnet :41-82, Encoder :85-143, Decoder :146-207, LatentDiffusionAutoencoder.encode/decode :305-331;
  Downsample / Upsample: fluxion/layers/sampling.py:41-161; SelfAttention2d: fluxion/layers/attentions.py:388-489.
TEST INFRASTRUCTURE - see oracle/__init__.py.
"""

from __future__ import annotations

from typing import Mapping

import torch
from torch import Tensor
from torch.nn import functional as F

from oracle import ops

SD = Mapping[str, Tensor]
WIDTHS = (128, 256, 512, 512, 512)
ENCODER_SCALE = 0.18125


def _conv(sd: SD, prefix: str, x: Tensor, stride: int = 1, padding: int = 0) -> Tensor:
    return ops.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)


# Tiled inference (auto_encoder.py:209-251 FixedGroupNorm): when ``frozen_stats`` is a dict, every GroupNorm - keyed by
# its state-dict prefix - normalises with the (mean, var) per (sample, group) of the FIRST tensor it saw.
frozen_stats: dict | None = None


def _gn(sd: SD, prefix: str, x: Tensor, eps: float) -> Tensor:
    if frozen_stats is None:
        return ops.group_norm(x, 32, sd[prefix + ".weight"], sd[prefix + ".bias"], eps)
    B, C = x.shape[:2]
    g = x.reshape(B, 32, -1)
    if prefix not in frozen_stats:
        frozen_stats[prefix] = (g.mean(dim=2, keepdim=True), g.var(dim=2, keepdim=True, correction=0))
    mean, var = frozen_stats[prefix]
    y = ((g - mean) / torch.sqrt(var + eps)).reshape(x.shape)
    shape = (1, C) + (1,) * (x.ndim - 2)
    return y * sd[prefix + ".weight"].reshape(shape) + sd[prefix + ".bias"].reshape(shape)


def resnet(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """Sum(shortcut, Chain(GN, SiLU, conv3x3, GN, SiLU, conv3x3)) (:41-82); GroupNorm eps is the default 1e-5;
    the shortcut is a 1x1 conv only when its key exists (width change), else the identity."""
    h = _conv(sd, prefix + ".Chain.Conv2d_1", ops.silu(_gn(sd, prefix + ".Chain.GroupNorm_1", x, 1e-5)), padding=1)
    h = _conv(sd, prefix + ".Chain.Conv2d_2", ops.silu(_gn(sd, prefix + ".Chain.GroupNorm_2", h, 1e-5)), padding=1)
    skip = _conv(sd, prefix + ".Conv2d", x) if (prefix + ".Conv2d.weight") in sd else x
    return skip + h


def bottleneck_attention(sd: SD, prefix: str, x: Tensor) -> Tensor:
    """Residual(GroupNorm(eps 1e-6), SelfAttention2d(512 channels, ONE head)) (:113-116, :176-180)."""
    B, C, H, W = x.shape
    t = _gn(sd, prefix + ".GroupNorm", x, 1e-6).reshape(B, C, H * W).transpose(1, 2)
    p = prefix + ".SelfAttention2d"
    lin = lambda name, v: ops.linear(v, sd[f"{p}.{name}.weight"], sd.get(f"{p}.{name}.bias"))
    a = ops.sdpa(lin("Distribute.Linear_1", t), lin("Distribute.Linear_2", t), lin("Distribute.Linear_3", t), 1)
    return x + lin("Linear", a).transpose(1, 2).reshape(B, C, H, W)


def encoder(sd: SD, x: Tensor, prefix: str = "Encoder") -> Tensor:
    """:85-143.  Downsample(scale 2, padding 0) pads right/bottom by one pixel, then a stride-2 3x3 conv."""
    h = _conv(sd, prefix + ".Conv2d", x, padding=1)
    for i in range(5):
        p = f"{prefix}.Chain_1.Chain_{i + 1}"
        h = resnet(sd, p + ".Resnet_1", h)
        if i == 4:
            h = bottleneck_attention(sd, p + ".Residual", h)
        h = resnet(sd, p + ".Resnet_2", h)
        if i < 3:
            h = _conv(sd, p + ".Downsample.Conv2d", F.pad(h, (0, 1, 0, 1)), stride=2)
    h = _conv(sd, prefix + ".Chain_2.Conv2d", ops.silu(_gn(sd, prefix + ".Chain_2.GroupNorm", h, 1e-6)), padding=1)
    return _conv(sd, prefix + ".Chain_3.Conv2d", h)[:, :4]


def decoder(sd: SD, z: Tensor, prefix: str = "Decoder") -> Tensor:
    """:146-207.  Level 0 (512 ch): Resnet, attention, Resnet; levels 1-4: three Resnets; levels 1-3 end with the
    nearest-2x Upsample (+3x3 conv): `layer.insert(-1, Upsample)` APPENDS - fluxion's Chain.insert counts a negative
    index from one past the end (chain.py insert: index = len + index + 1)."""
    h = _conv(sd, prefix + ".Conv2d_2", _conv(sd, prefix + ".Conv2d_1", z), padding=1)
    for i in range(5):
        p = f"{prefix}.Chain_1.Chain_{i + 1}"
        h = resnet(sd, p + ".Resnet_1", h)
        if i == 0:
            h = bottleneck_attention(sd, p + ".Residual", h)
        h = resnet(sd, p + ".Resnet_2", h)
        if i == 0:
            continue
        h = resnet(sd, p + ".Resnet_3", h)
        if i <= 3:
            h = ops.nearest_upsample(h, (h.shape[-2] * 2, h.shape[-1] * 2))
            h = _conv(sd, p + ".Upsample.Conv2d", h, padding=1)
    return _conv(sd, prefix + ".Chain_2.Conv2d", ops.silu(_gn(sd, prefix + ".Chain_2.GroupNorm", h, 1e-6)), padding=1)


def encode(sd: SD, image: Tensor) -> Tensor:
    return ENCODER_SCALE * encoder(sd, image)


def decode(sd: SD, latents: Tensor) -> Tensor:
    return decoder(sd, latents / ENCODER_SCALE)


# ------------------------------------------------------------------------------------------ tiled inference
def blending_mask(height: int, width: int, blending: int, edges: tuple[bool, bool, bool, bool]) -> Tensor:
    """auto_encoder.py:254-279: linear ramps towards the borders that are not image borders (top, bottom, left, right)."""
    mask = torch.ones(height, width)
    if blending == 0:
        return mask
    n = min(blending, min(height, width) // 2)
    ramp = torch.linspace(0, 1, n)
    if not edges[0]:
        mask[:n] *= ramp[:, None]
    if not edges[1]:
        mask[-n:] *= ramp.flip(0)[:, None]
    if not edges[2]:
        mask[:, :n] *= ramp[None, :]
    if not edges[3]:
        mask[:, -n:] *= ramp.flip(0)[None, :]
    return mask


def latent_tiles(height: int, width: int, tile_h: int, tile_w: int, overlap: int) -> list[tuple[int, int, int, int]]:
    """auto_encoder.py:411-428: (top, left, bottom, right), columns outermost."""
    return [
        (y, x, min(height, y + tile_h), min(width, x + tile_w))
        for x in range(0, max(width - overlap, 1), tile_w - overlap)
        for y in range(0, max(height - overlap, 1), tile_h - overlap)
    ]


def capture_statistics(sd: SD, image01: Tensor, small01: Tensor) -> None:
    """auto_encoder.py:430-456: the resized copy ``small01`` (values in [0, 1]) is clamped to the range of the full image,
    moved to its per-channel mean / standard deviation, and run through encode + decode with ``frozen_stats`` empty."""
    small = small01.clamp(min=image01.min(), max=image01.max())
    std, mean = torch.std_mean(image01, dim=[0, 2, 3], keepdim=True)
    small_std, small_mean = torch.std_mean(small, dim=[0, 2, 3], keepdim=True)
    small = (small - small_mean) * (std / small_std) + mean
    decode(sd, encode(sd, 2 * small - 1))


def tiled(sd: SD, run, source: Tensor, latent_hw: tuple[int, int], tile_hw: tuple[int, int], blending: int, k_in: int, k_out: int,
          channels: int) -> Tensor:
    """auto_encoder.py:465-583 (encode: k_in 8, k_out 1; decode: k_in 1, k_out 8; ``tile_hw`` in pixels)."""
    H, W = latent_hw
    tiles = latent_tiles(H, W, tile_hw[0] // 8, tile_hw[1] // 8, blending // 8)
    if len(tiles) == 1:
        return run(sd, source)
    out = torch.zeros(1, channels, H * k_out, W * k_out)
    weight = torch.zeros_like(out)
    for top, left, bottom, right in tiles:
        piece = run(sd, source[:, :, top * k_in : bottom * k_in, left * k_in : right * k_in])
        mask = blending_mask((bottom - top) * k_out, (right - left) * k_out, blending * k_out // 8, (top == 0, bottom == H, left == 0, right == W))
        out[:, :, top * k_out : bottom * k_out, left * k_out : right * k_out] += piece * mask
        weight[:, :, top * k_out : bottom * k_out, left * k_out : right * k_out] += mask
    return out / weight
