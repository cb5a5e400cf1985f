"""Functional restatement of the SD1.5 and SDXL UNet forward passes over a reference state dict.

The functions walk explicit block tables (not a module tree) and read weights by their
reference state-dict keys, so they are independent both of the reference's Chain machinery and
of refiners_b200's fluxion mirror.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Reference (paths under /root/reference/src/refiners/foundationals/latent_diffusion/):
  unet.py:6-79                      ResidualBlock / ResidualAccumulator / ResidualConcatenator
  range_adapter.py:11-86            sinusoidal embedding, RangeEncoder, RangeAdapter2d
  cross_attention.py:25-175         CrossAttentionBlock / CrossAttentionBlock2d
  stable_diffusion_1/unet.py:16-249 SD1UNet
  stable_diffusion_xl/unet.py:20-351 SDXLUNet
"""

from __future__ import annotations

from typing import Mapping

import torch
from torch import Tensor

from oracle import ops

SD = Mapping[str, Tensor]


class Weights(dict):
    """A state dict (reference keys) plus the adapter weights that the reference keeps in injected
    modules.  Everything is addressed by the ORIGINAL path of the adapted leaf, so the restatement does
    not depend on how an adapter renames its target's state-dict key:
      loras[path]  list of (down [r, in], up [out, r], scale) on the Linear at ``path``
                   (fluxion/adapters/lora.py:383-448: Sum(target, Chain(down, up, Multiply(scale))...))
      ip[path]     (W_k', W_v') of the IP-Adapter ImageCrossAttention on the Attention at ``path``
                   (latent_diffusion/image_prompt.py:237-309), with ``ip_scale`` and ``ip_embedding``."""

    def __init__(self, tensors: Mapping[str, Tensor], loras: dict | None = None, ip: dict | None = None,
                 ip_scale: float = 1.0, ip_embedding: Tensor | None = None) -> None:
        super().__init__(tensors)
        self.loras = loras or {}
        self.ip = ip or {}
        self.ip_scale = ip_scale
        self.ip_embedding = ip_embedding


def _lin(sd: SD, prefix: str, x: Tensor) -> Tensor:
    loras = getattr(sd, "loras", None)
    if loras and prefix in loras:
        return ops.lora_linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), loras[prefix])
    return ops.linear(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"))


def _conv(sd: SD, prefix: str, x: Tensor, stride: int = 1, padding: int = 0) -> Tensor:
    return ops.conv2d(x, sd[prefix + ".weight"], sd.get(prefix + ".bias"), stride=stride, padding=padding)


def range_encoder(sd: SD, prefix: str, timestep: Tensor, dtype: torch.dtype) -> Tensor:
    """sinusoid(320) -> cast -> Linear -> SiLU -> Linear  (range_adapter.py:25-44)."""
    e = ops.sinusoidal_embedding(timestep, 320).to(dtype)
    return _lin(sd, prefix + ".Linear_2", ops.silu(_lin(sd, prefix + ".Linear_1", e)))


def residual_block(sd: SD, prefix: str, x: Tensor, temb: Tensor, eps: float = 1e-5) -> Tensor:
    """unet.py:6-51 with the RangeAdapter2d injected around the first conv (sdxl/unet.py:286-295):
    h = conv1(silu(gn1(x))) + Linear(silu(temb))[:, :, None, None]; h = conv2(silu(gn2(h)));
    out = h + shortcut(x)."""
    c = prefix + ".Chain"
    h = ops.silu(ops.group_norm(x, 32, sd[c + ".GroupNorm_1.weight"], sd[c + ".GroupNorm_1.bias"], eps))
    h = _conv(sd, c + ".RangeAdapter2d.Conv2d", h, padding=1)
    t = _lin(sd, c + ".RangeAdapter2d.Chain.Linear", ops.silu(temb))
    h = h + t.reshape(t.shape[0], -1, 1, 1)
    h = ops.silu(ops.group_norm(h, 32, sd[c + ".GroupNorm_2.weight"], sd[c + ".GroupNorm_2.bias"], eps))
    h = _conv(sd, c + ".Conv2d", h, padding=1)
    shortcut = _conv(sd, prefix + ".Conv2d", x) if (prefix + ".Conv2d.weight") in sd else x
    return h + shortcut


# Self-attention guidance probes (self_attention_guidance.py:22-59, placed by stable_diffusion_1/ and
# stable_diffusion_xl/self_attention_guidance.py:20-31): when ``middle_probe`` is a dict, a UNet pass leaves in it the
# middle block's feature-map size ("shape") and the attention probabilities of its FIRST self-attention ("map").
middle_probe: dict | None = None


# StyleAligned (style_aligned.py:136-265): when ``style_aligned_scale`` is a float, every SELF-attention shares the first
# image of each guidance half: q, k are moved to its per-channel token statistics, k and v get its (scaled) rows appended.
style_aligned_scale: float | None = None


def _style_aligned(q: Tensor, k: Tensor, v: Tensor, scale: float, eps: float = 1e-8) -> tuple[Tensor, Tensor, Tensor]:
    half = q.shape[0] // 2

    def reference(x: Tensor) -> Tensor:  # ExtractReferenceFeatures :13-44
        return torch.stack((x[0], x[half])).repeat_interleave(half, dim=0)

    def adain(x: Tensor, ref: Tensor) -> Tensor:  # :47-92 (torch.std: unbiased)
        mean, std = x.mean(dim=-2, keepdim=True), x.std(dim=-2, keepdim=True)
        return (x - mean) / (std + eps) * ref.std(dim=-2, keepdim=True) + ref.mean(dim=-2, keepdim=True)

    def scaled(ref: Tensor) -> Tensor:  # ScaleReferenceFeatures :95-133: every image but the first of its half
        factor = torch.full((q.shape[0], 1, 1), scale, dtype=ref.dtype, device=ref.device)
        factor[0], factor[half] = 1.0, 1.0
        return ref * factor

    rq, rk, rv = reference(q), reference(k), reference(v)
    return adain(q, rq), torch.cat((adain(k, rk), scaled(rk)), dim=-2), torch.cat((v, scaled(rv)), dim=-2)


def attention(sd: SD, prefix: str, q_in: Tensor, kv_in: Tensor, heads: int) -> Tensor:
    """fl.Attention: Distribute(Wq, Wk, Wv) -> SDPA -> Wo  (fluxion/layers/attentions.py:205-316)."""
    q = _lin(sd, prefix + ".Distribute.Linear_1", q_in)
    k = _lin(sd, prefix + ".Distribute.Linear_2", kv_in)
    v = _lin(sd, prefix + ".Distribute.Linear_3", kv_in)
    if style_aligned_scale is not None and prefix.endswith(".SelfAttention"):
        q, k, v = _style_aligned(q, k, v, style_aligned_scale)
    if middle_probe is not None and "MiddleBlock" in prefix and prefix.endswith(".SelfAttention") and "map" not in middle_probe:
        split = lambda t: t.reshape(t.shape[0], t.shape[1], heads, -1).transpose(1, 2)  # noqa: E731
        qh, kh = split(q), split(k)
        middle_probe["map"] = torch.softmax(qh @ kh.transpose(-1, -2) / qh.shape[-1] ** 0.5, dim=-1)
    o = ops.sdpa(q, k, v, heads)
    ip = getattr(sd, "ip", None)
    if ip and prefix in ip:
        # CrossAttentionAdapter (image_prompt.py:312-347): the SDPA becomes Sum(SDPA(q, k, v),
        # Chain(SDPA(q, W_k' e, W_v' e), Multiply(scale))) in front of the output projection
        wk, wv = ip[prefix]
        e = sd.ip_embedding  # type: ignore[attr-defined]
        o = o + sd.ip_scale * ops.sdpa(q, ops.linear(e, wk), ops.linear(e, wv), heads)  # type: ignore[attr-defined]
    return _lin(sd, prefix + ".Linear", o)


def cross_attention_block(sd: SD, prefix: str, x: Tensor, context: Tensor, heads: int) -> Tensor:
    """cross_attention.py:25-73: three pre-LN residual branches (self-attn, cross-attn, GEGLU MLP)."""
    r1, r2, r3 = prefix + ".Residual_1", prefix + ".Residual_2", prefix + ".Residual_3"
    h = ops.layer_norm(x, sd[r1 + ".LayerNorm.weight"], sd[r1 + ".LayerNorm.bias"], 1e-5)
    x = x + attention(sd, r1 + ".SelfAttention", h, h, heads)
    h = ops.layer_norm(x, sd[r2 + ".LayerNorm.weight"], sd[r2 + ".LayerNorm.bias"], 1e-5)
    x = x + attention(sd, r2 + ".Attention", h, context, heads)
    h = ops.layer_norm(x, sd[r3 + ".LayerNorm.weight"], sd[r3 + ".LayerNorm.bias"], 1e-5)
    h = _lin(sd, r3 + ".Linear_2", ops.glu_gelu(_lin(sd, r3 + ".Linear_1", h)))
    return x + h


def cross_attention_2d(sd: SD, prefix: str, x: Tensor, context: Tensor, heads: int, layers: int, linear_proj: bool) -> Tensor:
    """cross_attention.py:92-175: GN(eps 1e-6) -> project in -> [B, HW, C] -> N blocks -> project out,
    residual around everything.  SDXL projects with Linear, SD1.5 with 1x1 conv."""
    B, C, H, W = x.shape
    if middle_probe is not None and "MiddleBlock" in prefix:
        middle_probe["shape"] = (H, W)
    c1, c2, c3 = prefix + ".Chain_1", prefix + ".Chain_2", prefix + ".Chain_3"
    h = ops.group_norm(x, 32, sd[c1 + ".GroupNorm.weight"], sd[c1 + ".GroupNorm.bias"], 1e-6)
    if linear_proj:
        h = _lin(sd, c1 + ".Linear", h.flatten(2).transpose(1, 2))
    else:
        h = _conv(sd, c1 + ".Conv2d", h).flatten(2).transpose(1, 2)
    for i in range(layers):
        name = c2 + (".CrossAttentionBlock" if layers == 1 else f".CrossAttentionBlock_{i + 1}")
        h = cross_attention_block(sd, name, h, context, heads)
    if linear_proj:
        h = _lin(sd, c3 + ".Linear", h).transpose(1, 2).reshape(B, C, H, W)
    else:
        h = _conv(sd, c3 + ".Conv2d", h.transpose(1, 2).reshape(B, C, H, W))
    return h + x


# ------------------------------------------------------------------------------------ SD 1.5
# (kind, *args): "res" = ResidualBlock(+attention?), "down", per reference DownBlocks/UpBlocks
_SD1_DOWN = [("in",), ("res", True), ("res", True), ("down",), ("res", True), ("res", True), ("down",),
             ("res", True), ("res", True), ("down",), ("res", False), ("res", False)]
_SD1_UP = [(False, False), (False, False), (False, True), (True, False), (True, False), (True, True),
           (True, False), (True, False), (True, True), (True, False), (True, False), (True, False)]


def condition_encoder(sd: SD, prefix: str, image: Tensor) -> Tensor:
    """ControlNet ConditionEncoder (stable_diffusion_1/controlnet.py:16-68): conv-SiLU stem, three
    (conv, SiLU, stride-2 conv, SiLU) stages, output conv to 320 channels at 1/8 resolution."""
    h = ops.silu(_conv(sd, prefix + ".Chain_1.Conv2d", image, padding=1))
    for i in (2, 3, 4):
        h = ops.silu(_conv(sd, f"{prefix}.Chain_{i}.Conv2d_1", h, padding=1))
        h = ops.silu(_conv(sd, f"{prefix}.Chain_{i}.Conv2d_2", h, stride=2, padding=1))
    return _conv(sd, prefix + ".Conv2d", h, padding=1)


def sd1_controlnet(
    sd: SD, x: Tensor, timestep: Tensor, clip_text_embedding: Tensor, condition: Tensor, scale: float = 1.0,
    scale_decay: float = 1.0, prefix: str = "Controlnet",
) -> list[Tensor]:
    """Controlnet (stable_diffusion_1/controlnet.py:71-173): the UNet encoder + middle block run on
    the first four latent channels (:103), the encoded condition is added after the input conv
    (:111-116), and after every one of the 12 down entries and after the middle block a 1x1 conv taps
    the activation; tap n contributes ``tap * scale * scale_decay ** (12 - n)`` to residual slot n
    (:153-173).  The taps are Passthroughs: the control copy's own activations are not modified."""
    dtype = x.dtype
    temb = range_encoder(sd, prefix + ".TimestepEncoder.RangeEncoder", timestep, dtype)
    h = x[:, :4]
    deltas: list[Tensor] = []
    for i, entry in enumerate(_SD1_DOWN):
        p = f"{prefix}.DownBlocks.Chain_{i + 1}"
        if entry[0] == "in":
            h = _conv(sd, p + ".Conv2d", h, padding=1)
            h = h + condition_encoder(sd, p + ".Residual.ConditionEncoder", condition)
        elif entry[0] == "down":
            h = _conv(sd, p + ".Downsample.Conv2d", h, stride=2, padding=1)
        else:
            h = residual_block(sd, p + ".ResidualBlock", h, temb)
            if entry[1]:
                h = cross_attention_2d(sd, p + ".CLIPLCrossAttention", h, clip_text_embedding, 8, 1, False)
        deltas.append(_conv(sd, p + ".Passthrough.Conv2d", h) * scale * scale_decay ** float(12 - i))
    m = prefix + ".MiddleBlock"
    h = residual_block(sd, m + ".ResidualBlock_1", h, temb)
    h = cross_attention_2d(sd, m + ".CLIPLCrossAttention", h, clip_text_embedding, 8, 1, False)
    h = residual_block(sd, m + ".ResidualBlock_2", h, temb)
    deltas.append(_conv(sd, m + ".Passthrough.Conv2d", h) * scale * scale_decay ** 0.0)
    return deltas


def sd1_unet(sd: SD, x: Tensor, timestep: Tensor, clip_text_embedding: Tensor, residuals: list[Tensor] | None = None,
             t2i: tuple[tuple[Tensor, ...], float] | None = None) -> Tensor:
    """SD1UNet forward (stable_diffusion_1/unet.py:165-249): 12 down entries each recording a
    skip, the middle block (added to the 13th residual slot: 0.0 for the plain UNet), 12 up entries
    each consuming one skip.  ``residuals`` are the 13 ControlNet corrections already sitting in the
    slots when the UNet runs (the control copy is child 0 of the UNet, controlnet.py:196-203)."""
    dtype = x.dtype
    temb = range_encoder(sd, "TimestepEncoder.RangeEncoder", timestep, dtype)
    skips: list[Tensor] = []
    h = x
    for i, entry in enumerate(_SD1_DOWN):
        p = f"DownBlocks.Chain_{i + 1}"
        if entry[0] == "in":
            h = _conv(sd, p + ".Conv2d", h, padding=1)
        elif entry[0] == "down":
            h = _conv(sd, p + ".Downsample.Conv2d", h, stride=2, padding=1)
        else:
            h = residual_block(sd, p + ".ResidualBlock", h, temb)
            if entry[1]:
                h = cross_attention_2d(sd, p + ".CLIPLCrossAttention", h, clip_text_embedding, 8, 1, False)
        if t2i is not None and i in (2, 5, 8, 11):
            # T2IFeatures (stable_diffusion_1/t2i_adapter.py:24-31): a Residual in front of the tap - the encoder continues from it
            h = h + t2i[1] * t2i[0][(2, 5, 8, 11).index(i)]
        # ResidualAccumulator: residuals[n] = x + residuals[n]; the accumulator is a Passthrough, so the
        # encoder itself continues from the uncorrected h
        skips.append(h if residuals is None else h + residuals[i])
    m = "Sum.MiddleBlock"
    mid = residual_block(sd, m + ".ResidualBlock_1", h, temb)
    mid = cross_attention_2d(sd, m + ".CLIPLCrossAttention", mid, clip_text_embedding, 8, 1, False)
    mid = residual_block(sd, m + ".ResidualBlock_2", mid, temb)
    # Sum(UseContext residuals[-1], MiddleBlock): the 13th residual slot is never written by the
    # plain UNet (it exists for ControlNet), so this adds the initial 0.0
    h = (0.0 if residuals is None else residuals[12]) + mid
    for n, (attn, up) in enumerate(_SD1_UP):
        p = f"UpBlocks.Chain_{n + 1}"
        h = torch.cat([h, skips[-n - 1]], dim=1)  # ResidualConcatenator(-n-2) over 12 skips + 1 spare slot
        h = residual_block(sd, p + ".ResidualBlock", h, temb)
        if attn:
            h = cross_attention_2d(sd, p + ".CLIPLCrossAttention", h, clip_text_embedding, 8, 1, False)
        if up:
            h = ops.nearest_upsample(h, (h.shape[-2] * 2, h.shape[-1] * 2))
            h = _conv(sd, p + ".Upsample.Conv2d", h, padding=1)
    h = ops.silu(ops.group_norm(h, 32, sd["Chain.GroupNorm.weight"], sd["Chain.GroupNorm.bias"], 1e-5))
    return _conv(sd, "Chain.Conv2d", h, padding=1)


# -------------------------------------------------------------------------------------- SDXL
# (cout, attention layers, heads) per level entry; None marks Downsample
_SDXL_DOWN = [("in",), ("res", 0, 0), ("res", 0, 0), ("down",), ("res", 2, 10), ("res", 2, 10), ("down",),
              ("res", 10, 20), ("res", 10, 20)]
_SDXL_UP = [(10, 20, False), (10, 20, False), (10, 20, True), (2, 10, False), (2, 10, False), (2, 10, True),
            (0, 0, False), (0, 0, False), (0, 0, False)]


def sdxl_timestep_embedding(sd: SD, timestep: Tensor, pooled_text_embedding: Tensor, time_ids: Tensor, dtype: torch.dtype) -> Tensor:
    """TimestepEncoder (sdxl/unet.py:62-90) = RangeEncoder(timestep) + TextTimeEmbedding (:20-59):
    MLP(cat(pooled[B,1280], sinusoid256(time_ids[B,6]) flattened to [B,1536]))."""
    t = range_encoder(sd, "TimestepEncoder.Sum.Chain.RangeEncoder", timestep, dtype)
    ids = ops.sinusoidal_embedding(time_ids.unsqueeze(-1), 256)  # [B, 1, 6, 256]
    ids = ids.reshape(ids.shape[0], -1)
    tt = torch.cat([pooled_text_embedding, ids], dim=1).to(dtype)
    p = "TimestepEncoder.Sum.TextTimeEmbedding"
    tt = _lin(sd, p + ".Linear_2", ops.silu(_lin(sd, p + ".Linear_1", tt)))
    return t + tt


def sdxl_control_lora(
    sd: SD, own: SD, x: Tensor, timestep: Tensor, clip_text_embedding: Tensor, pooled_text_embedding: Tensor, time_ids: Tensor,
    condition: Tensor, scale: float = 1.0,
) -> list[Tensor]:
    """ControlLora (stable_diffusion_xl/control_lora.py:144-248): a structural copy of TimestepEncoder +
    DownBlocks + MiddleBlock whose weighted leaves are the UNet's own, with LoRAs attached inside the copy
    only (``sd.loras``, keyed by the leaf's path inside the copy), the encoded condition added at the end of
    the first block (:190-202) and a 1x1 ZeroConvolution after each of the 9 down entries and after the middle
    block contributing ``scale * conv(h)`` to residual slot n (:90-132, :203-233).  ``own`` holds the copy's
    own parameters (ConditionEncoder, ZeroConvolutions) under their paths inside the copy."""
    dtype = x.dtype
    temb = sdxl_timestep_embedding(sd, timestep, pooled_text_embedding, time_ids, dtype)
    deltas: list[Tensor] = []
    h = x
    for i, entry in enumerate(_SDXL_DOWN):
        p = f"DownBlocks.Chain_{i + 1}"
        if entry[0] == "in":
            h = _conv(sd, p + ".Conv2d", h, padding=1)
        elif entry[0] == "down":
            h = _conv(sd, p + ".Downsample.Conv2d", h, stride=2, padding=1)
        else:
            h = residual_block(sd, p + ".ResidualBlock", h, temb)
            if entry[1]:
                h = cross_attention_2d(sd, p + ".SDXLCrossAttention", h, clip_text_embedding, entry[2], entry[1], True)
        deltas.append(_conv(own, p + ".ZeroConvolution.Conv2d", h) * scale)
        if entry[0] == "in":
            # the condition Residual is APPENDED to the first block (control_lora.py:190-202), i.e. it sits after the
            # accumulator that becomes ZeroConvolution 0: tap 0 sees conv_in(x) without the condition
            h = h + condition_encoder(own, p + ".Residual.ConditionEncoder", condition)
    m = "MiddleBlock"
    h = residual_block(sd, m + ".ResidualBlock_1", h, temb)
    h = cross_attention_2d(sd, m + ".SDXLCrossAttention", h, clip_text_embedding, 20, 10, True)
    h = residual_block(sd, m + ".ResidualBlock_2", h, temb)
    deltas.append(_conv(own, m + ".ZeroConvolution.Conv2d", h) * scale)
    return deltas


def sdxl_unet(
    sd: SD, x: Tensor, timestep: Tensor, clip_text_embedding: Tensor, pooled_text_embedding: Tensor, time_ids: Tensor,
    residuals: list[Tensor] | None = None, t2i: tuple[tuple[Tensor, ...], float] | None = None,
) -> Tensor:
    """SDXLUNet forward (sdxl/unet.py:258-351).  ``t2i`` = (the four T2I-Adapter feature maps, scale): added in front of the
    taps of encoder entries 3, 5, 8 and at the end of the middle block (stable_diffusion_xl/t2i_adapter.py:25-40).  ``residuals``: the 10 ControlLora corrections already sitting
    in the residual slots when the UNet runs (the control copy is child 0 of the UNet, control_lora.py:283-286)."""
    dtype = x.dtype
    temb = sdxl_timestep_embedding(sd, timestep, pooled_text_embedding, time_ids, dtype)
    skips: list[Tensor] = []
    h = x
    for i, entry in enumerate(_SDXL_DOWN):
        p = f"DownBlocks.Chain_{i + 1}"
        if entry[0] == "in":
            h = _conv(sd, p + ".Conv2d", h, padding=1)
        elif entry[0] == "down":
            h = _conv(sd, p + ".Downsample.Conv2d", h, stride=2, padding=1)
        else:
            h = residual_block(sd, p + ".ResidualBlock", h, temb)
            if entry[1]:
                h = cross_attention_2d(sd, p + ".SDXLCrossAttention", h, clip_text_embedding, entry[2], entry[1], True)
        if t2i is not None and i in (3, 5, 8):
            h = h + t2i[1] * t2i[0][(3, 5, 8).index(i)]
        skips.append(h if residuals is None else h + residuals[i])  # ResidualAccumulator is a Passthrough
    m = "MiddleBlock"
    h = residual_block(sd, m + ".ResidualBlock_1", h, temb)
    h = cross_attention_2d(sd, m + ".SDXLCrossAttention", h, clip_text_embedding, 20, 10, True)
    h = residual_block(sd, m + ".ResidualBlock_2", h, temb)
    if t2i is not None:
        h = h + t2i[1] * t2i[0][3]
    # Residual(UseContext residuals[-1]): the spare 10th slot keeps its initial 0.0 unless ControlLora wrote it
    h = h + (0.0 if residuals is None else residuals[9])
    for n, (layers, heads, up) in enumerate(_SDXL_UP):
        p = f"UpBlocks.Chain_{n + 1}"
        h = torch.cat([h, skips[-n - 1]], dim=1)  # ResidualConcatenator(-n-2) on a list with one spare slot
        h = residual_block(sd, p + ".ResidualBlock", h, temb)
        if layers:
            h = cross_attention_2d(sd, p + ".SDXLCrossAttention", h, clip_text_embedding, heads, layers, True)
        if up:
            h = ops.nearest_upsample(h, (h.shape[-2] * 2, h.shape[-1] * 2))
            h = _conv(sd, p + ".Upsample.Conv2d", h, padding=1)
    h = ops.silu(ops.group_norm(h, 32, sd["OutputBlock.GroupNorm.weight"], sd["OutputBlock.GroupNorm.bias"], 1e-5))
    return _conv(sd, "OutputBlock.Conv2d", h, padding=1)
