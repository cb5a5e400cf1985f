"""BASELINE-size parity cases (128x128 latents = 1024^2 pixels; SAM at 1024^2) shared by the pin script
and the tests.  TEST INFRASTRUCTURE - see oracle/__init__.py.

Every case is described once, against an API namespace, so that the SAME construction runs on the real
reference (``refiners.*``, in oracle/pin_against_reference.py - where the fixtures are recorded) and on the
refiners_b200 mirror (in tests/).  Inputs and adapter weights are *keyed*: drawn from a numpy PCG64 stream
seeded by crc32(name), so nothing but the reference's outputs has to be stored (a 1024^2 condition image
alone would be 25 MB).

Cases:
  config 2   SDXLUNet, UNet batch 2, keyed weights (seed 2)
  config 3   + two rank-16 LinearLoras (scales 1.0 / 1.4) on every Linear under a CrossAttentionBlock
             (700 adapters, as unet_inclusions=["CrossAttentionBlock"], tests/e2e/test_diffusion.py:1616 of
             the reference) + SDXLIPAdapter sub-adapters with 4 image tokens (image_prompt.py:237-455)
  config 4   + ControlLoraAdapter("canny") with rank-8 LoRAs on a spread of shared leaves, keyed
             ConditionEncoder / ZeroConvolution weights, 3 x 1024 x 1024 condition
  config 5   SAMViTH image encoder on one 1024^2 image
  step       StableDiffusion_XL(x, step=...) with CFG + Euler (A17) at latent batch 1
"""

from __future__ import annotations

import zlib
from types import SimpleNamespace
from typing import Any

import numpy as np
import torch
from torch import Tensor

from oracle.weights import keyed_state_dict, keyed_tensor

SQRT3 = float(np.sqrt(3.0))


def keyed_input(name: str, shape: tuple[int, ...], scale: float = 1.0) -> Tensor:
    """Unit-variance uniform values in [-sqrt 3, sqrt 3) * scale, a pure function of (name, shape)."""
    rng = np.random.Generator(np.random.PCG64(zlib.crc32(("input:" + name).encode())))
    u = (rng.random(size=shape, dtype=np.float32) * 2.0 - 1.0) * np.float32(SQRT3 * scale)
    return torch.from_numpy(np.ascontiguousarray(u))


def reference_api() -> SimpleNamespace:
    """The reference's own modules (needs /root/reference on sys.path, see pin_against_reference.py)."""
    import refiners.fluxion.layers as fl
    from refiners.fluxion.adapters.lora import LinearLora, LoraAdapter
    from refiners.foundationals.clip.image_encoder import CLIPImageEncoderH
    from refiners.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock
    from refiners.foundationals.latent_diffusion.solvers import Euler
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import (
        ConditionEncoder,
        ControlLoraAdapter,
        ZeroConvolution,
    )
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.image_prompt import SDXLIPAdapter
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.model import StableDiffusion_XL
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet
    from refiners.foundationals.segment_anything.image_encoder import SAMViTH

    return SimpleNamespace(**{k: v for k, v in locals().items()})


def engine_api() -> SimpleNamespace:
    """The refiners_b200 mirror of the same names."""
    import refiners_b200.fluxion.layers as fl
    from refiners_b200.fluxion.adapters import LinearLora, LoraAdapter
    from refiners_b200.foundationals.clip.image_encoder import CLIPImageEncoderH
    from refiners_b200.foundationals.latent_diffusion import Euler, SDXLUNet, StableDiffusion_XL
    from refiners_b200.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock
    from refiners_b200.foundationals.latent_diffusion.image_prompt import SDXLIPAdapter
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import (
        ConditionEncoder,
        ControlLoraAdapter,
        ZeroConvolution,
    )
    from refiners_b200.foundationals.segment_anything import SAMViTH

    return SimpleNamespace(**{k: v for k, v in locals().items()})


# ------------------------------------------------------------------------------------- inputs
def sdxl_inputs(tag: str, batch: int, latent: int = 128) -> dict[str, Tensor]:
    return {
        "x": keyed_input(f"{tag}.x", (batch, 4, latent, latent)),
        "ctx": keyed_input(f"{tag}.ctx", (batch, 77, 2048)),
        "pooled": keyed_input(f"{tag}.pooled", (batch, 1280)),
        "time_ids": torch.tensor([[1024, 1024, 0, 0, 1024, 1024]]).repeat(batch, 1),
        "timestep": torch.tensor([981.0]),
    }


def set_sdxl_contexts(unet: Any, inp: dict[str, Tensor], device: Any = "cpu", dtype: torch.dtype = torch.float32) -> None:
    unet.set_timestep(inp["timestep"].to(device))
    unet.set_clip_text_embedding(inp["ctx"].to(device, dtype))
    unet.set_pooled_text_embedding(inp["pooled"].to(device, dtype))
    unet.set_time_ids(inp["time_ids"].to(device))


def sdxl_base_weights(api: SimpleNamespace, seed: int = 2) -> dict[str, Tensor]:
    shapes = {k: tuple(v.shape) for k, v in api.SDXLUNet(4, device="meta").state_dict().items()}
    return keyed_state_dict(shapes, seed=seed)


def build_sdxl(api: SimpleNamespace, weights: dict[str, Tensor], device: Any = "cpu", dtype: torch.dtype = torch.float32) -> Any:
    unet = api.SDXLUNet(4, device="meta")
    unet.load_state_dict({k: v.to(device, dtype) for k, v in weights.items()}, assign=True)
    return unet


def module_paths(root: Any) -> dict[int, str]:
    """id(module) -> dotted path below ``root`` = the module's state-dict prefix."""
    return {id(m): name for name, m in root.named_modules()}


# ----------------------------------------------------------------------------------- config 3
LORA_RANK = 16
LORA_SCALES = (1.0, 1.4)
IP_SCALE = 0.6
IP_TOKENS = 4


def lora_factors(path: str, j: int, in_features: int, out_features: int, rank: int = LORA_RANK) -> tuple[Tensor, Tensor]:
    """down ~ unit-variance preserving, up at 0.2 of that: a visible but not dominating LoRA term."""
    down = keyed_tensor(f"lora{j}.{path}.down", (rank, in_features))
    up = keyed_tensor(f"lora{j}.{path}.up", (out_features, rank)) * 0.2
    return down, up


def ip_factors(path: str, inner_dim: int, ctx_dim: int = 2048) -> tuple[Tensor, Tensor]:
    return keyed_tensor(f"ip.{path}.k", (inner_dim, ctx_dim)), keyed_tensor(f"ip.{path}.v", (inner_dim, ctx_dim))


def attach_config3(api: SimpleNamespace, unet: Any, batch: int, device: Any = "cpu", dtype: torch.dtype = torch.float32) -> tuple[Any, dict]:
    """Inject the config-3 adapters into ``unet`` (weights already loaded).  Returns the IP adapter and the
    oracle-side description {loras, ip, ip_scale, ip_embedding} (fp32 CPU tensors)."""
    fl = api.fl
    paths = module_paths(unet)
    targets = []
    for lin, parent in unet.walk(fl.Linear, recurse=True):
        if any(isinstance(a, api.CrossAttentionBlock) for a in [*parent.get_parents(), parent]):
            targets.append((lin, parent, paths[id(lin)]))
    cross = {id(a): paths[id(a)] for a in unet.layers(fl.Attention) if type(a) is not fl.SelfAttention}
    loras: dict[str, list[tuple[Tensor, Tensor, float]]] = {}
    for lin, parent, path in targets:
        mods = []
        for j, scale in enumerate(LORA_SCALES):
            down, up = lora_factors(path, j, lin.in_features, lin.out_features)
            lora = api.LinearLora(f"lora{j}", in_features=lin.in_features, out_features=lin.out_features, rank=LORA_RANK, scale=scale,
                                  device=device, dtype=dtype)
            lora.down.weight.data.copy_(down.to(device, dtype))
            lora.up.weight.data.copy_(up.to(device, dtype))
            mods.append(lora)
            loras.setdefault(path, []).append((down, up, scale))
        api.LoraAdapter(lin, *mods).inject(parent)
    ip = api.SDXLIPAdapter(unet, clip_image_encoder=api.CLIPImageEncoderH(device="meta"), scale=IP_SCALE)
    ip_w: dict[str, tuple[Tensor, Tensor]] = {}
    for sub in ip.sub_adapters:
        path = cross[id(sub.target)]
        wk, wv = ip_factors(path, sub.target.inner_dim)
        sub.load_weights(wk.to(device, dtype), wv.to(device, dtype))
        ip_w[path] = (wk, wv)
    ip.inject()
    emb = keyed_input("cfg3.clip_image_embedding", (batch, IP_TOKENS, 2048))
    ip.set_clip_image_embedding(emb.to(device, dtype))
    return ip, {"loras": loras, "ip": ip_w, "ip_scale": IP_SCALE, "ip_embedding": emb, "n_lora_adapters": len(targets)}


# ----------------------------------------------------------------------------------- config 4
CL_RANK = 8
CL_SCALE = 0.8


def attach_config4(api: SimpleNamespace, unet: Any, batch: int, device: Any = "cpu", dtype: torch.dtype = torch.float32,
                   cond_size: int = 1024) -> tuple[Any, dict]:
    """ControlLoraAdapter("canny") with keyed own weights and rank-8 LoRAs on every 7th Linear / Conv-free
    leaf of the control copy.  Returns the adapter and the oracle-side description."""
    fl = api.fl
    adapter = api.ControlLoraAdapter("canny", unet, scale=CL_SCALE)
    cl = adapter.control_lora
    # the copy's own parameters (fresh leaves): keyed by their path inside the copy
    own: dict[str, Tensor] = {}
    for owner in [*cl.layers(api.ConditionEncoder), *cl.layers(api.ZeroConvolution)]:
        base = module_paths(cl)[id(owner)]
        for k, prm in owner.state_dict(keep_vars=True).items():
            value = keyed_tensor(f"cl.{base}.{k}", tuple(prm.shape), seed=4)
            if "ZeroConvolution" in base and k.endswith(".weight"):
                value = value * 0.5
            own[f"{base}.{k}"] = value
        owner.load_state_dict({k: own[f"{base}.{k}"].to(device, dtype) for k in owner.state_dict()}, assign=True)
    paths = module_paths(cl)
    linears = [(lin, paths[id(lin)]) for lin, _ in cl.walk(fl.Linear, recurse=True)]
    sd: dict[str, Tensor] = {}
    loras: dict[str, list[tuple[Tensor, Tensor, float]]] = {}
    for lin, path in linears[::7]:
        down = keyed_tensor(f"cl.lora.{path}.down", (CL_RANK, lin.in_features))
        up = keyed_tensor(f"cl.lora.{path}.up", (lin.out_features, CL_RANK)) * 0.2
        sd[f"ControlLora.{path}.down"], sd[f"ControlLora.{path}.up"] = down, up
        loras[path] = [(down, up, 1.0)]
    api.ControlLoraAdapter.load_lora_layers("canny", {k: v.to(device, dtype) for k, v in sd.items()}, cl)
    adapter.inject()
    cond = keyed_input("cfg4.condition", (batch, 3, cond_size, cond_size)).abs().clamp(max=1.0)
    adapter.set_condition(cond.to(device, dtype))
    return adapter, {"own": own, "loras": loras, "scale": CL_SCALE, "condition": cond, "n_loras": len(loras)}


# ----------------------------------------------------------------------------------- config 5
def sam_inputs() -> Tensor:
    return keyed_input("cfg5.image", (1, 3, 1024, 1024))


def build_sam(api: SimpleNamespace, device: Any = "cpu", dtype: torch.dtype = torch.float32) -> tuple[Any, dict[str, Tensor]]:
    sam = api.SAMViTH(device="meta")
    shapes = {k: tuple(v.shape) for k, v in sam.state_dict().items()}
    sd = keyed_state_dict(shapes, seed=8)
    for k in sd:  # relative-position tables and the learned positions: small, as in a trained model
        if "embedding" in k or k.startswith("PositionalEncoder"):
            sd[k] = sd[k] * 0.3
    sam.load_state_dict({k: v.to(device, dtype) for k, v in sd.items()}, assign=True)
    return sam, sd


# --------------------------------------------------------------------------------------- step
STEP_CASES = ((0, 5.0), (13, 7.5), (29, 5.0))  # (step, condition_scale)


def step_inputs() -> dict[str, Tensor]:
    return {
        "x": keyed_input("step.x", (1, 4, 128, 128)),  # multiplied by init_noise_sigma by the caller
        "ctx": keyed_input("step.ctx", (2, 77, 2048)),
        "pooled": keyed_input("step.pooled", (2, 1280)),
        "time_ids": torch.tensor([[1024, 1024, 0, 0, 1024, 1024]]).repeat(2, 1),
    }
