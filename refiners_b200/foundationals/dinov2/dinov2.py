"""The published DINOv2 sizes (arXiv:2304.07193; `_reg` = with 4 register tokens, arXiv:2309.16588).

Configurations as in the reference's ``foundationals/dinov2/dinov2.py:26-336``: patch 14, pre-training grid 518/14 = 37,
head dim 64 everywhere; the giant models use a SwiGLU MLP of hidden width 4096."""

from __future__ import annotations

from typing import Any

import torch

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.dinov2.vit import ViT

Device = torch.device
DType = torch.dtype

#            embedding_dim, num_layers, num_heads
_SIZES = {"small": (384, 12, 6), "base": (768, 12, 12), "large": (1024, 24, 16), "giant": (1536, 40, 24)}


def _config(size: str, registers: bool) -> dict[str, Any]:
    width, depth, heads = _SIZES[size]
    cfg: dict[str, Any] = dict(embedding_dim=width, patch_size=14, image_size=518, num_layers=depth, num_heads=heads)
    if size == "giant":
        cfg.update(feedforward_dim=4096, activation=fl.GLU(fl.SiLU()))
    if registers:
        cfg.update(num_registers=4, interpolate_antialias=True)
    return cfg


def _variant(name: str, size: str, registers: bool) -> type[ViT]:
    def __init__(self: ViT, device: Device | str | None = None, dtype: DType | None = None) -> None:
        ViT.__init__(self, **_config(size, registers), device=device, dtype=dtype)

    return type(name, (ViT,), {"__init__": __init__, "__doc__": f"DINOv2 {size}{' with registers' if registers else ''}.",
                               "__module__": __name__})


DINOv2_small = _variant("DINOv2_small", "small", False)
DINOv2_base = _variant("DINOv2_base", "base", False)
DINOv2_large = _variant("DINOv2_large", "large", False)
DINOv2_giant = _variant("DINOv2_giant", "giant", False)
DINOv2_small_reg = _variant("DINOv2_small_reg", "small", True)
DINOv2_base_reg = _variant("DINOv2_base_reg", "base", True)
DINOv2_large_reg = _variant("DINOv2_large_reg", "large", True)
DINOv2_giant_reg = _variant("DINOv2_giant_reg", "giant", True)


def preprocess(img: Any, dim: int = 224) -> torch.Tensor:
    """Resize to ``dim`` x ``dim`` (no crop) and normalise with the ImageNet statistics: fp32 ``[3, dim, dim]``."""
    from refiners_b200.fluxion.utils import image_to_tensor

    t = image_to_tensor(img.convert("RGB").resize((dim, dim))).squeeze()
    mean = torch.tensor([0.485, 0.456, 0.406]).reshape(3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).reshape(3, 1, 1)
    return (t - mean) / std
