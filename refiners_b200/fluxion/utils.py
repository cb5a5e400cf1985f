"""Small tensor helpers used by the fluxion mirror.

Covers the subset of /root/reference/src/refiners/fluxion/utils.py that the hot
path touches (seed, no_grad, pad, interpolate, tensor summary for ChainError,
safetensors IO). Image helpers live outside the denoising path and are not mirrored.
"""

from pathlib import Path
from typing import Any, Iterable

import torch
from torch import Tensor
from torch.nn import functional as F


def manual_seed(seed: int) -> None:
    torch.manual_seed(seed)


class no_grad(torch.no_grad):
    def __new__(cls, orig_func: Any | None = None) -> "no_grad":
        return object.__new__(cls)


def norm(x: Tensor) -> Tensor:
    return torch.linalg.vector_norm(x)


def pad(x: Tensor, pad: Iterable[int], value: float = 0.0, mode: str = "constant") -> Tensor:
    return F.pad(x, tuple(pad), mode=mode, value=value)


def interpolate(x: Tensor, size: torch.Size, mode: str = "nearest", antialias: bool = False) -> Tensor:
    return F.interpolate(x, size=tuple(size), mode=mode, antialias=antialias)


def summarize_tensor(tensor: Tensor, /) -> str:
    """One-line description used in ChainError messages (utils.py:235-279 in the reference)."""
    parts = [
        f"shape=({', '.join(str(d) for d in tensor.shape)})",
        f"dtype={str(tensor.dtype).removeprefix('torch.')}",
        f"device={tensor.device}",
    ]
    if tensor.numel() and not tensor.is_meta:
        try:
            if tensor.is_complex():
                t = tensor.abs().float()
            else:
                t = tensor.float()
            parts += [
                f"min={t.min().item():.2f}",
                f"max={t.max().item():.2f}",
                f"mean={t.mean().item():.2f}",
                f"std={t.std().item():.2f}" if t.numel() > 1 else "std=nan",
                f"norm={norm(t).item():.2f}",
                f"grad={tensor.requires_grad}",
            ]
        except Exception:  # a failed device op must not mask the original error
            parts.append("stats=unavailable")
    return "Tensor(" + ", ".join(parts) + ")"


def load_from_safetensors(path: Path | str, device: torch.device | str = "cpu") -> dict[str, Tensor]:
    from safetensors.torch import load_file

    return load_file(str(path), device=str(device))


def save_to_safetensors(path: Path | str, tensors: dict[str, Tensor], metadata: dict[str, str] | None = None) -> None:
    from safetensors.torch import save_file

    save_file(tensors, str(path), metadata)


def load_tensors(path: Path | str, /, device: torch.device | str = "cpu") -> dict[str, Tensor]:
    if str(path).endswith(".safetensors"):
        return load_from_safetensors(path, device=device)
    return torch.load(path, map_location=device, weights_only=True)
