"""Small tensor helpers used by the fluxion mirror.

Covers the subset of /root/reference/src/refiners/fluxion/utils.py that the hot
path touches (seed, no_grad, pad, interpolate, tensor summary for ChainError,
safetensors IO, PIL <-> tensor, the Gaussian blur of self-attention guidance).
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


def gaussian_blur(
    tensor: Tensor, kernel_size: int | tuple[int, int], sigma: float | tuple[float, float] | None = None
) -> Tensor:
    """Depthwise Gaussian blur of the last two axes with reflected borders (reference utils.py:65-113, the
    torchvision recipe): normalised 1-D taps on ``[-(k-1)/2, (k-1)/2]``, their outer product as ONE 2-D kernel per
    channel, ``sigma`` defaulting to ``0.15 k + 0.35``.  ``kernel_size`` / ``sigma`` pairs are (x, y)."""
    assert torch.is_floating_point(tensor)
    sizes = (kernel_size, kernel_size) if isinstance(kernel_size, int) else tuple(kernel_size)
    if sigma is None:
        sigmas: tuple[float, ...] = tuple(0.15 * k + 0.35 for k in sizes)
    elif isinstance(sigma, float):
        sigmas = (sigma, sigma)
    else:
        assert isinstance(sigma, tuple)
        sigmas = sigma

    def taps(k: int, s: float) -> Tensor:
        reach = 0.5 * (k - 1)
        grid = torch.linspace(-reach, reach, steps=k, device=tensor.device, dtype=tensor.dtype)
        bell = torch.exp(-0.5 * (grid / s).pow(2))
        return bell / bell.sum()

    (kx, ky), (sx, sy) = sizes, sigmas
    window = torch.mm(taps(ky, sy)[:, None], taps(kx, sx)[None, :])
    channels = tensor.shape[-3]
    framed = pad(tensor, (kx // 2, kx // 2, ky // 2, ky // 2), mode="reflect")
    return F.conv2d(framed, weight=window.expand(channels, 1, ky, kx), groups=channels)


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
    # a pickle: only plain tensors are unpickled (anything else raises UnpicklingError), and what comes back must be a
    # flat name -> tensor mapping (contract: the reference's tests/fluxion/test_utils.py:102-125)
    loaded = torch.load(path, map_location=device, weights_only=True)
    assert isinstance(loaded, dict), f"{path}: expected a mapping of tensors, found {type(loaded).__name__}"
    stray = [key for key, value in loaded.items() if not (isinstance(key, str) and isinstance(value, Tensor))]
    assert not stray, f"{path}: entries that are not str -> Tensor: {stray[:5]}"
    return loaded


# ---------------------------------------------------------------------------- PIL <-> tensors
# (utils.py:116-200 in the reference: values in [0, 1], batch axis first, channels by PIL mode)
def image_to_tensor(image: Any, device: torch.device | str | None = None, dtype: torch.dtype | None = None) -> Tensor:
    """PIL image -> ``[1, C, H, W]`` in [0, 1]; C = 1 (mode L), 3 (RGB) or 4 (RGBA)."""
    import numpy as np

    pixels = torch.tensor(np.array(image).astype(np.float32) / 255.0, device=device, dtype=dtype)
    if image.mode == "L":
        pixels = pixels.unsqueeze(0)
    elif image.mode in ("RGB", "RGBA"):
        pixels = pixels.permute(2, 0, 1)
    else:
        raise ValueError(f"Unsupported image mode: {image.mode}")
    return pixels.unsqueeze(0)


def images_to_tensor(images: list[Any], device: torch.device | str | None = None, dtype: torch.dtype | None = None) -> Tensor:
    return torch.cat([image_to_tensor(image, device=device, dtype=dtype) for image in images])


def tensor_to_image(tensor: Tensor) -> Any:
    """``[1, C, H, W]`` (clamped to [0, 1]) -> PIL image of mode L / RGB / RGBA."""
    from PIL import Image

    assert tensor.ndim == 4 and tensor.shape[0] == 1, f"Unsupported tensor shape: {tensor.shape}"
    channels = tensor.shape[1]
    data = tensor.clamp(0, 1).squeeze(0).to(torch.float32)  # numpy has no bfloat16
    if channels == 1:
        data = data.squeeze(0)
    elif channels in (3, 4):
        data = data.permute(1, 2, 0)
    else:
        raise ValueError(f"Unsupported number of channels: {channels}")
    return Image.fromarray((data.cpu().numpy() * 255).astype("uint8"))


def tensor_to_images(tensor: Tensor) -> list[Any]:
    return [tensor_to_image(t) for t in tensor.split(1)]
