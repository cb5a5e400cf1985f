"""What sits either side of the SAM image encoder: an image of any size goes in as a 1024 x 1024 tensor (longest side
scaled to the encoder resolution, ImageNet-normalised in 0..255 units, zero-padded at the bottom / right), and masks or
point prompts are mapped between the original image's pixel grid and that frame.

Function names, argument order and results are the contract of
/root/reference/src/refiners/foundationals/segment_anything/utils.py:7-130 (the reference's own
tests/foundationals/segment_anything/test_utils.py runs against this file: tests/test_reference_own_tests.py).
Host-side, once per image: none of this is on the timed path."""

from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

from refiners_b200.fluxion.utils import image_to_tensor, interpolate, pad

Device = torch.device
DType = torch.dtype

# ImageNet statistics in 0..255 units, the scale SAM's encoder was trained on
PIXEL_MEAN = (123.675, 116.28, 103.53)
PIXEL_STD = (58.395, 57.12, 57.375)


def compute_scaled_size(size: tuple[int, int], image_encoder_resolution: int) -> tuple[int, int]:
    """(h, w) of the image once its longest side equals the encoder resolution (aspect ratio kept, halves round up)."""
    factor = image_encoder_resolution / max(size)
    height, width = (int(side * factor + 0.5) for side in size)
    return height, width


def image_to_scaled_tensor(
    image: Any, scaled_size: tuple[int, int], device: Device | None = None, dtype: DType | None = None
) -> Tensor:
    """PIL image -> ``[1, C, h, w]`` in 0..255 (bilinear resize to ``scaled_size`` = (h, w))."""
    from PIL import Image

    height, width = scaled_size
    shrunk = image.resize((width, height), resample=Image.Resampling.BILINEAR)
    return 255.0 * image_to_tensor(shrunk, device=device, dtype=dtype)


def pad_image_tensor(image_tensor: Tensor, scaled_size: tuple[int, int], image_encoder_resolution: int) -> Tensor:
    """Zero rows below and zero columns to the right, up to the encoder's square input."""
    assert image_tensor.ndim == 4, f"expected [B, C, H, W], got {tuple(image_tensor.shape)}"
    assert max(image_tensor.shape[-2:]) <= image_encoder_resolution, "the image is larger than the encoder's input"
    height, width = scaled_size
    return pad(image_tensor, (0, image_encoder_resolution - width, 0, image_encoder_resolution - height))


def preprocess_image(
    image: Any, image_encoder_resolution: int, device: Device | None = None, dtype: DType | None = None
) -> Tensor:
    """PIL image -> the ``[1, 3, R, R]`` tensor `SAMViT` consumes."""
    scaled_size = compute_scaled_size((image.height, image.width), image_encoder_resolution)
    pixels = image_to_scaled_tensor(image, scaled_size, device=device, dtype=dtype)
    mean = torch.tensor(PIXEL_MEAN, device=pixels.device, dtype=pixels.dtype).view(1, -1, 1, 1)
    std = torch.tensor(PIXEL_STD, device=pixels.device, dtype=pixels.dtype).view(1, -1, 1, 1)
    return pad_image_tensor((pixels - mean) / std, scaled_size, image_encoder_resolution)


def postprocess_masks(low_res_masks: Tensor, original_size: tuple[int, int], image_encoder_resolution: int) -> Tensor:
    """Low-resolution mask logits -> the original image's (h, w): up to the encoder frame, padding cut off, then to size."""
    height, width = compute_scaled_size(original_size, image_encoder_resolution)
    frame = interpolate(low_res_masks, size=torch.Size((image_encoder_resolution, image_encoder_resolution)), mode="bilinear")
    return interpolate(frame[..., :height, :width], size=torch.Size(original_size), mode="bilinear")


def normalize_coordinates(coordinates: Tensor, original_size: tuple[int, int], image_encoder_resolution: int) -> Tensor:
    """Pixel coordinates ``[..., (x, y)]`` of the original image -> [0, 1] in the encoder frame (pixel centres; IN PLACE,
    like the reference)."""
    height, width = compute_scaled_size(original_size, image_encoder_resolution)
    for axis, (scaled, original) in enumerate(((width, original_size[1]), (height, original_size[0]))):
        coordinates[:, :, axis] = (coordinates[:, :, axis] * (scaled / original) + 0.5) / image_encoder_resolution
    return coordinates
