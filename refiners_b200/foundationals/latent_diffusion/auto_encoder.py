"""The latent-diffusion VAE (the step right after the denoising loop).

Module trees and state-dict keys are those of the reference's
``foundationals/latent_diffusion/auto_encoder.py`` (`Resnet` :41-82, `Encoder` :85-143, `Decoder` :146-207,
`LatentDiffusionAutoencoder` :282-413); ``tests/test_reference_structure.py`` compares them node by node and
``tests/golden/vae.safetensors`` pins encode / decode against the reference's outputs.

Status: ``encode`` / ``decode`` and the PIL helpers run on the host and on the GPU (every conv, GroupNorm(+SiLU) and
residual add through ``librefiners_b200.so``; the single 512-wide attention head of the bottleneck runs on the CUDA-core
flash kernel, which takes head dims up to 512 - it executes once per image, not per denoising step; GPU parity in
tests/test_models_golden.py::test_vae_gpu).  Tiled inference (:415-621 in the reference) is not built.
"""

from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.context import Contexts

Device = torch.device
DType = torch.dtype

_WIDTHS = (128, 256, 512, 512, 512)  # channels per resolution level, full resolution first


class Resnet(fl.Sum):
    """shortcut(x) + conv(SiLU(GN(conv(SiLU(GN(x)))))); the shortcut is a 1x1 conv when the width changes."""

    def __init__(
        self, in_channels: int, out_channels: int, num_groups: int = 32, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.in_channels, self.out_channels = in_channels, out_channels
        kw = dict(device=device, dtype=dtype)

        def stage(cin: int) -> list[fl.Module]:
            return [fl.GroupNorm(channels=cin, num_groups=num_groups, **kw), fl.SiLU(),
                    fl.Conv2d(cin, out_channels, kernel_size=3, padding=1, **kw)]

        skip = fl.Identity() if in_channels == out_channels else fl.Conv2d(in_channels, out_channels, kernel_size=1, **kw)
        super().__init__(skip, fl.Chain(*stage(in_channels), *stage(out_channels)))


def _bottleneck_attention(channels: int, kw: dict) -> fl.Residual:
    """x + SelfAttention2d(GN(x)): one head over all pixels of the lowest-resolution map."""
    return fl.Residual(
        fl.GroupNorm(channels=channels, num_groups=32, eps=1e-6, **kw),
        fl.SelfAttention2d(channels=channels, **kw),
    )


def _head(cin: int, cout: int, kw: dict) -> fl.Chain:
    return fl.Chain(fl.GroupNorm(channels=cin, num_groups=32, eps=1e-6, **kw), fl.SiLU(),
                    fl.Conv2d(cin, cout, kernel_size=3, padding=1, **kw))


class Encoder(fl.Chain):
    """``[B, 3, H, W]`` in [-1, 1] -> ``[B, 4, H/8, W/8]`` (the mean half of the 8-channel moments)."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        kw = dict(device=device, dtype=dtype)
        levels: list[fl.Chain] = []
        for i, width in enumerate(_WIDTHS):
            previous = _WIDTHS[i - 1] if i else width
            levels.append(fl.Chain(Resnet(previous, width, **kw), Resnet(width, width, **kw)))
        for level in levels[:3]:  # three 2x reductions
            level.append(fl.Downsample(channels=level[-1].out_channels, scale_factor=2, **kw))
        levels[-1].insert_after_type(Resnet, _bottleneck_attention(_WIDTHS[-1], kw))
        super().__init__(
            fl.Conv2d(3, _WIDTHS[0], kernel_size=3, padding=1, **kw),
            fl.Chain(*levels),
            _head(_WIDTHS[-1], 8, kw),
            fl.Chain(fl.Conv2d(8, 8, kernel_size=1, **kw), fl.Slicing(dim=1, end=4)),
        )

    def init_context(self) -> Contexts:
        return {"sampling": {"shapes": []}}


class Decoder(fl.Chain):
    """``[B, 4, h, w]`` latents -> ``[B, 3, 8h, 8w]`` image in [-1, 1]."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.resnet_sizes: list[int] = list(_WIDTHS)
        self.latent_dim: int = 4
        self.output_channels: int = 3
        kw = dict(device=device, dtype=dtype)
        widths = _WIDTHS[::-1]  # lowest resolution first
        levels: list[fl.Chain] = []
        for i, width in enumerate(widths):
            previous = widths[i - 1] if i else width
            blocks = [Resnet(previous, width, **kw), Resnet(width, width, **kw)]
            if i:
                blocks.append(Resnet(width, width, **kw))
            levels.append(fl.Chain(*blocks))
        levels[0].insert(1, _bottleneck_attention(widths[0], kw))
        for level in levels[1:4]:  # three 2x upsamplings; Chain.insert(-1, ...) appends, so each level ENDS with its Upsample
            level.insert(-1, fl.Upsample(channels=level.layer(-1, Resnet).out_channels, upsample_factor=2, **kw))
        super().__init__(
            fl.Conv2d(self.latent_dim, self.latent_dim, kernel_size=1, **kw),
            fl.Conv2d(self.latent_dim, widths[0], kernel_size=3, padding=1, **kw),
            fl.Chain(*levels),
            _head(widths[-1], self.output_channels, kw),
        )


class LatentDiffusionAutoencoder(fl.Chain):
    """Encoder + Decoder with the latent scale of the diffusion models (``encoder_scale``)."""

    encoder_scale = 0.18125

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(Encoder(device=device, dtype=dtype), Decoder(device=device, dtype=dtype))

    def encode(self, x: Tensor) -> Tensor:
        return self.encoder_scale * self[0](x)

    def decode(self, x: Tensor) -> Tensor:
        return self[1](x / self.encoder_scale)

    # -- PIL helpers (values in [0, 1] on the image side, [-1, 1] on the VAE side)
    def images_to_latents(self, images: list[Any]) -> Tensor:
        from refiners_b200.fluxion.utils import images_to_tensor

        return self.encode(2 * images_to_tensor(images, device=self.device, dtype=self.dtype) - 1)

    def image_to_latents(self, image: Any) -> Tensor:
        return self.images_to_latents([image])

    def latents_to_images(self, x: Tensor) -> list[Any]:
        from refiners_b200.fluxion.utils import tensor_to_images

        return tensor_to_images((self.decode(x) + 1) / 2)

    def latents_to_image(self, x: Tensor) -> Any:
        if x.shape[0] != 1:
            raise ValueError(f"Expected batch size of 1, got {x.shape[0]}")
        return self.latents_to_images(x)[0]

    def tiled_inference(self, *args: Any, **kwargs: Any) -> Any:
        raise NotImplementedError("tiled VAE inference is not built in refiners_b200 yet (auto_encoder.py:415-621 in the reference)")
