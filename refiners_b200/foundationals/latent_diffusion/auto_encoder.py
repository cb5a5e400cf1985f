"""The latent-diffusion VAE (the step right after the denoising loop).

Module trees and state-dict keys are those of the reference's
``foundationals/latent_diffusion/auto_encoder.py`` (`Resnet` :41-82, `Encoder` :85-143, `Decoder` :146-207,
`LatentDiffusionAutoencoder` :282-413); ``tests/test_reference_structure.py`` compares them node by node and
``tests/golden/vae.safetensors`` pins encode / decode against the reference's outputs.

Status: ``encode`` / ``decode`` and the PIL helpers run on the host and on the GPU (every conv, GroupNorm(+SiLU) and
residual add through ``librefiners_b200.so``; the single 512-wide attention head of the bottleneck runs on the CUDA-core
flash kernel, which takes head dims up to 512 - it executes once per image, not per denoising step; GPU parity in
tests/test_models_golden.py::test_vae_gpu).

Tiled inference (`FixedGroupNorm` :209-251, blending mask :254-279, tiling :411-621 in the reference) bounds the
activation memory of very large images: the image is processed tile by tile and the overlapping borders are blended
with linear ramps.  A tile must not see its own GroupNorm statistics (the seams would show), so for the duration of
``tiled_inference`` every GroupNorm is wrapped in a ``FixedGroupNorm`` that captures the statistics of ONE pass over a
downscaled copy of the whole image and applies those to every tile.  On CUDA the frozen statistics are an fp32
``[B, groups, 2]`` buffer consumed by ``rb200_group_norm_fixed`` - a frozen tile pass skips the statistics kernels
altogether and reads each activation once.
"""

from __future__ import annotations

from contextlib import contextmanager
from typing import Any, Callable, Iterator, NamedTuple

import torch
import torch.nn.functional as F
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.utils import image_to_tensor, no_grad, tensor_to_image

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


class FixedGroupNorm(fl.Chain, Adapter[fl.GroupNorm]):
    """A GroupNorm whose per-(sample, group) statistics are those of the FIRST tensor it sees."""

    mean: Tensor | None
    var: Tensor | None

    def __init__(self, target: fl.GroupNorm) -> None:
        self.mean = None
        self.var = None
        self._frozen: list[Tensor] = []  # CUDA: the fp32 [B, groups, 2] (mean, rstd) buffer of rb200_group_norm_fixed
        with self.setup_adapter(target):
            super().__init__(fl.Lambda(self.compute_group_norm))

    def compute_group_norm(self, x: Tensor) -> Tensor:
        norm = self.target
        if x.is_cuda:
            y, stats = B.group_norm_fixed(x, norm.num_groups, norm.weight, norm.bias, norm.eps, self._frozen[0] if self._frozen else None)
            if not self._frozen:
                self._frozen.append(stats)
                self.mean = stats[..., 0].flatten()
                self.var = stats[..., 1].flatten().pow(-2) - norm.eps
            return y
        batch, channels, height, width = x.shape
        # one "channel" of an evaluation-mode batch norm per (sample, group) reproduces a group norm with given statistics
        grouped = x.reshape(1, batch * norm.num_groups, channels // norm.num_groups, height, width)
        if self.mean is None or self.var is None:
            self.var, self.mean = torch.var_mean(grouped, dim=(0, 2, 3, 4), correction=0)
        normalised = F.batch_norm(grouped, self.mean, self.var, None, None, False, 0, norm.eps).reshape(x.shape)
        return normalised * norm.weight.reshape(1, -1, 1, 1) + norm.bias.reshape(1, -1, 1, 1)


class _ImageSize(NamedTuple):
    height: int
    width: int


class _Tile(NamedTuple):
    top: int
    left: int
    bottom: int
    right: int


def _create_blending_mask(
    size: _ImageSize, blending: int, num_channels: int, device: Device | None = None, dtype: DType | None = None,
    is_edge: tuple[bool, bool, bool, bool] = (False, False, False, False),
) -> Tensor:
    """Weight of a tile's pixels: 1 inside, a linear 0 -> 1 ramp of ``blending`` pixels towards every border that is
    not an image border (``is_edge`` = top, bottom, left, right), ramps multiplying in the corners."""
    mask = torch.ones(size, device=device, dtype=dtype)
    reach = min(blending, min(size) // 2)
    if blending == 0:
        return mask
    rise = torch.linspace(0, 1, steps=reach, device=device, dtype=dtype)
    fall = rise.flip(0)
    top, bottom, left, right = is_edge
    if not top:
        mask[:reach, :] *= rise[:, None]
    if not bottom:
        mask[-reach:, :] *= fall[:, None]
    if not left:
        mask[:, :reach] *= rise[None, :]
    if not right:
        mask[:, -reach:] *= fall[None, :]
    return mask[None, None].expand(1, num_channels, *size)


class LatentDiffusionAutoencoder(fl.Chain):
    """Encoder + Decoder with the latent scale of the diffusion models (``encoder_scale``)."""

    encoder_scale = 0.18125

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(Encoder(device=device, dtype=dtype), Decoder(device=device, dtype=dtype))
        self._tile_size: _ImageSize | None = None
        self._blending: int | None = None

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

    # -- tiled inference --------------------------------------------------------------------------------------
    @contextmanager
    def tiled_inference(self, image: Any, tile_size: tuple[int, int] = (512, 512), blending: int = 64) -> Iterator[None]:
        """Inside this context ``tiled_image_to_latents`` / ``tiled_latents_to_image`` work on ``tile_size`` =
        (width, height) pixel tiles overlapping by ``blending`` pixels, all normalised with the GroupNorm statistics of
        ``image`` shrunk to one tile.  (The untiled methods keep working but see the frozen statistics too.)"""
        try:
            self._blending = blending
            self._tile_size = _ImageSize(width=tile_size[0], height=tile_size[1])
            self._add_fixed_group_norm(image, inference_size=self._tile_size)
            yield
        finally:
            self._remove_fixed_group_norm()
            self._tile_size = None
            self._blending = None

    def _active_tiling(self) -> tuple[_ImageSize, int]:
        if self._tile_size is None:
            raise ValueError("Tiled inference context manager not active. Use `tiled_inference` method to activate.")
        assert self._blending is not None
        return self._tile_size, self._blending

    def tiled_image_to_latents(self, image: Any) -> Tensor:
        tile_size, blending = self._active_tiling()
        pixels = 2 * image_to_tensor(image, device=self.device, dtype=self.dtype) - 1
        return self._tiled_encode(pixels, tile_size, blending)

    def tiled_latents_to_image(self, x: Tensor) -> Any:
        tile_size, blending = self._active_tiling()
        return tensor_to_image((self._tiled_decode(x, tile_size, blending) + 1) / 2)

    @staticmethod
    def _generate_latent_tiles(size: _ImageSize, tile_size: _ImageSize, overlap: int = 8) -> list[_Tile]:
        """Tiles of (at most) ``tile_size`` covering ``size``, neighbours sharing ``overlap`` rows / columns; column by
        column, top to bottom (the accumulation order of the blend)."""
        lefts = range(0, max(size.width - overlap, 1), tile_size.width - overlap)
        tops = range(0, max(size.height - overlap, 1), tile_size.height - overlap)
        return [
            _Tile(top=top, left=left, bottom=min(size.height, top + tile_size.height), right=min(size.width, left + tile_size.width))
            for left in lefts
            for top in tops
        ]

    @no_grad()
    def _add_fixed_group_norm(self, image: Any, inference_size: _ImageSize) -> None:
        """Wrap every GroupNorm in a ``FixedGroupNorm`` and let them capture their statistics from one encode + decode
        of ``image`` resized to ``inference_size`` - after pulling the resized copy's value range and per-channel mean /
        spread back to those of the full image, which resampling alters."""
        for norm, parent in [*self.walk(fl.GroupNorm)]:
            FixedGroupNorm(norm).inject(parent)
        full = image_to_tensor(image, device=self.device, dtype=self.dtype)
        small = image_to_tensor(image.resize((inference_size.width, inference_size.height)), device=self.device, dtype=self.dtype)
        small.clamp_(min=full.min(), max=full.max())
        spread, centre = torch.std_mean(full, dim=[0, 2, 3], keepdim=True)
        small_spread, small_centre = torch.std_mean(small, dim=[0, 2, 3], keepdim=True)
        small = (small - small_centre) * (spread / small_spread) + centre
        self.decode(self.encode(2 * small - 1))

    def _remove_fixed_group_norm(self) -> None:
        for fixed in [*self.layers(FixedGroupNorm)]:
            fixed.eject()

    def _blend_tiles(
        self, run: Callable[[Tensor], Tensor], source: Tensor, latent_size: _ImageSize, tile_size: _ImageSize, blending: int,
        source_scale: int, result_scale: int, result_channels: int,
    ) -> Tensor:
        """``run`` over every tile of ``source`` (tiles are laid out on the latent grid; ``source`` / the result have
        ``source_scale`` / ``result_scale`` pixels per latent), accumulated with the blending windows and normalised by
        the accumulated window weight."""
        tiles = self._generate_latent_tiles(
            latent_size, tile_size=_ImageSize(height=tile_size.height // 8, width=tile_size.width // 8), overlap=blending // 8
        )
        if len(tiles) == 1:
            return run(source)
        shape = (1, result_channels, latent_size.height * result_scale, latent_size.width * result_scale)
        total = torch.zeros(shape, device=self.device, dtype=self.dtype)
        weight = torch.zeros_like(total)
        for tile in tiles:
            k = source_scale
            piece = run(source[:, :, tile.top * k : tile.bottom * k, tile.left * k : tile.right * k])
            k = result_scale
            window = _create_blending_mask(
                _ImageSize(height=(tile.bottom - tile.top) * k, width=(tile.right - tile.left) * k),
                blending * k // 8, num_channels=result_channels, device=self.device, dtype=self.dtype,
                is_edge=(tile.top == 0, tile.bottom == latent_size.height, tile.left == 0, tile.right == latent_size.width),
            )
            where = (slice(None), slice(None), slice(tile.top * k, tile.bottom * k), slice(tile.left * k, tile.right * k))
            total[where] += piece * window
            weight[where] += window
        return total / weight

    @no_grad()
    def _tiled_encode(self, image_tensor: Tensor, tile_size: _ImageSize, blending: int = 64) -> Tensor:
        latent_size = _ImageSize(height=image_tensor.shape[2] // 8, width=image_tensor.shape[3] // 8)
        return self._blend_tiles(self.encode, image_tensor, latent_size, tile_size, blending, source_scale=8, result_scale=1, result_channels=4)

    @no_grad()
    def _tiled_decode(self, latents: Tensor, tile_size: _ImageSize, blending: int = 64) -> Tensor:
        latent_size = _ImageSize(height=latents.shape[2], width=latents.shape[3])
        return self._blend_tiles(self.decode, latents, latent_size, tile_size, blending, source_scale=1, result_scale=8, result_channels=3)
