"""SAM ViT image encoder (ViT-H by default): [B, 3, 1024, 1024] image -> [B, 256, 64, 64] embedding.

Contract - class names, constructor signatures, children (= state-dict keys and ``repr``), the two relative-position
tables and the addition order of the two bias terms - from
/root/reference/src/refiners/foundationals/segment_anything/image_encoder.py: `PatchEncoder` :9-34,
`PositionalEncoder` :37-55, `RelativePositionAttention` :58-143, `FusedSelfAttention` :146-190, `FeedForward`,
`WindowPartition` / `WindowMerge` :206-236, `TransformerLayer` :239-283, `Neck` :286-310, `SAMViT` / `SAMViTH` :316-368.

Execution.  On CUDA the attention is one call into the C ABI (rb200_sam_attention): the bias tables
``q . R[offset]`` are computed for the 2S-1 distinct offsets of each axis and consumed inside a tcgen05 flash kernel;
the [B x heads, HW, HW] logits (512 MB per image in the global layers) never exist, and window partition / merge are
one launch each.  The host path below evaluates the same decomposition with dense tensor algebra: it first projects
the queries on all 2S-1 table rows (one matmul per axis) and then *gathers* the Toeplitz structure, instead of
gathering a [S, S, d] table and contracting it as the reference does - same dot products, S/2 times fewer of them.
"""

from __future__ import annotations

from typing import Any

import torch
from torch import Tensor, nn

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.utils import pad

Device = torch.device
DType = torch.dtype


def _placement(device: Device | str | None, dtype: DType | None) -> dict[str, Any]:
    return {"device": device, "dtype": dtype}


class PatchEncoder(fl.Chain):
    """Non-overlapping patches: a convolution with kernel = stride = patch size, then NCHW -> NHWC."""

    def __init__(
        self, in_channels: int, out_channels: int, patch_size: int = 16, use_bias: bool = True,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.patch_size = patch_size
        self.use_bias = use_bias
        square = (patch_size, patch_size)
        super().__init__(
            fl.Conv2d(in_channels, out_channels, kernel_size=square, stride=square, use_bias=use_bias, **_placement(device, dtype)),
            fl.Permute(0, 2, 3, 1),
        )


class PositionalEncoder(fl.Residual):
    """Adds a learned [H, W, C] position map."""

    def __init__(
        self, embedding_dim: int, image_embedding_size: tuple[int, int], device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.image_embedding_size = image_embedding_size
        super().__init__(fl.Parameter(*image_embedding_size, embedding_dim, **_placement(device, dtype)))


class RelativePositionAttention(fl.WeightedModule):
    """Multi-head attention over a fused [B, H, W, 3C] projection with decomposed relative-position bias:

        logits[(h, w), (kh, kw)] = q . k / sqrt(d)  +  q . R_v[h - kh + H - 1]  +  q . R_h[w - kw + W - 1]

    ``vertical_embedding`` holds R_v ([2H-1, d]), ``horizontal_embedding`` R_h ([2W-1, d]); q enters the bias terms
    unscaled; the vertical term is added first (the reference pins that order)."""

    def __init__(
        self, embedding_dim: int, num_heads: int, spatial_size: tuple[int, int], device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        super().__init__()
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.head_dim = embedding_dim // num_heads
        self.spatial_size = spatial_size

        def table(extent: int) -> nn.Parameter:
            return nn.Parameter(torch.zeros(2 * extent - 1, self.head_dim, **_placement(device, dtype)))

        self.horizontal_embedding = table(spatial_size[0])
        self.vertical_embedding = table(spatial_size[1])

    @property
    def device(self) -> Device:
        return self.horizontal_embedding.device

    @property
    def dtype(self) -> DType:
        return self.horizontal_embedding.dtype

    def compute_relative_coords(self, size: int) -> Tensor:
        """[size, size] table-row index of (query position, key position): q - k + size - 1."""
        positions = torch.arange(size)
        return (positions.unsqueeze(1) - positions.unsqueeze(0)) + (size - 1)

    def compute_relative_embedding(self, x: Tensor) -> tuple[Tensor, Tensor]:
        """Bias terms of queries ``x`` [B', H*W, d]: (horizontal [B', H, W, 1, Wk], vertical [B', H, W, Hk, 1])."""
        extent_h, extent_v = self.spatial_size
        queries = x.reshape(x.shape[0], extent_h, extent_v, -1)
        # q against EVERY table row, then pick, per (query position, key position), the row of their offset
        on_h_rows = queries @ self.horizontal_embedding.transpose(0, 1)          # [B', H, W, 2W-1]
        on_v_rows = queries @ self.vertical_embedding.transpose(0, 1)            # [B', H, W, 2H-1]
        pick_h = self.compute_relative_coords(extent_h).to(x.device)             # [W, Wk]: indexed by the query's w
        pick_v = self.compute_relative_coords(extent_v).to(x.device)             # [H, Hk]: indexed by the query's h
        batch = x.shape[0]
        horizontal = on_h_rows.gather(-1, pick_h.view(1, 1, extent_h, extent_h).expand(batch, extent_h, -1, -1))
        vertical = on_v_rows.gather(-1, pick_v.view(1, extent_v, 1, extent_v).expand(batch, -1, extent_v, -1))
        return horizontal.unsqueeze(-2), vertical.unsqueeze(-1)

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            return B.sam_attention(x, self.vertical_embedding, self.horizontal_embedding, self.num_heads)
        batch, height, width, _ = x.shape
        heads, tokens = self.num_heads, height * width
        # [B, HW, (3, heads, d)] -> three [B * heads, HW, d] operands
        query, key, value = (
            x.reshape(batch, tokens, 3, heads, self.head_dim).permute(2, 0, 3, 1, 4).reshape(3, batch * heads, tokens, self.head_dim)
        )
        bias_h, bias_v = self.compute_relative_embedding(query)
        logits = (query * self.head_dim**-0.5) @ key.transpose(-2, -1)
        logits = ((logits.view(-1, height, width, height, width) + bias_v) + bias_h).view(-1, tokens, tokens)
        mixed = logits.softmax(dim=-1) @ value
        return mixed.view(batch, heads, height, width, self.head_dim).permute(0, 2, 3, 1, 4).reshape(batch, height, width, -1)


class FusedSelfAttention(fl.Chain):
    """One Linear to q | k | v, the relative-position attention, an output projection."""

    def __init__(
        self, embedding_dim: int = 768, spatial_size: tuple[int, int] = (64, 64), num_heads: int = 1,
        use_bias: bool = True, is_causal: bool = False, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        assert embedding_dim % num_heads == 0, (
            f"Embedding dim (embedding_dim={embedding_dim}) must be divisible by num heads (num_heads={num_heads})"
        )
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.use_bias = use_bias
        self.is_causal = is_causal
        where = _placement(device, dtype)
        super().__init__(
            fl.Linear(embedding_dim, 3 * embedding_dim, bias=use_bias, **where),
            RelativePositionAttention(embedding_dim, num_heads, spatial_size, **where),
            fl.Linear(embedding_dim, embedding_dim, bias=True, **where),
        )


class FeedForward(fl.Chain):
    def __init__(
        self, embedding_dim: int, feedforward_dim: int, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.feedforward_dim = feedforward_dim
        where = _placement(device, dtype)
        super().__init__(
            fl.Linear(embedding_dim, feedforward_dim, bias=True, **where),
            fl.GeLU(),
            fl.Linear(feedforward_dim, embedding_dim, bias=True, **where),
        )


def _round_up(value: int, multiple: int) -> int:
    return -(-value // multiple) * multiple


class WindowPartition(fl.ContextModule):
    """[B, H, W, C] -> [B * nH * nW, ws, ws, C], zero padded up to whole windows; the geometry is left in the
    ``window_partition`` context for the matching ``WindowMerge``."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        batch, height, width, channels = x.shape
        geometry = self.use_context("window_partition")
        ws = geometry["window_size"]
        padded_h, padded_w = _round_up(height, ws), _round_up(width, ws)
        geometry.update(original_height=height, original_width=width, padded_height=padded_h, padded_width=padded_w)
        if x.is_cuda and channels % (16 // x.element_size()) == 0:
            return B.window_partition(x, ws)  # pad + partition in one launch
        if (padded_h, padded_w) != (height, width):
            x = pad(x, (0, 0, 0, padded_w - width, 0, padded_h - height))
        tiles = x.view(batch, padded_h // ws, ws, padded_w // ws, ws, channels).transpose(2, 3)
        return tiles.reshape(-1, ws, ws, channels)


class WindowMerge(fl.ContextModule):
    """Inverse of ``WindowPartition``; the padding is cropped away."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        geometry = self.use_context("window_partition")
        ws = geometry["window_size"]
        height, width = geometry["original_height"], geometry["original_width"]
        if x.is_cuda and x.shape[-1] % (16 // x.element_size()) == 0:
            return B.window_merge(x, ws, height, width)  # merge + crop in one launch
        rows, cols = geometry["padded_height"] // ws, geometry["padded_width"] // ws
        batch = x.shape[0] // (rows * cols)
        merged = x.view(batch, rows, cols, ws, ws, -1).transpose(2, 3).reshape(batch, rows * ws, cols * ws, -1)
        return merged[:, :height, :width, :].contiguous()


class TransformerLayer(fl.Chain):
    """Pre-norm block: x + attention(LN(x)) - windowed (``window_size``) or global - then x + MLP(LN(x))."""

    def __init__(
        self, embedding_dim: int, num_heads: int, feedforward_dim: int, image_embedding_size: tuple[int, int],
        window_size: int | None = None, layer_norm_eps: float = 1e-6, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.feedforward_dim = feedforward_dim
        self.window_size = window_size
        self.layer_norm_eps = layer_norm_eps
        self.image_embedding_size = image_embedding_size
        where = _placement(device, dtype)
        if window_size is None:  # global attention over the whole map; the trailing Reshape restores [H, W, C]
            enter, extent, leave = fl.Identity(), image_embedding_size, fl.Reshape(*image_embedding_size, embedding_dim)
        else:
            enter, extent, leave = WindowPartition(), (window_size, window_size), WindowMerge()
        super().__init__(
            fl.Residual(
                fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **where),
                enter,
                FusedSelfAttention(embedding_dim=embedding_dim, num_heads=num_heads, spatial_size=extent, **where),
                leave,
            ),
            fl.Residual(
                fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **where),
                FeedForward(embedding_dim=embedding_dim, feedforward_dim=feedforward_dim, **where),
            ),
        )

    def init_context(self) -> Contexts:
        return {"window_partition": {"window_size": self.window_size}}


class Neck(fl.Chain):
    """NHWC -> NCHW, 1x1 conv to 256 channels, LayerNorm2d, 3x3 conv, LayerNorm2d (no biases)."""

    def __init__(self, in_channels: int = 768, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.in_channels = in_channels
        where = _placement(device, dtype)
        width = 256
        super().__init__(
            fl.Permute(0, 3, 1, 2),
            fl.Conv2d(in_channels, width, kernel_size=1, use_bias=False, **where),
            fl.LayerNorm2d(channels=width, **where),
            fl.Conv2d(width, width, kernel_size=3, padding=1, use_bias=False, **where),
            fl.LayerNorm2d(channels=width, **where),
        )


class Transformer(fl.Chain):
    """The stack of ``TransformerLayer``s (a named Chain so that its key appears in the state dict)."""


class SAMViT(fl.Chain):
    """Patch embedding, learned positions, ``num_layers`` transformer layers (global attention at
    ``global_attention_indices``, 14 x 14 windows elsewhere), neck."""

    def __init__(
        self, embedding_dim: int, num_layers: int, num_heads: int,
        global_attention_indices: tuple[int, ...] | None = None, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.num_layers = num_layers
        self.num_heads = num_heads
        self.image_size = (1024, 1024)
        self.patch_size = 16
        self.window_size = 14
        self.image_embedding_size = tuple(side // self.patch_size for side in self.image_size)
        self.feed_forward_dim = 4 * embedding_dim
        self.global_attention_indices = global_attention_indices or tuple()
        where = _placement(device, dtype)
        grid = self.image_embedding_size

        def layer(index: int) -> TransformerLayer:
            return TransformerLayer(
                embedding_dim=embedding_dim, num_heads=num_heads, feedforward_dim=self.feed_forward_dim,
                window_size=None if index in self.global_attention_indices else self.window_size,
                image_embedding_size=grid, **where,
            )

        super().__init__(
            PatchEncoder(in_channels=3, out_channels=embedding_dim, patch_size=self.patch_size, **where),
            PositionalEncoder(embedding_dim=embedding_dim, image_embedding_size=grid, **where),
            Transformer(*map(layer, range(num_layers))),
            Neck(in_channels=embedding_dim, **where),
        )


class SAMViTH(SAMViT):
    """The published "huge" configuration: 1280 wide, 32 layers, 16 heads, global attention every 8th layer."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(
            embedding_dim=1280, num_layers=32, num_heads=16, global_attention_indices=(7, 15, 23, 31), device=device, dtype=dtype
        )
