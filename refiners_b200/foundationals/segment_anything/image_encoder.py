"""SAM ViT image encoder (ViT-H by default).

Tree shape follows /root/reference/src/refiners/foundationals/segment_anything/image_encoder.py:
`PatchEncoder` :9-34, `PositionalEncoder` :37-55, `RelativePositionAttention` :58-143,
`FusedSelfAttention` :146-190, `FeedForward`, `WindowPartition` / `WindowMerge` :206-236,
`TransformerLayer` :239-283, `Neck` :286-310, `SAMViT` / `SAMViTH` :316-368.

On CUDA `RelativePositionAttention` is one call into the C ABI (rb200_sam_attention): the
reference materialises the [B*heads, HW, HW] logits (512 MB per image in the global layers) and
rebuilds index tensors on the host every call; the kernel path never forms the logits in HBM.
"""

from __future__ import annotations

import torch
from torch import Tensor, nn

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.context import Contexts
from refiners_b200.fluxion.utils import pad

Device = torch.device
DType = torch.dtype


class PatchEncoder(fl.Chain):
    """Non-overlapping patch embedding: conv(k = stride = patch) then NCHW -> NHWC."""

    def __init__(
        self, in_channels: int, out_channels: int, patch_size: int = 16, use_bias: bool = True,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.in_channels, self.out_channels, self.patch_size, self.use_bias = in_channels, out_channels, patch_size, use_bias
        super().__init__(
            fl.Conv2d(
                in_channels,
                out_channels,
                kernel_size=(patch_size, patch_size),
                stride=(patch_size, patch_size),
                use_bias=use_bias,
                device=device,
                dtype=dtype,
            ),
            fl.Permute(0, 2, 3, 1),
        )


class PositionalEncoder(fl.Residual):
    def __init__(
        self, embedding_dim: int, image_embedding_size: tuple[int, int], device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.image_embedding_size = embedding_dim, image_embedding_size
        super().__init__(
            fl.Parameter(image_embedding_size[0], image_embedding_size[1], embedding_dim, device=device, dtype=dtype)
        )


class RelativePositionAttention(fl.WeightedModule):
    """Multi-head attention over an [B, H, W, 3C] fused projection with decomposed relative
    position terms: logits = q k^T d^-1/2 + q . R_v[h - kh] + q . R_h[w - kw]."""

    def __init__(
        self, embedding_dim: int, num_heads: int, spatial_size: tuple[int, int], device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        super().__init__()
        self.embedding_dim, self.num_heads = embedding_dim, num_heads
        self.head_dim = embedding_dim // num_heads
        self.spatial_size = spatial_size
        self.horizontal_embedding = nn.Parameter(torch.zeros(2 * spatial_size[0] - 1, self.head_dim, device=device, dtype=dtype))
        self.vertical_embedding = nn.Parameter(torch.zeros(2 * spatial_size[1] - 1, self.head_dim, device=device, dtype=dtype))

    @property
    def device(self) -> Device:
        return self.horizontal_embedding.device

    @property
    def dtype(self) -> DType:
        return self.horizontal_embedding.dtype

    def forward(self, x: Tensor) -> Tensor:
        if x.is_cuda:
            return B.sam_attention(x, self.vertical_embedding, self.horizontal_embedding, self.num_heads)
        batch, height, width, _ = x.shape
        heads, d = self.num_heads, self.head_dim
        qkv = x.reshape(batch, height * width, 3, heads, d).permute(2, 0, 3, 1, 4).reshape(3, batch * heads, height * width, d)
        q, k, v = qkv.unbind(0)
        rel_h, rel_v = self.compute_relative_embedding(q)
        logits = (q * d**-0.5) @ k.transpose(-2, -1)
        # vertical term first, then horizontal (the order is part of the reference's numerics)
        logits = ((logits.reshape(-1, height, width, height, width) + rel_v) + rel_h).reshape(logits.shape)
        out = logits.softmax(dim=-1) @ v
        return out.reshape(batch, heads, height, width, d).permute(0, 2, 3, 1, 4).reshape(batch, height, width, -1)

    def compute_relative_coords(self, size: int) -> Tensor:
        idx = torch.arange(size)
        return idx[:, None] - idx[None, :] + size - 1

    def compute_relative_embedding(self, x: Tensor) -> tuple[Tensor, Tensor]:
        width, height = self.spatial_size
        emb_h = self.horizontal_embedding[self.compute_relative_coords(width)]
        emb_v = self.vertical_embedding[self.compute_relative_coords(height)]
        x = x.reshape(x.shape[0], width, height, -1)
        rel_h = torch.einsum("bhwc,wkc->bhwk", x, emb_h).unsqueeze(-2)
        rel_v = torch.einsum("bhwc,hkc->bhwk", x, emb_v).unsqueeze(-1)
        return rel_h, rel_v


class FusedSelfAttention(fl.Chain):
    def __init__(
        self, embedding_dim: int = 768, spatial_size: tuple[int, int] = (64, 64), num_heads: int = 1,
        use_bias: bool = True, is_causal: bool = False, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        assert embedding_dim % num_heads == 0, (
            f"Embedding dim (embedding_dim={embedding_dim}) must be divisible by num heads (num_heads={num_heads})"
        )
        self.embedding_dim, self.num_heads, self.use_bias, self.is_causal = embedding_dim, num_heads, use_bias, is_causal
        super().__init__(
            fl.Linear(embedding_dim, 3 * embedding_dim, bias=use_bias, device=device, dtype=dtype),
            RelativePositionAttention(embedding_dim, num_heads, spatial_size, device=device, dtype=dtype),
            fl.Linear(embedding_dim, embedding_dim, bias=True, device=device, dtype=dtype),
        )


class FeedForward(fl.Chain):
    def __init__(
        self, embedding_dim: int, feedforward_dim: int, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.feedforward_dim = embedding_dim, feedforward_dim
        super().__init__(
            fl.Linear(embedding_dim, feedforward_dim, bias=True, device=device, dtype=dtype),
            fl.GeLU(),
            fl.Linear(feedforward_dim, embedding_dim, bias=True, device=device, dtype=dtype),
        )


class WindowPartition(fl.ContextModule):
    """[B, H, W, C] -> [B * nH * nW, ws, ws, C] (zero padding up to multiples of the window)."""

    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        batch, height, width, channels = x.shape
        ctx = self.use_context("window_partition")
        ws = ctx["window_size"]
        ph, pw = (ws - height % ws) % ws, (ws - width % ws) % ws
        hp, wp = height + ph, width + pw
        ctx.update({"original_height": height, "original_width": width, "padded_height": hp, "padded_width": wp})
        if x.is_cuda and channels % (16 // x.element_size()) == 0:
            return B.window_partition(x, ws)  # pad + partition in one launch
        if ph or pw:
            x = pad(x, (0, 0, 0, pw, 0, ph))
        x = x.view(batch, hp // ws, ws, wp // ws, ws, channels)
        return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, channels)


class WindowMerge(fl.ContextModule):
    def __init__(self) -> None:
        super().__init__()

    def forward(self, x: Tensor) -> Tensor:
        ctx = self.use_context("window_partition")
        ws = ctx["window_size"]
        hp, wp = ctx["padded_height"], ctx["padded_width"]
        height, width = ctx["original_height"], ctx["original_width"]
        if x.is_cuda and x.shape[-1] % (16 // x.element_size()) == 0:
            return B.window_merge(x, ws, height, width)  # merge + crop in one launch
        batch = x.shape[0] // (hp * wp // ws // ws)
        x = x.view(batch, hp // ws, wp // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).contiguous().view(batch, hp, wp, -1)
        if hp > height or wp > width:
            x = x[:, :height, :width, :].contiguous()
        return x


class TransformerLayer(fl.Chain):
    def __init__(
        self, embedding_dim: int, num_heads: int, feedforward_dim: int, image_embedding_size: tuple[int, int],
        window_size: int | None = None, layer_norm_eps: float = 1e-6, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.num_heads, self.feedforward_dim, self.window_size = embedding_dim, num_heads, feedforward_dim, window_size
        self.layer_norm_eps, self.image_embedding_size = layer_norm_eps, image_embedding_size
        windowed = window_size is not None
        spatial = (window_size, window_size) if windowed else image_embedding_size
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            fl.Residual(
                fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **kw),
                WindowPartition() if windowed else fl.Identity(),
                FusedSelfAttention(embedding_dim=embedding_dim, num_heads=num_heads, spatial_size=spatial, **kw),
                WindowMerge() if windowed else fl.Reshape(image_embedding_size[0], image_embedding_size[1], embedding_dim),
            ),
            fl.Residual(
                fl.LayerNorm(embedding_dim, eps=layer_norm_eps, **kw),
                FeedForward(embedding_dim=embedding_dim, feedforward_dim=feedforward_dim, **kw),
            ),
        )

    def init_context(self) -> Contexts:
        return {"window_partition": {"window_size": self.window_size}}


class Neck(fl.Chain):
    def __init__(self, in_channels: int = 768, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.in_channels = in_channels
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            fl.Permute(0, 3, 1, 2),
            fl.Conv2d(in_channels, 256, kernel_size=1, use_bias=False, **kw),
            fl.LayerNorm2d(channels=256, **kw),
            fl.Conv2d(256, 256, kernel_size=3, padding=1, use_bias=False, **kw),
            fl.LayerNorm2d(channels=256, **kw),
        )


class Transformer(fl.Chain):
    pass


class SAMViT(fl.Chain):
    """[B, 3, 1024, 1024] image -> [B, 256, 64, 64] embedding."""

    def __init__(
        self, embedding_dim: int, num_layers: int, num_heads: int,
        global_attention_indices: tuple[int, ...] | None = None, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.num_layers, self.num_heads = embedding_dim, num_layers, num_heads
        self.image_size = (1024, 1024)
        self.patch_size, self.window_size = 16, 14
        self.image_embedding_size = (self.image_size[0] // self.patch_size, self.image_size[1] // self.patch_size)
        self.feed_forward_dim = 4 * self.embedding_dim
        self.global_attention_indices = global_attention_indices or tuple()
        kw = dict(device=device, dtype=dtype)
        super().__init__(
            PatchEncoder(in_channels=3, out_channels=embedding_dim, patch_size=self.patch_size, **kw),
            PositionalEncoder(embedding_dim=embedding_dim, image_embedding_size=self.image_embedding_size, **kw),
            Transformer(
                TransformerLayer(
                    embedding_dim=embedding_dim,
                    num_heads=num_heads,
                    feedforward_dim=self.feed_forward_dim,
                    window_size=None if i in self.global_attention_indices else self.window_size,
                    image_embedding_size=self.image_embedding_size,
                    **kw,
                )
                for i in range(num_layers)
            ),
            Neck(in_channels=embedding_dim, **kw),
        )


class SAMViTH(SAMViT):
    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(
            embedding_dim=1280,
            num_layers=32,
            num_heads=16,
            global_attention_indices=(7, 15, 23, 31),
            device=device,
            dtype=dtype,
        )
