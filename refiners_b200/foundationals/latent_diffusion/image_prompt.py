"""IP-Adapter: image-prompt conditioning injected into every text cross-attention.

Per-step behaviour follows /root/reference/src/refiners/foundationals/latent_diffusion/image_prompt.py:
`ImageProjection` :24-45, `ImageCrossAttention` :237-278, `CrossAttentionAdapter` :281-347,
`IPAdapter` :350-455, and stable_diffusion_xl/image_prompt.py:9-65 (`SDXLIPAdapter`).

Scope: the CLIP image encoder runs once per prompt, outside the denoising loop - it
is an optional constructor argument here and the benchmarks feed a synthetic ``clip_image_embedding`` through
``set_clip_image_embedding``.  The fine-grained image projection (`PerceiverResampler`, perceiver.py) is built.

B200 addition: after injection each cross-attention contains ``Sum(SDPA, ImageCrossAttention)``;
on CUDA that Sum runs as ONE flash-attention launch with two key/value sets and two independent
softmaxes (``o = A(q,k_t,v_t) + s * A(q,k_i,v_i)``) - registered below as a Sum fuser, the tree is
untouched.
"""

from __future__ import annotations

from typing import Any, Generic, TypeVar

import torch
from torch import Tensor, nn

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.engine import fusion
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.layers.leaves import ScaledDotProductAttention
from refiners_b200.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock2d
from refiners_b200.foundationals.latent_diffusion.perceiver import PerceiverResampler

T = TypeVar("T", bound=fl.Chain)
TIPAdapter = TypeVar("TIPAdapter", bound="IPAdapter[Any]")
Device = torch.device
DType = torch.dtype


class ImageProjection(fl.Chain):
    """CLIP image embedding [B, E] -> ``num_tokens`` text-space tokens [B, T, C]."""

    def __init__(
        self,
        clip_image_embedding_dim: int = 1024,
        clip_text_embedding_dim: int = 768,
        num_tokens: int = 4,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.clip_image_embedding_dim = clip_image_embedding_dim
        self.clip_text_embedding_dim = clip_text_embedding_dim
        self.num_tokens = num_tokens
        super().__init__(
            fl.Linear(clip_image_embedding_dim, clip_text_embedding_dim * num_tokens, device=device, dtype=dtype),
            fl.Reshape(num_tokens, clip_text_embedding_dim),
            fl.LayerNorm(normalized_shape=clip_text_embedding_dim, device=device, dtype=dtype),
        )


class ImageCrossAttention(fl.Chain):
    """(q, k_text, v_text) -> scale * SDPA(q, Wk' e, Wv' e) with e = ``ip_adapter.clip_image_embedding``."""

    def __init__(self, text_cross_attention: fl.Attention, scale: float = 1.0) -> None:
        self._multiply = [fl.Multiply(scale)]
        tca = text_cross_attention
        kw = dict(bias=tca.use_bias, device=tca.device, dtype=tca.dtype)
        super().__init__(
            fl.Distribute(
                fl.Identity(),
                fl.Chain(
                    fl.UseContext(context="ip_adapter", key="clip_image_embedding"),
                    fl.Linear(tca.key_embedding_dim, tca.inner_dim, **kw),
                ),
                fl.Chain(
                    fl.UseContext(context="ip_adapter", key="clip_image_embedding"),
                    fl.Linear(tca.value_embedding_dim, tca.inner_dim, **kw),
                ),
            ),
            ScaledDotProductAttention(num_heads=tca.num_heads, is_causal=tca.is_causal),
            self.multiply,
        )

    @property
    def multiply(self) -> fl.Multiply:
        return self._multiply[0]

    @property
    def scale(self) -> float:
        return self.multiply.scale

    @scale.setter
    def scale(self, value: float) -> None:
        self.multiply.scale = value


class CrossAttentionAdapter(fl.Chain, Adapter[fl.Attention]):
    def __init__(self, target: fl.Attention, scale: float = 1.0) -> None:
        with self.setup_adapter(target):
            super().__init__(target)
        self._image_cross_attention = [ImageCrossAttention(text_cross_attention=target, scale=scale)]

    def inject(self, parent: fl.Chain | None = None) -> "CrossAttentionAdapter":
        sdpa = self.target.ensure_find(ScaledDotProductAttention)
        # the text SDPA becomes Sum(text SDPA, image cross-attention)
        self.target.replace(old_module=sdpa, new_module=fl.Sum(sdpa, self.image_cross_attention))
        return super().inject(parent)

    def eject(self) -> None:
        holder = self.target.ensure_find_parent(self.image_cross_attention)
        holder.remove(self.image_cross_attention)
        sdpa = holder.layer("ScaledDotProductAttention", ScaledDotProductAttention)
        self.target.replace(old_module=holder, new_module=sdpa)
        super().eject()

    @property
    def image_cross_attention(self) -> ImageCrossAttention:
        return self._image_cross_attention[0]

    @property
    def image_key_projection(self) -> fl.Linear:
        return self.image_cross_attention.layer(("Distribute", 1, "Linear"), fl.Linear)

    @property
    def image_value_projection(self) -> fl.Linear:
        return self.image_cross_attention.layer(("Distribute", 2, "Linear"), fl.Linear)

    @property
    def scale(self) -> float:
        return self.image_cross_attention.scale

    @scale.setter
    def scale(self, value: float) -> None:
        self.image_cross_attention.scale = value

    def load_weights(self, key_tensor: Tensor, value_tensor: Tensor) -> None:
        self.image_key_projection.weight = nn.Parameter(key_tensor)
        self.image_value_projection.weight = nn.Parameter(value_tensor)
        self.image_cross_attention.to(self.device, self.dtype)


class IPAdapter(Generic[T], fl.Chain, Adapter[T]):
    """Image-prompt adapter for a latent-diffusion UNet: one `CrossAttentionAdapter` per text
    cross-attention (70 in SDXL)."""

    def __init__(
        self,
        target: T,
        clip_image_encoder: fl.Chain | None,
        image_proj: fl.Module | None,
        scale: float = 1.0,
        fine_grained: bool = False,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        with self.setup_adapter(target):
            super().__init__(target)
        self.fine_grained = fine_grained
        # list-wrapped: these are not part of the UNet's state dict
        self._clip_image_encoder = [clip_image_encoder]
        if fine_grained and clip_image_encoder is not None:
            self._grid_image_encoder = [self.convert_to_grid_features(clip_image_encoder)]
        self._image_proj = [image_proj]
        self.sub_adapters = [
            CrossAttentionAdapter(target=attn, scale=scale)
            for attn in target.layers(fl.Attention)
            if type(attn) is not fl.SelfAttention
        ]
        if weights is not None:
            if image_proj is not None:
                image_proj.load_state_dict(
                    {k.removeprefix("image_proj."): v for k, v in weights.items() if k.startswith("image_proj.")}
                )
            for i, sub in enumerate(self.sub_adapters):
                pair = [v for k, v in weights.items() if k.startswith(f"ip_adapter.{i:03d}.")]
                assert len(pair) == 2
                sub.load_weights(*pair)

    @property
    def clip_image_encoder(self) -> fl.Chain:
        enc = self._clip_image_encoder[0]
        assert enc is not None, "no CLIP image encoder attached (it runs outside the denoising loop)"
        return enc

    @property
    def image_proj(self) -> fl.Module:
        proj = self._image_proj[0]
        assert proj is not None, "no image projection attached"
        return proj

    def inject(self: TIPAdapter, parent: fl.Chain | None = None) -> TIPAdapter:
        for sub in self.sub_adapters:
            sub.inject()
        return super().inject(parent)

    def eject(self) -> None:
        for sub in self.sub_adapters:
            sub.eject()
        super().eject()

    @property
    def scale(self) -> float:
        return self.sub_adapters[0].scale

    @scale.setter
    def scale(self, value: float) -> None:
        for sub in self.sub_adapters:
            sub.scale = value

    def set_clip_image_embedding(self, image_embedding: Tensor) -> None:
        """[B, T, C] image tokens read by every `ImageCrossAttention` (T = 4, or 16 fine-grained)."""
        self.set_context("ip_adapter", {"clip_image_embedding": image_embedding})

    # ---- once per image prompt (outside the denoising loop): image(s) -> the context tensor above
    @property
    def grid_image_encoder(self) -> fl.Chain:
        assert hasattr(self, "_grid_image_encoder"), "fine-grained prompts need a CLIP image encoder at construction"
        return self._grid_image_encoder[0]

    @staticmethod
    def convert_to_grid_features(clip_image_encoder: fl.Chain) -> fl.Chain:
        """The encoder truncated to patch features for the PerceiverResampler: a structural copy (shared weights)
        without [CLS] pooling, final LayerNorm and projection, and without the last transformer layer - the
        penultimate hidden states of ViT-H (image_prompt.py:553-565 in the reference)."""
        grid = clip_image_encoder.structural_copy()
        assert isinstance(grid[-1], fl.Linear) and isinstance(grid[-2], fl.LayerNorm) and isinstance(grid[-3], fl.Lambda)
        for _ in range(3):
            grid.pop()
        layers = grid[-1]
        assert isinstance(layers, fl.Chain) and len(layers) == 32
        layers.pop()
        return grid

    def preprocess_image(
        self, image: Any, size: tuple[int, int] = (224, 224), mean: list[float] | None = None, std: list[float] | None = None,
    ) -> Tensor:
        """PIL image -> ``[1, 3, H, W]`` resized and normalised with OpenAI CLIP's statistics, on the UNet's device/dtype."""
        from refiners_b200.fluxion.utils import image_to_tensor

        pixels = image_to_tensor(image.resize(size), device=self.target.device, dtype=self.target.dtype)
        stats = lambda values: torch.tensor(values, device=pixels.device, dtype=pixels.dtype).reshape(1, -1, 1, 1)  # noqa: E731
        mean = [0.48145466, 0.4578275, 0.40821073] if mean is None else mean
        std = [0.26862954, 0.26130258, 0.27577711] if std is None else std
        return (pixels - stats(mean)) / stats(std)

    def _compute_clip_image_embedding(self, image_prompt: Tensor) -> tuple[Tensor, Tensor]:
        encoder = self.grid_image_encoder if self.fine_grained else self.clip_image_encoder
        features = encoder(image_prompt)
        conditional = self.image_proj(features)
        # the unconditional prompt: a zero embedding (plain) or the features of a black image (fine-grained)
        blank = encoder(torch.zeros_like(image_prompt)) if self.fine_grained else torch.zeros_like(features)
        return self.image_proj(blank), conditional

    def compute_clip_image_embedding(
        self, image_prompt: Any, weights: list[float] | None = None, concat_batches: bool = True,
    ) -> Tensor:
        """Image prompt(s) (PIL image, list of PIL images, or a preprocessed tensor) -> the ``[2, T', C]`` (or
        ``[2B, T, C]``) tensor for `set_clip_image_embedding`, unconditional half first.  ``weights`` scale each
        image's conditional tokens; ``concat_batches`` joins several images into one longer token sequence
        (image_prompt.py:457-511 in the reference)."""
        if isinstance(image_prompt, list):
            image_prompt = torch.cat([self.preprocess_image(image) for image in image_prompt])
        elif not isinstance(image_prompt, Tensor):
            image_prompt = self.preprocess_image(image_prompt)
        negative, conditional = self._compute_clip_image_embedding(image_prompt)
        count = image_prompt.shape[0]
        if weights is not None:
            assert len(weights) == count, f"Got {len(weights)} weights for {count} images"
            if any(w != 1.0 for w in weights):
                conditional *= torch.tensor(weights, device=conditional.device, dtype=conditional.dtype).unsqueeze(-1).unsqueeze(-1)
        if count > 1 and concat_batches:
            negative = torch.cat(negative.chunk(count), dim=1)
            conditional = torch.cat(conditional.chunk(count), dim=1)
        return torch.cat((negative, conditional))

    def project_image_embedding(self, clip_embedding: Tensor, negative: Tensor | None = None) -> Tensor:
        """image_proj(clip embedding), with the unconditional half first when given
        (reference image_prompt.py:457-511, minus the encoder call)."""
        cond = self.image_proj(clip_embedding)
        if negative is None:
            return cond
        return torch.cat((self.image_proj(negative), cond))


class SDXLIPAdapter(IPAdapter[fl.Chain]):
    def __init__(
        self,
        target: fl.Chain,
        clip_image_encoder: fl.Chain | None = None,
        image_proj: fl.Module | None = None,
        scale: float = 1.0,
        fine_grained: bool = False,
        weights: dict[str, Tensor] | None = None,
        clip_image_embedding_dim: int = 1024,
    ) -> None:
        if image_proj is None:
            xattn = target.ensure_find(CrossAttentionBlock2d)
            if fine_grained:
                # IP-Adapter "plus": 16 tokens resampled from the CLIP-H patch features (1280 wide before the
                # final projection); the latent width stays 1280 whatever the text width (sdxl/image_prompt.py:43-53)
                image_proj = PerceiverResampler(
                    latents_dim=1280, num_attention_layers=4, num_attention_heads=20, head_dim=64, num_tokens=16,
                    input_dim=1280, output_dim=xattn.context_embedding_dim, device=target.device, dtype=target.dtype,
                )
            else:
                image_proj = ImageProjection(
                    clip_image_embedding_dim=clip_image_embedding_dim,
                    clip_text_embedding_dim=xattn.context_embedding_dim,
                    device=target.device,
                    dtype=target.dtype,
                )
        elif fine_grained:
            assert isinstance(image_proj, PerceiverResampler)
        super().__init__(
            target=target,
            clip_image_encoder=clip_image_encoder,
            image_proj=image_proj,
            scale=scale,
            fine_grained=fine_grained,
            weights=weights,
        )


class SD1IPAdapter(IPAdapter[fl.Chain]):
    """IP-Adapter on an SD 1.5 UNet (contract: stable_diffusion_1/image_prompt.py:9-56 of the reference): four image tokens
    from the pooled CLIP-H embedding, or ("plus", ``fine_grained``) 16 tokens resampled from its patch features with a
    768-wide, 12-head PerceiverResampler."""

    def __init__(
        self,
        target: fl.Chain,
        clip_image_encoder: fl.Chain | None = None,
        image_proj: fl.Module | None = None,
        scale: float = 1.0,
        fine_grained: bool = False,
        weights: dict[str, Tensor] | None = None,
    ) -> None:
        if clip_image_encoder is None:
            from refiners_b200.foundationals.clip.image_encoder import CLIPImageEncoderH

            clip_image_encoder = CLIPImageEncoderH(device=target.device, dtype=target.dtype)
        if image_proj is not None:
            assert not fine_grained or isinstance(image_proj, PerceiverResampler)
        else:
            text_width = target.ensure_find(CrossAttentionBlock2d).context_embedding_dim
            on = dict(device=target.device, dtype=target.dtype)
            if fine_grained:
                image_proj = PerceiverResampler(
                    latents_dim=text_width, num_attention_layers=4, num_attention_heads=12, head_dim=64, num_tokens=16,
                    input_dim=clip_image_encoder.embedding_dim, output_dim=text_width, **on,
                )
            else:
                image_proj = ImageProjection(
                    clip_image_embedding_dim=clip_image_encoder.output_dim, clip_text_embedding_dim=text_width, **on,
                )
        super().__init__(
            target=target, clip_image_encoder=clip_image_encoder, image_proj=image_proj, scale=scale, fine_grained=fine_grained,
            weights=weights,
        )


# --------------------------------------------------------------------------- fused execution
def _fuse_text_image_attention(chain: fl.Sum, inputs: tuple[Any, ...]) -> Any:
    """Sum(SDPA, ImageCrossAttention)(q, k, v) as one dual-KV attention launch."""
    if len(chain) != 2 or len(inputs) != 3:
        return NotImplemented
    sdpa, ica = chain[0], chain[1]
    if type(sdpa) is not ScaledDotProductAttention or type(ica) is not ImageCrossAttention:
        return NotImplemented
    q, k, v = inputs
    if not (isinstance(q, Tensor) and q.is_cuda) or sdpa.is_causal or sdpa.slice_size:
        return NotImplemented
    kids = list(ica)
    if len(kids) != 3 or type(kids[0]) is not fl.Distribute or type(kids[1]) is not ScaledDotProductAttention:
        return NotImplemented
    if type(kids[2]) is not fl.Multiply or kids[2].bias != 0.0 or kids[1].num_heads != sdpa.num_heads or kids[1].is_causal:
        return NotImplemented
    dist = kids[0]
    if len(dist) != 3 or type(dist[0]) is not fl.Identity:
        return NotImplemented
    if fusion._hooked(sdpa, ica, kids[1], kids[2]):
        return NotImplemented
    k2 = dist[1](k)   # Chain(UseContext, Linear): ignores its input, projects the image tokens
    v2 = dist[2](v)
    return B.sdpa(q, k, v, sdpa.num_heads, False, k2=k2, v2=v2, scale2=float(kids[2].scale))


fusion.register_sum_fuser(_fuse_text_image_attention)
