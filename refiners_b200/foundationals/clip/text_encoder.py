"""CLIP text towers: the prompt side of the diffusion models (runs once per prompt).

Module tree / state-dict keys / constructor arguments follow /root/reference/src/refiners/foundationals/clip/text_encoder.py
(`TokenEncoder` :8-29, `TransformerLayer` :32-91, `CLIPTextEncoder` :94-188 and the L / H / G presets :191-251):
tokenizer -> ids on the model's device -> token + position embeddings -> N pre-LN transformer layers with CAUSAL
self-attention -> final LayerNorm.  On CUDA every LayerNorm / Linear / (quick-)GeLU is one of the library's kernels; the
causal 77-token attention runs on the CUDA-core flash kernel (the tensor-core kernels do not take a causal mask: at 77 keys
the whole tower is launch-bound, not attention-bound).
"""

from __future__ import annotations

import torch

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.clip.common import FeedForward, PositionalEncoder
from refiners_b200.foundationals.clip.tokenizer import CLIPTokenizer

Device = torch.device
DType = torch.dtype


class TokenEncoder(fl.Embedding):
    def __init__(self, vocabulary_size: int, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.vocabulary_size = vocabulary_size
        self.embedding_dim = embedding_dim
        super().__init__(num_embeddings=vocabulary_size, embedding_dim=embedding_dim, device=device, dtype=dtype)


class TransformerLayer(fl.Chain):
    def __init__(
        self, embedding_dim: int, feedforward_dim: int, num_attention_heads: int = 1, layer_norm_eps: float = 1e-5,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.num_attention_heads = num_attention_heads
        self.feedforward_dim = feedforward_dim
        self.layer_norm_eps = layer_norm_eps
        where = {"device": device, "dtype": dtype}
        norm = lambda: fl.LayerNorm(normalized_shape=embedding_dim, eps=layer_norm_eps, **where)  # noqa: E731
        super().__init__(
            fl.Residual(norm(), fl.SelfAttention(embedding_dim=embedding_dim, num_heads=num_attention_heads, is_causal=True, **where)),
            fl.Residual(norm(), FeedForward(embedding_dim=embedding_dim, feedforward_dim=feedforward_dim, **where)),
        )


class CLIPTextEncoder(fl.Chain):
    def __init__(
        self,
        embedding_dim: int = 768,
        max_sequence_length: int = 77,
        vocabulary_size: int = 49408,
        num_layers: int = 12,
        num_attention_heads: int = 12,
        feedforward_dim: int = 3072,
        layer_norm_eps: float = 1e-5,
        use_quick_gelu: bool = False,
        tokenizer: CLIPTokenizer | None = None,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.max_sequence_length = max_sequence_length
        self.vocabulary_size = vocabulary_size
        self.num_layers = num_layers
        self.num_attention_heads = num_attention_heads
        self.feedforward_dim = feedforward_dim
        self.layer_norm_eps = layer_norm_eps
        self.use_quick_gelu = use_quick_gelu
        where = {"device": device, "dtype": dtype}
        super().__init__(
            tokenizer or CLIPTokenizer(sequence_length=max_sequence_length),
            fl.Converter(set_dtype=False),  # ids move to the tower's device, and stay integers
            fl.Sum(
                TokenEncoder(vocabulary_size=vocabulary_size, embedding_dim=embedding_dim, **where),
                PositionalEncoder(max_sequence_length=max_sequence_length, embedding_dim=embedding_dim, **where),
            ),
            *(
                TransformerLayer(
                    embedding_dim=embedding_dim, num_attention_heads=num_attention_heads, feedforward_dim=feedforward_dim,
                    layer_norm_eps=layer_norm_eps, **where,
                )
                for _ in range(num_layers)
            ),
            fl.LayerNorm(normalized_shape=embedding_dim, eps=layer_norm_eps, **where),
        )
        if use_quick_gelu:  # OpenAI's original towers: x * sigmoid(1.702 x)
            for gelu, parent in [*self.walk(fl.GeLU)]:
                parent.replace(old_module=gelu, new_module=fl.GeLU(approximation=fl.GeLUApproximation.SIGMOID))


class CLIPTextEncoderL(CLIPTextEncoder):
    """CLIP ViT-L/14's text tower (SD 1.5, first SDXL encoder): 768 wide, 12 layers, 12 heads, quick GeLU."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(embedding_dim=768, num_layers=12, num_attention_heads=12, feedforward_dim=3072, use_quick_gelu=True,
                         device=device, dtype=dtype)


class CLIPTextEncoderH(CLIPTextEncoder):
    """OpenCLIP ViT-H/14's text tower (SD 2.x): 1024 wide, 23 layers, 16 heads."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(embedding_dim=1024, num_layers=23, num_attention_heads=16, feedforward_dim=4096, device=device, dtype=dtype)


class CLIPTextEncoderG(CLIPTextEncoder):
    """OpenCLIP ViT-bigG/14's text tower (second SDXL encoder): 1280 wide, 32 layers, 20 heads; padded with token 0."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        tokenizer = CLIPTokenizer(pad_token_id=0)
        super().__init__(embedding_dim=1280, num_layers=32, num_attention_heads=20, feedforward_dim=5120, tokenizer=tokenizer,
                         device=device, dtype=dtype)
