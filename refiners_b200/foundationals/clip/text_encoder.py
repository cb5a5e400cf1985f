"""CLIP text towers: the prompt side of the diffusion models (runs once per prompt).

Module tree / state-dict keys / constructor arguments are the contract of
/root/reference/src/refiners/foundationals/clip/text_encoder.py (`TokenEncoder` :8-29, `TransformerLayer` :32-91,
`CLIPTextEncoder` :94-188, the L / H / G presets :191-251):

    tokenizer -> ids moved to the tower's device -> token + position embeddings
              -> N x [ x + CausalSelfAttention(LN(x)),  x + FeedForward(LN(x)) ]  -> LayerNorm

On CUDA every LayerNorm / Linear / (quick-)GeLU is one of the library's kernels; the causal 77-token attention runs on the
CUDA-core flash kernel (the tensor-core kernels take no causal mask: at 77 keys a tower is launch-bound, not attention-bound).
"""

from __future__ import annotations

from typing import Any

import torch

import refiners_b200.fluxion.layers as fl
from refiners_b200.foundationals.clip.common import FeedForward, PositionalEncoder
from refiners_b200.foundationals.clip.tokenizer import CLIPTokenizer

Device = torch.device
DType = torch.dtype

# published towers: (width, layers, heads, feed-forward width, OpenAI's sigmoid GeLU?, padding token of the tokenizer)
_TOWERS: dict[str, tuple[int, int, int, int, bool, int | None]] = {
    "L": (768, 12, 12, 3072, True, None),      # CLIP ViT-L/14: SD 1.5, first SDXL encoder
    "H": (1024, 23, 16, 4096, False, None),    # OpenCLIP ViT-H/14: SD 2.x
    "G": (1280, 32, 20, 5120, False, 0),       # OpenCLIP ViT-bigG/14: second SDXL encoder; pads with token 0
}


class TokenEncoder(fl.Embedding):
    def __init__(self, vocabulary_size: int, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None) -> None:
        self.vocabulary_size = vocabulary_size
        self.embedding_dim = embedding_dim
        super().__init__(num_embeddings=vocabulary_size, embedding_dim=embedding_dim, device=device, dtype=dtype)


class TransformerLayer(fl.Chain):
    """Two pre-LayerNorm residual branches: causal self-attention, then the feed-forward."""

    def __init__(
        self, embedding_dim: int, feedforward_dim: int, num_attention_heads: int = 1, layer_norm_eps: float = 1e-5,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.num_attention_heads = num_attention_heads
        self.feedforward_dim = feedforward_dim
        self.layer_norm_eps = layer_norm_eps
        on: dict[str, Any] = {"device": device, "dtype": dtype}
        mixers = (
            fl.SelfAttention(embedding_dim=embedding_dim, num_heads=num_attention_heads, is_causal=True, **on),
            FeedForward(embedding_dim=embedding_dim, feedforward_dim=feedforward_dim, **on),
        )
        super().__init__(
            fl.Residual(fl.LayerNorm(normalized_shape=embedding_dim, eps=layer_norm_eps, **on), mixer) for mixer in mixers
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
        hyper = dict(
            embedding_dim=embedding_dim, max_sequence_length=max_sequence_length, vocabulary_size=vocabulary_size,
            num_layers=num_layers, num_attention_heads=num_attention_heads, feedforward_dim=feedforward_dim,
            layer_norm_eps=layer_norm_eps, use_quick_gelu=use_quick_gelu,
        )
        for name, value in hyper.items():  # echoed by repr(), read by adapters
            setattr(self, name, value)
        on: dict[str, Any] = {"device": device, "dtype": dtype}
        embed = fl.Sum(
            TokenEncoder(vocabulary_size=vocabulary_size, embedding_dim=embedding_dim, **on),
            PositionalEncoder(max_sequence_length=max_sequence_length, embedding_dim=embedding_dim, **on),
        )
        layers = [
            TransformerLayer(embedding_dim=embedding_dim, num_attention_heads=num_attention_heads, feedforward_dim=feedforward_dim,
                             layer_norm_eps=layer_norm_eps, **on)
            for _ in range(num_layers)
        ]
        super().__init__(
            tokenizer or CLIPTokenizer(sequence_length=max_sequence_length),
            fl.Converter(set_dtype=False),  # the ids go to the tower's device and stay integers
            embed,
            *layers,
            fl.LayerNorm(normalized_shape=embedding_dim, eps=layer_norm_eps, **on),
        )
        if use_quick_gelu:
            self._use_sigmoid_gelu()

    def _use_sigmoid_gelu(self) -> None:
        """OpenAI's original towers compute GeLU as x * sigmoid(1.702 x)."""
        for exact, parent in [*self.walk(fl.GeLU)]:
            parent.replace(old_module=exact, new_module=fl.GeLU(approximation=fl.GeLUApproximation.SIGMOID))


def _published(tag: str, device: Device | str | None, dtype: DType | None) -> dict[str, Any]:
    width, layers, heads, hidden, quick, pad = _TOWERS[tag]
    extra: dict[str, Any] = {} if pad is None else {"tokenizer": CLIPTokenizer(pad_token_id=pad)}
    return dict(embedding_dim=width, num_layers=layers, num_attention_heads=heads, feedforward_dim=hidden, use_quick_gelu=quick,
                device=device, dtype=dtype, **extra)


class CLIPTextEncoderL(CLIPTextEncoder):
    """768 wide, 12 layers, 12 heads, quick GeLU."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(**_published("L", device, dtype))


class CLIPTextEncoderH(CLIPTextEncoder):
    """1024 wide, 23 layers, 16 heads."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(**_published("H", device, dtype))


class CLIPTextEncoderG(CLIPTextEncoder):
    """1280 wide, 32 layers, 20 heads; its tokenizer pads with token 0."""

    def __init__(self, device: Device | str | None = None, dtype: DType | None = None) -> None:
        super().__init__(**_published("G", device, dtype))
