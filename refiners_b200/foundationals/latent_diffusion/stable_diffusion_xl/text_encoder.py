"""SDXL's two text towers as one module: prompt -> ([B, 77, 768 + 1280] token embedding, [B, 1280] pooled embedding).

Contract (class names, trees, context name, call results) from
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/text_encoder.py: `TextEncoderWithPooling`
:13-62, `DoubleTextEncoder` :65-101.

Both towers are read at their PENULTIMATE layer (the chain minus its last transformer layer and final LayerNorm).  The
bigG tower additionally produces the pooled embedding SDXL feeds to its timestep embedding: the remaining layer and the
final norm are run on a side branch, projected (bias-free 1280 x 1280 Linear) and read at each prompt's end-of-text
position, which a probe behind the tokenizer records.  `TextEncoderWithPooling` is an ADAPTER around the bigG tower, so
that the tower keeps its plain tree (and checkpoint keys) when ejected.
"""

from __future__ import annotations

from typing import cast

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.context import Contexts
from refiners_b200.foundationals.clip.text_encoder import CLIPTextEncoderG, CLIPTextEncoderL
from refiners_b200.foundationals.clip.tokenizer import CLIPTokenizer

Device = torch.device
DType = torch.dtype

POOLING = "text_encoder_pooling"


class TextEncoderWithPooling(fl.Chain, Adapter[CLIPTextEncoderG]):
    def __init__(self, target: CLIPTextEncoderG, projection: fl.Linear | None = None) -> None:
        with self.setup_adapter(target=target):
            head = fl.Chain(
                target[-2:],  # last transformer layer + final LayerNorm
                projection or fl.Linear(in_features=1280, out_features=1280, bias=False, device=target.device, dtype=target.dtype),
                fl.Lambda(func=self.pool),
            )
            super().__init__(
                target.ensure_find(CLIPTokenizer),
                fl.SetContext(context=POOLING, key="end_of_text_index", callback=self.set_end_of_text_index),
                target[1:-2],  # ids -> penultimate hidden states
                fl.Parallel(fl.Identity(), head),
            )

    def init_context(self) -> Contexts:
        return {POOLING: {"end_of_text_index": []}}

    def __call__(self, text: str | list[str]) -> tuple[Tensor, Tensor]:
        return super().__call__(text)

    @property
    def tokenizer(self) -> CLIPTokenizer:
        return self.ensure_find(CLIPTokenizer)

    def set_end_of_text_index(self, end_of_text_index: list[int], tokens: Tensor) -> None:
        """Position of the first end-of-text token of every prompt (the bigG tokenizer pads with 0, so it is unique)."""
        marks = tokens == self.tokenizer.end_of_text_token_id
        for row in marks:
            (position,) = row.nonzero(as_tuple=True)
            end_of_text_index.append(cast(int, position.item()))

    def pool(self, x: Tensor) -> Tensor:
        positions = self.use_context(context_name=POOLING).get("end_of_text_index", [])
        assert len(positions) == x.shape[0], "End of text index not found."
        return torch.cat([x[i : i + 1, at, :] for i, at in enumerate(positions)], dim=0)


class DoubleTextEncoder(fl.Chain):
    def __init__(
        self,
        text_encoder_l: CLIPTextEncoderL | None = None,
        text_encoder_g: CLIPTextEncoderG | None = None,
        projection: fl.Linear | None = None,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        text_encoder_l = text_encoder_l or CLIPTextEncoderL(device=device, dtype=dtype)
        text_encoder_g = text_encoder_g or CLIPTextEncoderG(device=device, dtype=dtype)
        super().__init__(fl.Parallel(text_encoder_l[:-2], text_encoder_g), fl.Lambda(self.concatenate_embeddings))
        TextEncoderWithPooling(target=text_encoder_g, projection=projection).inject(self.layer("Parallel", fl.Parallel))

    def __call__(self, text: str | list[str]) -> tuple[Tensor, Tensor]:
        return super().__call__(text)

    def concatenate_embeddings(self, text_embedding_l: Tensor, text_embedding_with_pooling: tuple[Tensor, Tensor]) -> tuple[Tensor, Tensor]:
        text_embedding_g, pooled_text_embedding = text_embedding_with_pooling
        return torch.cat((text_embedding_l, text_embedding_g), dim=-1), pooled_text_embedding

    def structural_copy(self: "DoubleTextEncoder") -> "DoubleTextEncoder":
        """The pooling adapter refuses to be copied while injected: it is taken out, the plain tree copied, put back, and
        a new adapter (sharing the projection) is built around the copy's bigG tower."""
        pooling = self.ensure_find(TextEncoderWithPooling)
        pooling.eject()
        twin = super().structural_copy()
        pooling.inject()
        projection = pooling.layer(("Parallel", "Chain", "Linear"), fl.Linear)
        TextEncoderWithPooling(target=twin.ensure_find(CLIPTextEncoderG), projection=projection).inject(twin.layer("Parallel", fl.Parallel))
        return twin
