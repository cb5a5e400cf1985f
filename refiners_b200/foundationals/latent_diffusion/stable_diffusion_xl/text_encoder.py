"""SDXL's two text towers as one module: prompt -> ([B, 77, 768 + 1280] token embedding, [B, 1280] pooled embedding).

Class names, trees (= checkpoint keys), the context name and the call results are the contract of
/root/reference/src/refiners/foundationals/latent_diffusion/stable_diffusion_xl/text_encoder.py (`TextEncoderWithPooling`
:13-62, `DoubleTextEncoder` :65-101).

Both towers are read at their PENULTIMATE layer: the chain without its last transformer layer and final LayerNorm.  The bigG
tower also yields the pooled embedding SDXL feeds to its timestep embedding: a side branch runs the remaining layer and the
final norm, projects (bias-free 1280 x 1280 Linear) and reads each prompt at its end-of-text token, whose position a probe
behind the tokenizer has noted.  `TextEncoderWithPooling` is an ADAPTER around the bigG tower, so the tower keeps its plain
tree and keys when the adapter is ejected.
"""

from __future__ import annotations

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200.fluxion.adapters.adapter import Adapter
from refiners_b200.fluxion.context import Contexts
from refiners_b200.foundationals.clip.text_encoder import CLIPTextEncoderG, CLIPTextEncoderL
from refiners_b200.foundationals.clip.tokenizer import CLIPTokenizer

Device = torch.device
DType = torch.dtype

_POOLING = "text_encoder_pooling"     # context
_POSITIONS = "end_of_text_index"      # its key: one position per prompt of the batch, appended in order
_TAIL = 2                             # modules cut off a tower to read it at the penultimate layer: last layer + final norm


def first_end_of_text(tokens: Tensor, end_of_text_token_id: int) -> list[int]:
    """Column of the first end-of-text token in every row (argmax returns the first maximum)."""
    return (tokens == end_of_text_token_id).to(torch.uint8).argmax(dim=1).tolist()


class TextEncoderWithPooling(fl.Chain, Adapter[CLIPTextEncoderG]):
    def __init__(self, target: CLIPTextEncoderG, projection: fl.Linear | None = None) -> None:
        with self.setup_adapter(target=target):
            if projection is None:
                projection = fl.Linear(in_features=1280, out_features=1280, bias=False, device=target.device, dtype=target.dtype)
            body, top = target[1:-_TAIL], target[-_TAIL:]
            probe = fl.SetContext(context=_POOLING, key=_POSITIONS, callback=self.set_end_of_text_index)
            pooled = fl.Chain(top, projection, fl.Lambda(func=self.pool))
            super().__init__(target.ensure_find(CLIPTokenizer), probe, body, fl.Parallel(fl.Identity(), pooled))

    def init_context(self) -> Contexts:
        return {_POOLING: {_POSITIONS: []}}

    def __call__(self, text: str | list[str]) -> tuple[Tensor, Tensor]:
        return super().__call__(text)

    @property
    def tokenizer(self) -> CLIPTokenizer:
        return self.ensure_find(CLIPTokenizer)

    def set_end_of_text_index(self, end_of_text_index: list[int], tokens: Tensor) -> None:
        end_of_text_index.extend(first_end_of_text(tokens, self.tokenizer.end_of_text_token_id))

    def pool(self, x: Tensor) -> Tensor:
        positions = self.use_context(context_name=_POOLING).get(_POSITIONS, [])
        assert len(positions) == x.shape[0], "End of text index not found."
        rows = torch.arange(x.shape[0], device=x.device)
        return x[rows, torch.tensor(positions, device=x.device)]


class DoubleTextEncoder(fl.Chain):
    def __init__(
        self,
        text_encoder_l: CLIPTextEncoderL | None = None,
        text_encoder_g: CLIPTextEncoderG | None = None,
        projection: fl.Linear | None = None,
        device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        small = text_encoder_l or CLIPTextEncoderL(device=device, dtype=dtype)
        big = text_encoder_g or CLIPTextEncoderG(device=device, dtype=dtype)
        super().__init__(fl.Parallel(small[:-_TAIL], big), fl.Lambda(self.concatenate_embeddings))
        self._wrap_big_tower(self, big, projection)

    @staticmethod
    def _wrap_big_tower(owner: "DoubleTextEncoder", tower: CLIPTextEncoderG, projection: fl.Linear | None) -> None:
        TextEncoderWithPooling(target=tower, projection=projection).inject(owner.layer("Parallel", fl.Parallel))

    def __call__(self, text: str | list[str]) -> tuple[Tensor, Tensor]:
        return super().__call__(text)

    def concatenate_embeddings(self, text_embedding_l: Tensor, text_embedding_with_pooling: tuple[Tensor, Tensor]) -> tuple[Tensor, Tensor]:
        text_embedding_g, pooled_text_embedding = text_embedding_with_pooling
        return torch.cat((text_embedding_l, text_embedding_g), dim=-1), pooled_text_embedding

    def structural_copy(self: "DoubleTextEncoder") -> "DoubleTextEncoder":
        """An injected adapter cannot be copied: the pooling adapter is taken out, the plain tree copied, the adapter put
        back, and a new one - sharing the projection - built around the copy's bigG tower."""
        pooling = self.ensure_find(TextEncoderWithPooling)
        shared_projection = pooling.layer(("Parallel", "Chain", "Linear"), fl.Linear)
        pooling.eject()
        try:
            twin = super().structural_copy()
        finally:
            pooling.inject()
        self._wrap_big_tower(twin, twin.ensure_find(CLIPTextEncoderG), shared_projection)
        return twin
