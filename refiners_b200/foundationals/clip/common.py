"""Blocks shared by the CLIP text and vision towers.

Names, constructor arguments and the resulting trees (= state-dict keys and ``repr``) are the contract of
/root/reference/src/refiners/foundationals/clip/common.py:7-49; there is nothing else in them."""

from __future__ import annotations

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl

Device = torch.device
DType = torch.dtype


class PositionalEncoder(fl.Chain):
    """Learned absolute positions: row ``i`` of the table for token ``i`` of whatever sequence comes in."""

    def __init__(
        self, max_sequence_length: int, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.max_sequence_length = max_sequence_length
        self.embedding_dim = embedding_dim
        table = fl.Embedding(num_embeddings=max_sequence_length, embedding_dim=embedding_dim, device=device, dtype=dtype)
        super().__init__(fl.Lambda(func=self.get_position_ids), table)

    def get_position_ids(self, x: Tensor) -> Tensor:
        length = x.shape[1]
        assert length <= self.max_sequence_length, f"{length} tokens, but only {self.max_sequence_length} positions were learned"
        return torch.arange(length, device=self.device)[None]

    @property
    def position_ids(self) -> Tensor:
        """All positions, ``[1, max_sequence_length]``."""
        return torch.arange(self.max_sequence_length, device=self.device)[None]


class FeedForward(fl.Chain):
    """widen -> GeLU -> narrow (the activation is swapped for the sigmoid approximation in OpenAI's original towers)."""

    def __init__(
        self, embedding_dim: int, feedforward_dim: int, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim = embedding_dim
        self.feedforward_dim = feedforward_dim
        widths = (embedding_dim, feedforward_dim, embedding_dim)
        widen, narrow = (fl.Linear(a, b, device=device, dtype=dtype) for a, b in zip(widths, widths[1:]))
        super().__init__(widen, fl.GeLU(), narrow)
