"""Blocks shared by the CLIP text and vision towers (contract: /root/reference/src/refiners/foundationals/clip/common.py:7-49)."""

from __future__ import annotations

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl

Device = torch.device
DType = torch.dtype


class PositionalEncoder(fl.Chain):
    """Learned absolute positions, looked up for the first ``x.shape[1]`` indices."""

    def __init__(
        self, max_sequence_length: int, embedding_dim: int, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.max_sequence_length, self.embedding_dim = max_sequence_length, embedding_dim
        super().__init__(
            fl.Lambda(func=self.get_position_ids),
            fl.Embedding(num_embeddings=max_sequence_length, embedding_dim=embedding_dim, device=device, dtype=dtype),
        )

    @property
    def position_ids(self) -> Tensor:
        return torch.arange(end=self.max_sequence_length, device=self.device).reshape(1, -1)

    def get_position_ids(self, x: Tensor) -> Tensor:
        return self.position_ids[:, : x.shape[1]]


class FeedForward(fl.Chain):
    def __init__(
        self, embedding_dim: int, feedforward_dim: int, device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.embedding_dim, self.feedforward_dim = embedding_dim, feedforward_dim
        kw = dict(device=device, dtype=dtype)
        super().__init__(fl.Linear(embedding_dim, feedforward_dim, **kw), fl.GeLU(), fl.Linear(feedforward_dim, embedding_dim, **kw))
