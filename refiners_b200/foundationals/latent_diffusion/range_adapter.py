"""Timestep conditioning: sinusoidal encoder and the per-ResidualBlock channel bias.

Follows /root/reference/src/refiners/foundationals/latent_diffusion/range_adapter.py
(`compute_sinusoidal_embedding` :11-22, `RangeEncoder` :25-44, `RangeAdapter2d` :47-86).

B200 addition: on CUDA, ``RangeAdapter2d`` = conv + Linear(SiLU(temb))[:, :, None, None] runs as
one conv launch whose epilogue adds the per-sample channel bias (the tiny [B, C] GEMM is a
second launch); decided per call, the tree is untouched.
"""

from __future__ import annotations

import math
from typing import Any

import torch
from torch import Tensor

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.adapters.adapter import Adapter

Device = torch.device
DType = torch.dtype


def compute_sinusoidal_embedding(x: Tensor, embedding_dim: int) -> Tensor:
    """[*, 1] -> [*, 1, embedding_dim]: cos then sin of x * 10000^(-i/half), computed in fp32."""
    half = embedding_dim // 2
    freqs = torch.exp(
        (-math.log(10000) * torch.arange(0, half, dtype=torch.float32, device=x.device)) / half
    )
    angles = x.unsqueeze(1).float() * freqs.unsqueeze(0)
    return torch.cat([torch.cos(angles), torch.sin(angles)], dim=-1)


class RangeEncoder(fl.Chain):
    def __init__(
        self, sinusoidal_embedding_dim: int, embedding_dim: int, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.sinusoidal_embedding_dim, self.embedding_dim = sinusoidal_embedding_dim, embedding_dim
        super().__init__(
            fl.Lambda(self.compute_sinusoidal_embedding),
            fl.Converter(set_device=False, set_dtype=True),
            fl.Linear(sinusoidal_embedding_dim, embedding_dim, device=device, dtype=dtype),
            fl.SiLU(),
            fl.Linear(embedding_dim, embedding_dim, device=device, dtype=dtype),
        )

    def compute_sinusoidal_embedding(self, x: Tensor) -> Tensor:
        return compute_sinusoidal_embedding(x, embedding_dim=self.sinusoidal_embedding_dim)


class RangeAdapter2d(fl.Sum, Adapter[fl.Conv2d]):
    def __init__(
        self, target: fl.Conv2d, channels: int, embedding_dim: int, context_key: str,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.channels, self.embedding_dim = channels, embedding_dim
        with self.setup_adapter(target):
            super().__init__(
                target,
                fl.Chain(
                    fl.UseContext("range_adapter", context_key),
                    fl.SiLU(),
                    fl.Linear(embedding_dim, channels, device=device, dtype=dtype),
                    fl.Reshape(channels, 1, 1),
                ),
            )

    @property
    def context_key(self) -> str:
        source = self.ensure_find(fl.UseContext)
        assert source.context == "range_adapter"
        return source.key

    @context_key.setter
    def context_key(self, value: str) -> None:
        source = self.ensure_find(fl.UseContext)
        assert source.context == "range_adapter"
        source.key = value

    def forward(self, *inputs: Any) -> Any:
        if len(inputs) == 1 and isinstance(inputs[0], Tensor) and inputs[0].is_cuda and B.fusion_enabled():
            children = list(self)
            if len(children) == 2 and type(children[0]) is fl.Conv2d and isinstance(children[1], fl.Chain):
                conv, side = children
                hooked = [conv, side, *side]
                no_hooks = not any(m._forward_hooks or m._forward_pre_hooks for m in hooked)
                if no_hooks and B.conv_supported(conv) and len(side) == 4 and type(side[3]) is fl.Reshape:
                    # side chain: UseContext -> SiLU -> Linear -> Reshape(C,1,1); run all but the reshape
                    bias = inputs[0]
                    for layer in list(side)[:3]:
                        bias = layer(bias)
                    bias = bias.reshape(bias.shape[0], -1)  # what Reshape(C, 1, 1) would see, minus the 1x1
                    if bias.shape[0] in (1, inputs[0].shape[0]) and bias.shape[1] == conv.out_channels:
                        bias = bias.expand(inputs[0].shape[0], -1)
                        return B.conv2d_module(inputs[0], conv, chan_bias=bias)
        return fl.Sum.forward(self, *inputs)
