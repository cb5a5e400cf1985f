"""Low-rank adapters.

Contract (class names, constructor signatures, children order ``down -> up -> Multiply(scale)``, initialisation
down ~ N(0, 1/rank) and up = 0, state-dict keys, the ``auto_attach`` search order and the error texts) from
/root/reference/src/refiners/fluxion/adapters/lora.py: `Lora` :14-178, `LinearLora` :181-266, `Conv2dLora` :269-380,
`LoraAdapter` :383-448, `auto_attach_loras` :479-536.  The reference's tests/adapters/test_lora.py runs against this
module (tests/test_reference_own_tests.py).

B200 execution: on CUDA a ``LoraAdapter`` around a ``Linear`` whose LoRAs are all ``LinearLora`` is ONE GEMM launch -
against the cached merged weight ``W + sum_i s_i B_i A_i`` (default; the factors and scales are step-invariant), or,
with merging switched off, the base GEMM with the rank terms appended as extra k-blocks after one rank-space GEMM
(refiners_b200.backend.linear).  The tree is never rewritten: the decision is taken per call from the adapter's
current children, and forward hooks anywhere on the path switch it off.
"""

from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Any, Generic, Iterator, TypeVar, cast

import torch
from torch import Tensor, nn

import refiners_b200.fluxion.layers as fl
from refiners_b200 import backend as B
from refiners_b200.fluxion.adapters.adapter import Adapter

T = TypeVar("T", bound=fl.WeightedModule)
Device = torch.device
DType = torch.dtype


class Lora(Generic[T], fl.Chain, ABC):
    """``x -> up(down(x)) * scale`` with a pair of weight-only layers of the adapted kind."""

    def __init__(
        self, name: str, /, rank: int = 16, scale: float = 1.0, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.name = name
        self._rank = rank
        self._scale = scale
        super().__init__(*self.lora_layers(device=device, dtype=dtype), fl.Multiply(scale))
        self.reset_parameters()

    # -- what a concrete kind provides ----------------------------------------------------------------
    @abstractmethod
    def lora_layers(self, device: Device | str | None = None, dtype: DType | None = None) -> tuple[T, T]:
        """The (down, up) pair, without biases."""

    @abstractmethod
    def is_compatible(self, layer: fl.WeightedModule, /) -> bool:
        """Whether this LoRA can sit beside ``layer``."""

    # -- the pair -------------------------------------------------------------------------------------
    def _factor(self, index: int) -> T:
        layer = self[index]
        assert isinstance(layer, fl.WeightedModule)
        return cast(T, layer)

    @property
    def down(self) -> T:
        return self._factor(0)

    @property
    def up(self) -> T:
        return self._factor(1)

    @property
    def rank(self) -> int:
        return self._rank

    @property
    def scale(self) -> float:
        return self._scale

    @scale.setter
    def scale(self, value: float) -> None:
        self._scale = value
        self.ensure_find(fl.Multiply).scale = value

    def reset_parameters(self) -> None:
        """A fresh LoRA is a no-op: random down-projection, zero up-projection."""
        nn.init.normal_(self.down.weight, std=1 / self.rank)
        nn.init.zeros_(self.up.weight)

    def load_weights(self, down_weight: Tensor, up_weight: Tensor) -> None:
        for layer, weight in ((self.down, down_weight), (self.up, up_weight)):
            assert weight.shape == layer.weight.shape
            layer.weight = nn.Parameter(weight.to(device=self.device, dtype=self.dtype))

    # -- construction from checkpoints ------------------------------------------------------------------
    @classmethod
    def from_weights(cls, name: str, /, down: Tensor, up: Tensor) -> "Lora[Any]":
        kinds = {2: LinearLora, 4: Conv2dLora}
        if down.ndim != up.ndim or down.ndim not in kinds:
            raise ValueError(f"Unsupported weight shapes: up={up.shape}, down={down.shape}")
        return kinds[down.ndim].from_weights(name, up=up, down=down)

    @classmethod
    def from_dict(cls, name: str, /, state_dict: dict[str, Tensor]) -> dict[str, "Lora[Any]"]:
        """``{<path>.down.weight: ..., <path>.up.weight: ...}`` -> ``{<path>: Lora}``.  Keys come in (down, up) pairs in
        dictionary order, as the reference expects of its checkpoints."""
        entries = [(key, tensor) for key, tensor in state_dict.items() if ".weight" in key]
        built: dict[str, Lora[Any]] = {}
        for (down_key, down), (_, up) in zip(entries[0::2], entries[1::2]):
            path = down_key.rsplit(".", 2)[0]  # strip ".down.weight"
            built[path] = cls.from_weights(name, down=down, up=up)
        return built

    # -- finding a home -------------------------------------------------------------------------------------
    def auto_attach(
        self, target: fl.Chain, include: list[str] | None = None, exclude: list[str] | None = None,
    ) -> "tuple[LoraAdapter, fl.Chain | None] | None":
        """Search ``target`` (depth first) for the first layer of the adapted kind that is compatible and not yet
        carrying a LoRA of this name.  Returns ``(adapter, parent)``: a new adapter to inject into ``parent``, or an
        existing adapter this LoRA was appended to (``parent`` is then None); None if nothing fits."""

        def lineage_ok(parent: fl.Chain) -> bool:
            if include is None and exclude is None:
                return True
            names = {type(node).__name__ for node in (*parent.get_parents(), parent)}
            wanted = include is None or not names.isdisjoint(include)
            banned = exclude is not None and not names.isdisjoint(exclude)
            return wanted and not banned

        for layer, parent in target.walk(type(self.up)):
            if isinstance(parent, Lora) or not lineage_ok(parent) or not self.is_compatible(layer):
                continue  # a LoRA's own factors are never adapted
            if not isinstance(parent, LoraAdapter):
                return LoraAdapter(layer, self), parent
            if self.name not in parent.names:
                parent.add_lora(self)
                return parent, None
        return None


class LinearLora(Lora[fl.Linear]):
    def __init__(
        self, name: str, /, in_features: int, out_features: int, rank: int = 16, scale: float = 1.0,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.in_features = in_features
        self.out_features = out_features
        super().__init__(name, rank=rank, scale=scale, device=device, dtype=dtype)

    @classmethod
    def from_weights(cls, name: str, /, down: Tensor, up: Tensor) -> "LinearLora":
        assert up.ndim == 2 and down.ndim == 2
        rank, in_features = down.shape
        out_features, up_rank = up.shape
        assert rank == up_rank, f"Rank mismatch: down rank={rank} and up rank={up_rank}"
        lora = cls(name, in_features=in_features, out_features=out_features, rank=rank, device=up.device, dtype=up.dtype)
        lora.load_weights(down_weight=down, up_weight=up)
        return lora

    def lora_layers(self, device: Device | str | None = None, dtype: DType | None = None) -> tuple[fl.Linear, fl.Linear]:
        widths = (self.in_features, self.rank, self.out_features)
        down, up = (fl.Linear(a, b, bias=False, device=device, dtype=dtype) for a, b in zip(widths, widths[1:]))
        return down, up

    def is_compatible(self, layer: fl.WeightedModule, /) -> bool:
        return isinstance(layer, fl.Linear) and (layer.in_features, layer.out_features) == (self.in_features, self.out_features)


class Conv2dLora(Lora[fl.Conv2d]):
    """Per-factor geometry comes as (down, up) pairs; the default is a 1x1 down- and a 3x3 up-convolution."""

    def __init__(
        self, name: str, /, in_channels: int, out_channels: int, rank: int = 16, scale: float = 1.0,
        kernel_size: tuple[int, int] = (1, 3), stride: tuple[int, int] = (1, 1), padding: tuple[int, int] = (0, 1),
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        super().__init__(name, rank=rank, scale=scale, device=device, dtype=dtype)

    @classmethod
    def from_weights(cls, name: str, /, down: Tensor, up: Tensor) -> "Conv2dLora":
        assert up.ndim == 4 and down.ndim == 4
        assert down.shape[0] == up.shape[1], f"Rank mismatch: down rank={down.shape[0]} and up rank={up.shape[1]}"
        kernels = (down.shape[2], up.shape[2])
        lora = cls(
            name, in_channels=down.shape[1], out_channels=up.shape[0], rank=down.shape[0], kernel_size=kernels,
            padding=tuple(1 if k == 3 else 0 for k in kernels),  # type: ignore[arg-type]
            device=up.device, dtype=up.dtype,
        )
        lora.load_weights(down_weight=down, up_weight=up)
        return lora

    def lora_layers(self, device: Device | str | None = None, dtype: DType | None = None) -> tuple[fl.Conv2d, fl.Conv2d]:
        channels = (self.in_channels, self.rank, self.out_channels)
        down, up = (
            fl.Conv2d(channels[i], channels[i + 1], kernel_size=self.kernel_size[i], stride=self.stride[i],
                      padding=self.padding[i], use_bias=False, device=device, dtype=dtype)
            for i in (0, 1)
        )
        return down, up

    def is_compatible(self, layer: fl.WeightedModule, /) -> bool:
        fits = isinstance(layer, fl.Conv2d) and (layer.in_channels, layer.out_channels) == (self.in_channels, self.out_channels)
        if fits:
            self.down.stride = layer.stride  # the down-projection strides like the convolution it shadows
        return fits


class LoraAdapter(fl.Sum, Adapter[fl.WeightedModule]):
    """``target(x) + sum_i lora_i(x)``; LoRAs are addressed by name."""

    def __init__(self, target: fl.WeightedModule, /, *loras: Lora[Any]) -> None:
        with self.setup_adapter(target):
            super().__init__(target, *loras)

    @property
    def lora_layers(self) -> Iterator[Lora[Any]]:
        return cast(Iterator[Lora[Any]], self.layers(Lora))

    @property
    def loras(self) -> dict[str, Lora[Any]]:
        return {lora.name: lora for lora in self.lora_layers}

    @property
    def names(self) -> list[str]:
        return [*self.loras]

    @property
    def scales(self) -> dict[str, float]:
        return {name: lora.scale for name, lora in self.loras.items()}

    # the reference exposes the per-name setter under the attribute ``scale`` (its getter returns ``scales``)
    @property
    def scale(self) -> dict[str, float]:
        return self.scales

    @scale.setter
    def scale(self, values: dict[str, float]) -> None:
        mine = self.loras
        for name, value in values.items():
            mine[name].scale = value

    def add_lora(self, lora: Lora[Any], /) -> None:
        assert lora.name not in self.names, f"LoRA layer with name {lora.name} already exists"
        self.append(lora)

    def remove_lora(self, name: str, /) -> Lora[Any] | None:
        lora = self.loras.get(name)
        if lora is not None:
            self.remove(lora)
        return lora

    # -- single-GEMM evaluation on CUDA ------------------------------------------------------------
    def _fusable(self) -> list[tuple[Tensor, Tensor, float]] | None:
        """(A, B, scale) of every LoRA when the whole Sum is expressible as one GEMM, else None: the target must be a
        hook-free ``Linear`` and every other child a hook-free ``LinearLora`` in its pristine down / up / Multiply form."""
        base, *others = self
        if type(base) is not fl.Linear or base._forward_hooks or base._forward_pre_hooks:
            return None
        triples: list[tuple[Tensor, Tensor, float]] = []
        for lora in others:
            if not isinstance(lora, LinearLora) or len(lora) != 3:
                return None
            down, up, mul = lora
            pristine = (
                type(down) is fl.Linear and type(up) is fl.Linear and type(mul) is fl.Multiply
                and down.bias is None and up.bias is None and mul.bias == 0.0
                and not any(m._forward_hooks or m._forward_pre_hooks for m in (lora, down, up, mul))
            )
            if not pristine:
                return None
            triples.append((down.weight, up.weight, float(mul.scale)))
        return triples

    def _forward_with_residual(self, inputs: tuple[Any, ...], residual: Tensor | None) -> Any:
        """``self(x) (+ residual)`` as one launch, or NotImplemented when the current children / inputs do not allow it."""
        if len(inputs) != 1 or not isinstance(inputs[0], Tensor) or not inputs[0].is_cuda or not B.fusion_enabled():
            return NotImplemented
        x = inputs[0]
        triples = self._fusable()
        if triples is None or not B.lora_fusable(x, triples):
            return NotImplemented
        base = cast(fl.Linear, self[0])
        if residual is not None and (x.shape[:-1] != residual.shape[:-1] or base.out_features != residual.shape[-1]):
            return NotImplemented
        return B.linear(x, base.weight, base.bias, loras=triples, residual=residual)

    def forward(self, *inputs: Any) -> Any:
        fused = self._forward_with_residual(inputs, None)
        return fl.Sum.forward(self, *inputs) if fused is NotImplemented else fused


def _attach_each(
    loras: dict[str, Lora[Any]], target: fl.Chain, include: list[str] | None, exclude: list[str] | None,
    debug_map: list[tuple[str, str]] | None,
) -> list[str]:
    """One attachment pass; returns the keys that found no home and logs (key, path) of the others."""
    homeless: list[str] = []
    for key, lora in loras.items():
        found = lora.auto_attach(target, include=include, exclude=exclude)
        if found is None:
            homeless.append(key)
            continue
        adapter, parent = found
        if debug_map is not None:
            debug_map.append((key, adapter.get_path() if parent is None else adapter.target.get_path(parent)))
        if parent is not None:
            adapter.inject(parent)
    return homeless


def auto_attach_loras(
    loras: dict[str, Lora[Any]], target: fl.Chain, /, include: list[str] | None = None,
    exclude: list[str] | None = None, sanity_check: bool = True, debug_map: list[tuple[str, str]] | None = None,
) -> list[str]:
    """Attach every LoRA to the first compatible layer of ``target`` (see ``Lora.auto_attach``); returns the keys that
    could not be attached.  With ``sanity_check`` all of them must attach, and a second pass with twin LoRAs of the same
    names must then attach nothing (each name has taken every place it could)."""
    if not sanity_check:
        return _attach_each(loras, target, include, exclude, debug_map)

    twins = {key: Lora.from_weights(lora.name, lora.down.weight, lora.up.weight) for key, lora in loras.items()}
    placed: list[tuple[str, str]] = []
    homeless = _attach_each(loras, target, include, exclude, placed)
    if debug_map is not None:
        debug_map.extend(placed)
    if homeless or len(placed) != len(loras):
        raise ValueError(f"sanity check failed: {len(placed)} / {len(loras)} LoRA layers attached, {len(homeless)} failed")

    placed_again: list[tuple[str, str]] = []
    skipped = _attach_each(twins, target, include, exclude, placed_again)
    if placed_again or len(skipped) != len(loras):
        raise ValueError(
            f"sanity check failed: {len(placed_again)} / {len(loras)} LoRA layers attached twice, {len(skipped)} skipped"
        )
    return homeless
