"""Low-rank adapters.

Contract from /root/reference/src/refiners/fluxion/adapters/lora.py: `Lora` :14 (a Chain
``down -> up -> Multiply(scale)``; down ~ N(0, 1/rank), up = 0), `LinearLora` :181,
`Conv2dLora` :269, `LoraAdapter` :383 (``Sum(target, *loras)``), `auto_attach_loras` :479.

B200 addition: on CUDA a ``LoraAdapter`` around a ``Linear`` whose LoRAs are all
``LinearLora`` runs as ONE launch - ``y = x Wt + b + sum_i s_i (x A_it) B_it`` with the rank
terms accumulated in the base GEMM's epilogue (fp32) - instead of the reference's
base GEMM + 2 GEMMs + Multiply + add per LoRA.  The tree is untouched: the fusion is decided
per call from the adapter's current children.
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
    def __init__(
        self, name: str, /, rank: int = 16, scale: float = 1.0, device: Device | str | None = None,
        dtype: DType | None = None,
    ) -> None:
        self.name, self._rank, self._scale = name, rank, scale
        down, up = self.lora_layers(device=device, dtype=dtype)
        super().__init__(down, up, fl.Multiply(scale))
        self.reset_parameters()

    def reset_parameters(self) -> None:
        nn.init.normal_(self.down.weight, std=1 / self.rank)
        nn.init.zeros_(self.up.weight)

    @abstractmethod
    def lora_layers(self, device: Device | str | None = None, dtype: DType | None = None) -> tuple[T, T]: ...

    @abstractmethod
    def is_compatible(self, layer: fl.WeightedModule, /) -> bool: ...

    @property
    def down(self) -> T:
        layer = self[0]
        assert isinstance(layer, fl.WeightedModule)
        return cast(T, layer)

    @property
    def up(self) -> T:
        layer = self[1]
        assert isinstance(layer, fl.WeightedModule)
        return cast(T, layer)

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

    @classmethod
    def from_weights(cls, name: str, /, down: Tensor, up: Tensor) -> "Lora[Any]":
        if up.ndim == 2 and down.ndim == 2:
            return LinearLora.from_weights(name, up=up, down=down)
        if up.ndim == 4 and down.ndim == 4:
            return Conv2dLora.from_weights(name, up=up, down=down)
        raise ValueError(f"Unsupported weight shapes: up={up.shape}, down={down.shape}")

    @classmethod
    def from_dict(cls, name: str, /, state_dict: dict[str, Tensor]) -> dict[str, "Lora[Any]"]:
        """Build LoRAs from ``{<prefix>.down.weight, <prefix>.up.weight, ...}`` (pairs are taken
        in dict order: down first, then up)."""
        weights = [(k, v) for k, v in state_dict.items() if ".weight" in k]
        loras: dict[str, Lora[Any]] = {}
        for (down_key, down), (_, up) in zip(weights[::2], weights[1::2]):
            loras[".".join(down_key.split(".")[:-2])] = cls.from_weights(name, down=down, up=up)
        return loras

    def auto_attach(
        self, target: fl.Chain, include: list[str] | None = None, exclude: list[str] | None = None,
    ) -> "tuple[LoraAdapter, fl.Chain | None] | None":
        """Find the first compatible, not yet adapted layer of ``target``.  Returns the adapter
        to inject and the parent to inject it into (``None`` when this LoRA was appended to an
        existing adapter)."""
        for layer, parent in target.walk(type(self.up)):
            if isinstance(parent, Lora):
                continue
            if include is not None or exclude is not None:
                lineage = {type(p).__name__ for p in parent.get_parents() + [parent]}
                if include is not None and lineage.isdisjoint(include):
                    continue
                if exclude is not None and not lineage.isdisjoint(exclude):
                    continue
            if not self.is_compatible(layer):
                continue
            if isinstance(parent, LoraAdapter):
                if self.name in parent.names:
                    continue
                parent.add_lora(self)
                return parent, None
            return LoraAdapter(layer, self), parent
        return None

    def load_weights(self, down_weight: Tensor, up_weight: Tensor) -> None:
        assert down_weight.shape == self.down.weight.shape
        assert up_weight.shape == self.up.weight.shape
        self.down.weight = nn.Parameter(down_weight.to(device=self.device, dtype=self.dtype))
        self.up.weight = nn.Parameter(up_weight.to(device=self.device, dtype=self.dtype))


class LinearLora(Lora[fl.Linear]):
    def __init__(
        self, name: str, /, in_features: int, out_features: int, rank: int = 16, scale: float = 1.0,
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.in_features, self.out_features = in_features, out_features
        super().__init__(name, rank=rank, scale=scale, device=device, dtype=dtype)

    @classmethod
    def from_weights(cls, name: str, /, down: Tensor, up: Tensor) -> "LinearLora":
        assert up.ndim == 2 and down.ndim == 2
        assert down.shape[0] == up.shape[1], f"Rank mismatch: down rank={down.shape[0]} and up rank={up.shape[1]}"
        lora = cls(
            name,
            in_features=down.shape[1],
            out_features=up.shape[0],
            rank=down.shape[0],
            device=up.device,
            dtype=up.dtype,
        )
        lora.load_weights(down_weight=down, up_weight=up)
        return lora

    def lora_layers(self, device: Device | str | None = None, dtype: DType | None = None) -> tuple[fl.Linear, fl.Linear]:
        return (
            fl.Linear(self.in_features, self.rank, bias=False, device=device, dtype=dtype),
            fl.Linear(self.rank, self.out_features, bias=False, device=device, dtype=dtype),
        )

    def is_compatible(self, layer: fl.WeightedModule, /) -> bool:
        return (
            isinstance(layer, fl.Linear)
            and layer.in_features == self.in_features
            and layer.out_features == self.out_features
        )


class Conv2dLora(Lora[fl.Conv2d]):
    def __init__(
        self, name: str, /, in_channels: int, out_channels: int, rank: int = 16, scale: float = 1.0,
        kernel_size: tuple[int, int] = (1, 3), stride: tuple[int, int] = (1, 1), padding: tuple[int, int] = (0, 1),
        device: Device | str | None = None, dtype: DType | None = None,
    ) -> None:
        self.in_channels, self.out_channels, self.kernel_size, self.stride = in_channels, out_channels, kernel_size, stride
        self.padding = padding
        super().__init__(name, rank=rank, scale=scale, device=device, dtype=dtype)

    @classmethod
    def from_weights(cls, name: str, /, down: Tensor, up: Tensor) -> "Conv2dLora":
        assert up.ndim == 4 and down.ndim == 4
        assert down.shape[0] == up.shape[1], f"Rank mismatch: down rank={down.shape[0]} and up rank={up.shape[1]}"
        k_down, k_up = down.shape[2], up.shape[2]
        lora = cls(
            name,
            in_channels=down.shape[1],
            out_channels=up.shape[0],
            rank=down.shape[0],
            kernel_size=(k_down, k_up),
            padding=(1 if k_down == 3 else 0, 1 if k_up == 3 else 0),
            device=up.device,
            dtype=up.dtype,
        )
        lora.load_weights(down_weight=down, up_weight=up)
        return lora

    def lora_layers(self, device: Device | str | None = None, dtype: DType | None = None) -> tuple[fl.Conv2d, fl.Conv2d]:
        return (
            fl.Conv2d(
                self.in_channels,
                self.rank,
                kernel_size=self.kernel_size[0],
                stride=self.stride[0],
                padding=self.padding[0],
                use_bias=False,
                device=device,
                dtype=dtype,
            ),
            fl.Conv2d(
                self.rank,
                self.out_channels,
                kernel_size=self.kernel_size[1],
                stride=self.stride[1],
                padding=self.padding[1],
                use_bias=False,
                device=device,
                dtype=dtype,
            ),
        )

    def is_compatible(self, layer: fl.WeightedModule, /) -> bool:
        if (
            isinstance(layer, fl.Conv2d)
            and layer.in_channels == self.in_channels
            and layer.out_channels == self.out_channels
        ):
            # the down projection inherits the stride of the layer it adapts
            self.down.stride = layer.stride
            return True
        return False


class LoraAdapter(fl.Sum, Adapter[fl.WeightedModule]):
    """``target(x) + sum_i lora_i(x)``."""

    def __init__(self, target: fl.WeightedModule, /, *loras: Lora[Any]) -> None:
        with self.setup_adapter(target):
            super().__init__(target, *loras)

    @property
    def lora_layers(self) -> Iterator[Lora[Any]]:
        return cast(Iterator[Lora[Any]], self.layers(Lora))

    @property
    def names(self) -> list[str]:
        return [lora.name for lora in self.lora_layers]

    @property
    def loras(self) -> dict[str, Lora[Any]]:
        return {lora.name: lora for lora in self.lora_layers}

    @property
    def scales(self) -> dict[str, float]:
        return {lora.name: lora.scale for lora in self.lora_layers}

    @scales.setter
    def scale(self, values: dict[str, float]) -> None:
        for name, value in values.items():
            self.loras[name].scale = value

    def add_lora(self, lora: Lora[Any], /) -> None:
        assert lora.name not in self.names, f"LoRA layer with name {lora.name} already exists"
        self.append(lora)

    def remove_lora(self, name: str, /) -> Lora[Any] | None:
        if name in self.names:
            lora = self.loras[name]
            self.remove(lora)
            return lora
        return None

    # -- fused execution --------------------------------------------------------------------
    def _fusable(self) -> list[tuple[Tensor, Tensor, float]] | None:
        """LoRA (A, B, scale) triples when the whole Sum can be one GEMM launch, else None."""
        children = list(self)
        base = children[0]
        if type(base) is not fl.Linear or base._forward_hooks or base._forward_pre_hooks:
            return None
        triples: list[tuple[Tensor, Tensor, float]] = []
        for child in children[1:]:
            if not isinstance(child, LinearLora) or len(child) != 3:
                return None
            down, up, mul = child[0], child[1], child[2]
            if type(down) is not fl.Linear or type(up) is not fl.Linear or type(mul) is not fl.Multiply:
                return None
            if mul.bias != 0.0 or down.bias is not None or up.bias is not None:
                return None
            if any(m._forward_hooks or m._forward_pre_hooks for m in (child, down, up, mul)):
                return None
            triples.append((down.weight, up.weight, float(mul.scale)))
        return triples

    def _forward_with_residual(self, inputs: tuple[Any, ...], residual: Tensor | None) -> Any:
        """One-launch evaluation (optionally adding a skip connection); NotImplemented if the
        adapter's current children do not allow it."""
        if len(inputs) == 1 and isinstance(inputs[0], Tensor) and inputs[0].is_cuda and B.fusion_enabled():
            x = inputs[0]
            triples = self._fusable()
            if triples is not None and B.lora_fusable(x, triples):
                base = cast(fl.Linear, self[0])
                if residual is not None and (x.shape[:-1] != residual.shape[:-1] or base.out_features != residual.shape[-1]):
                    return NotImplemented
                return B.linear(x, base.weight, base.bias, loras=triples, residual=residual)
        return NotImplemented

    def forward(self, *inputs: Any) -> Any:
        fused = self._forward_with_residual(inputs, None)
        if fused is not NotImplemented:
            return fused
        return fl.Sum.forward(self, *inputs)


def _auto_attach_loras(
    loras: dict[str, Lora[Any]], target: fl.Chain, /, include: list[str] | None = None,
    exclude: list[str] | None = None, debug_map: list[tuple[str, str]] | None = None,
) -> list[str]:
    failed: list[str] = []
    for key, lora in loras.items():
        attached = lora.auto_attach(target, include=include, exclude=exclude)
        if attached is None:
            failed.append(key)
            continue
        adapter, parent = attached
        if parent is None:
            if debug_map is not None:
                debug_map.append((key, adapter.get_path()))
            continue
        if debug_map is not None:
            debug_map.append((key, adapter.target.get_path(parent)))
        adapter.inject(parent)
    return failed


def auto_attach_loras(
    loras: dict[str, Lora[Any]], target: fl.Chain, /, include: list[str] | None = None,
    exclude: list[str] | None = None, sanity_check: bool = True, debug_map: list[tuple[str, str]] | None = None,
) -> list[str]:
    """Attach each LoRA to the first compatible layer of ``target``; returns the keys that
    found no home.  With ``sanity_check`` a second pass with copies must attach nothing."""
    if not sanity_check:
        return _auto_attach_loras(loras, target, include=include, exclude=exclude, debug_map=debug_map)

    twins = {key: Lora.from_weights(lora.name, lora.down.weight, lora.up.weight) for key, lora in loras.items()}
    first_map: list[tuple[str, str]] = []
    first_failed = _auto_attach_loras(loras, target, include=include, exclude=exclude, debug_map=first_map)
    if debug_map is not None:
        debug_map += first_map
    if len(first_map) != len(loras) or first_failed:
        raise ValueError(
            f"sanity check failed: {len(first_map)} / {len(loras)} LoRA layers attached, {len(first_failed)} failed"
        )
    second_map: list[tuple[str, str]] = []
    second_failed = _auto_attach_loras(twins, target, include=include, exclude=exclude, debug_map=second_map)
    if second_map or len(second_failed) != len(loras):
        raise ValueError(
            f"sanity check failed: {len(second_map)} / {len(loras)} LoRA layers attached twice, {len(second_failed)} skipped"
        )
    return first_failed
