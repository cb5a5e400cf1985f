"""CUDA-graph replay of a fluxion Chain.

The Chain walker stays pure Python (north star), but at ~2.7 k module calls and ~9.5 k context
registrations per SDXL forward the interpreter would cost as much as
the GPU work.  ``GraphedChain`` runs the walker ONCE under stream capture and replays the
recorded launch sequence afterwards.

Contexts: while a GraphedChain is attached, ``set_context`` calls on the chain (or on an
adapter wrapping it) are intercepted: tensor values are copied into static buffers that the
captured kernels read, so the usual API (``unet.set_timestep(t)`` ...) keeps working and costs a
device copy instead of a tree walk.  Non-tensor values are compared by equality.

The capture is dropped and redone when
  * the tree is edited anywhere (adapter inject/eject bump the structure epoch),
  * an input or context value changes shape, dtype or (for non-tensors) value,
  * ``invalidate()`` is called (do so after replacing parameter storages).
Python side effects of a forward (residual lists, size stacks) happen at capture time only and
are undone by ``Chain._reset_context`` as in eager mode.
"""

from __future__ import annotations

from typing import Any

import torch
from torch import Tensor

from refiners_b200 import backend as B
from refiners_b200.fluxion.layers import graph as _g
from refiners_b200.fluxion.layers.graph import Chain, structure_epoch


def _signature(v: Any) -> Any:
    if isinstance(v, Tensor):
        return ("tensor", tuple(v.shape), v.dtype, str(v.device))
    return ("value", repr(v))


class GraphedChain:
    def __init__(self, chain: Chain, warmup: int = 2) -> None:
        self.chain = chain
        self.warmup = warmup
        self._graph: torch.cuda.CUDAGraph | None = None
        self._epoch = -1
        self._pending: dict[tuple[int, str], dict[str, Any]] = {}   # (id(owner), context) -> values
        self._owners: dict[int, Chain] = {}
        self._static: dict[tuple[int, str, str], Tensor] = {}
        self._static_in: list[Tensor] = []
        self._static_out: Any = None
        self._sig: Any = None
        self._applying = False
        self.launches_per_replay = 0
        self.captures = 0
        self.replays = 0
        _g._context_listeners.append(self._on_set_context)

    def close(self) -> None:
        if self._on_set_context in _g._context_listeners:
            _g._context_listeners.remove(self._on_set_context)
        self._graph = None

    def invalidate(self) -> None:
        self._graph = None

    # -- context interception -----------------------------------------------------------------
    def _on_set_context(self, owner: Chain, context: str, value: Any) -> bool:
        if self._applying or not isinstance(value, dict):
            return False
        if owner is not self.chain and owner not in self.chain.get_parents():
            return False
        self._owners[id(owner)] = owner
        self._pending.setdefault((id(owner), context), {}).update(value)
        return True

    def _apply_contexts(self, use_static: bool) -> None:
        """Push the recorded context values (or their static mirrors) through the real API."""
        self._applying = True
        try:
            for (oid, context), values in self._pending.items():
                payload = {
                    key: (self._static[(oid, context, key)] if use_static and isinstance(v, Tensor) and v.is_cuda else v)
                    for key, v in values.items()
                }
                self._owners[oid].set_context(context, payload)
        finally:
            self._applying = False

    def _signature(self, inputs: tuple[Tensor, ...]) -> Any:
        ctx = tuple(
            (oid, context, key, _signature(v))
            for (oid, context), values in sorted(self._pending.items(), key=lambda kv: (kv[0][0], kv[0][1]))
            for key, v in sorted(values.items())
        )
        return (tuple(_signature(t) for t in inputs), ctx)

    def _capture(self, inputs: tuple[Tensor, ...]) -> None:
        self._static_in = [t.clone() for t in inputs]
        self._static = {
            (oid, context, key): v.clone()
            for (oid, context), values in self._pending.items()
            for key, v in values.items()
            if isinstance(v, Tensor) and v.is_cuda
        }
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):  # warm-up off the capture: packs weights, sizes the allocator
            for _ in range(self.warmup):
                self._apply_contexts(use_static=True)
                self.chain(*self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        self._apply_contexts(use_static=True)
        graph = torch.cuda.CUDAGraph()
        before = B.launch_count()
        with torch.cuda.graph(graph):
            self._static_out = self.chain(*self._static_in)
        self.launches_per_replay = B.launch_count() - before
        self._graph = graph
        self._epoch = structure_epoch()
        self.captures += 1

    def __call__(self, *inputs: Tensor) -> Any:
        sig = self._signature(inputs)
        if self._graph is None or self._epoch != structure_epoch() or sig != self._sig:
            self._capture(inputs)
            self._sig = sig
        else:
            for buf, new in zip(self._static_in, inputs):
                buf.copy_(new, non_blocking=True)
            for (oid, context), values in self._pending.items():
                for key, v in values.items():
                    if isinstance(v, Tensor) and v.is_cuda:
                        self._static[(oid, context, key)].copy_(v, non_blocking=True)
        assert self._graph is not None
        self._graph.replay()
        self.replays += 1
        return self._static_out
