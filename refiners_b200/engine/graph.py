"""CUDA-graph replay of a fluxion Chain.

The Chain walker stays pure Python (north star), but at ~2.7 k module calls and ~9.5 k context
registrations per SDXL forward the interpreter would cost as much as
the GPU work.  ``GraphedChain`` runs the walker ONCE under stream capture and replays the
recorded launch sequence afterwards.

What a capture bakes in, and how each is kept honest:

  * inputs and TENSOR context values   live in static device buffers; every call copies the current
                                       values in (CPU tensors too: they are mirrored on the chain's
                                       device, so a CPU ``timestep`` follows its value instead of being
                                       frozen into the capture)
  * non-tensor context values          compared by ``repr``; a change re-captures
  * the module tree                    ``structure_epoch()``: any edit anywhere (adapter inject / eject,
                                       append, replace ...) re-captures
  * Python scalars of modules          ``value_epoch()``: assigning a public int / float / bool / str /
                                       sequence attribute on any fluxion module (``Multiply.scale``
                                       behind ``Lora.scale``, ``ip_adapter.scale``, ControlNet /
                                       ControlLora ``scale`` ...) re-captures - these floats become
                                       kernel arguments and packed column scales
  * parameter storage                  a fingerprint of every parameter / buffer (address, version,
                                       shape, dtype): ``load_state_dict``, ``.to()``, in-place edits
                                       and swapped ``.weight`` objects re-capture (the eager pack caches
                                       use the same stamp); writes through ``.data`` are invisible to
                                       autograd's version counter - call ``invalidate()`` after those.

Step-invariant hoisting.  Some context tensors do not change from one denoising step to
the next: the text embedding, the IP-Adapter image embedding, a ControlNet / ControlLora condition image.  Everything
computed from them alone - the K / V projections of every cross-attention (cross_attention.py:52-65 of the reference),
the IP-Adapter's K' / V' (image_prompt.py:243-262), the ConditionEncoder (control_lora.py:193-201) - is computed ONCE,
outside the captured graph (backend.InvariantMemo), and recomputed only when the caller passes a different tensor (or
the same tensor with a bumped version) for such a context.  ``RB200_HOIST=0`` keeps everything inside the graph.

``set_context`` is observed, not swallowed: the providers keep receiving the values, so an eager call of
the same chain (debugging, a CPU fallback of the caller) still sees its contexts.  Python side effects
of a forward (residual lists, size stacks) happen at capture time only and are undone by
``Chain._reset_context`` as in eager mode.
"""

from __future__ import annotations

import os
import weakref
from typing import Any, Callable

import torch
from torch import Tensor

from refiners_b200 import backend as B
from refiners_b200.fluxion.layers import graph as _g
from refiners_b200.fluxion.layers.base import value_epoch
from refiners_b200.fluxion.layers.graph import Chain, structure_epoch


def step_invariant(context: str, key: str) -> bool:
    """Default classification of context entries that stay the same over a denoising loop."""
    return key in ("clip_text_embedding", "clip_image_embedding") or key.startswith("condition")


def _stamp_of(t: Tensor) -> tuple[int, int, tuple[int, ...]]:
    return (t.data_ptr(), t._version, tuple(t.shape))


def _signature(v: Any, nested: bool = False) -> Any:
    """What a capture depends on in a context value.  A tensor stored directly is mirrored in a static buffer (any tensor of
    that shape and dtype can be copied in); tensors INSIDE a container (the tuple of T2I-Adapter feature maps, a list of
    residuals) are used by the captured kernels at their own addresses, so they count by identity and version; everything
    else by ``repr``."""
    if isinstance(v, Tensor):
        return ("tensor", tuple(v.shape), v.dtype, *(_stamp_of(v) if nested else ()))
    if isinstance(v, (tuple, list)):
        return (type(v).__name__, tuple(_signature(item, nested=True) for item in v))
    return ("value", repr(v))


class GraphedChain:
    def __init__(self, chain: Chain, warmup: int = 2, invariant: Callable[[str, str], bool] | None = step_invariant) -> None:
        self.chain = chain
        self.warmup = warmup
        self.invariant = invariant if os.environ.get("RB200_HOIST", "1") != "0" else None
        self._memo: B.InvariantMemo | None = None
        self._seen: dict[tuple[int, str, str], Any] = {}   # identity stamps of the invariant context tensors last copied
        self.hoisted_ops = 0
        self.refreshes = 0
        self._graph: torch.cuda.CUDAGraph | None = None
        self._epochs: tuple[int, int] = (-1, -1)
        self._pending: dict[tuple[int, str], dict[str, Any]] = {}   # (id(owner), context) -> values
        self._owners: dict[int, Chain] = {}
        self._static: dict[tuple[int, str, str], Tensor] = {}
        self._static_in: list[Tensor] = []
        self._static_out: Any = None
        self._sig: Any = None
        self._tensors: list[Tensor] = []
        self._fingerprint: Any = None
        self._applying = False
        self.launches_per_replay = 0
        self.captures = 0
        self.replays = 0
        # the module-global listener list must not keep this runner (and, through it, the model and its
        # device memory) alive: it holds a weak reference and unregisters itself once the runner is gone
        ref = weakref.ref(self)

        def listener(owner: Chain, context: str, value: Any) -> bool:
            me = ref()
            if me is None:
                if listener in _g._context_listeners:
                    _g._context_listeners.remove(listener)
                return False
            me._observe(owner, context, value)
            return False  # never consumed: the provider is updated as usual

        self._listener = listener
        _g._context_listeners.append(listener)

    def close(self) -> None:
        if self._listener in _g._context_listeners:
            _g._context_listeners.remove(self._listener)
        self._graph = None
        self._static.clear()
        self._static_in = []
        self._static_out = None

    def invalidate(self) -> None:
        self._graph = None

    # -- context observation ------------------------------------------------------------------
    def _observe(self, owner: Chain, context: str, value: Any) -> None:
        if self._applying or not isinstance(value, dict):
            return
        if owner is not self.chain and owner not in self.chain.get_parents():
            return
        self._owners[id(owner)] = owner
        self._pending.setdefault((id(owner), context), {}).update(value)

    def _apply_contexts(self) -> None:
        """Hand the static mirrors of the recorded tensor values (and the other values as they are) to the
        real providers, so that the walker under capture reads the buffers the replays will refresh."""
        self._applying = True
        try:
            for (oid, context), values in self._pending.items():
                payload = {key: (self._static[(oid, context, key)] if isinstance(v, Tensor) else v) for key, v in values.items()}
                self._owners[oid].set_context(context, payload)
        finally:
            self._applying = False

    def _restore_contexts(self) -> None:
        """After a capture: give the providers the caller's own values back."""
        self._applying = True
        try:
            for (oid, context), values in self._pending.items():
                self._owners[oid].set_context(context, dict(values))
        finally:
            self._applying = False

    def _signature(self, inputs: tuple[Tensor, ...]) -> Any:
        ctx = tuple(
            (oid, context, key, _signature(v))
            for (oid, context), values in sorted(self._pending.items(), key=lambda kv: (kv[0][0], kv[0][1]))
            for key, v in sorted(values.items())
        )
        return (tuple((_signature(t), str(t.device)) for t in inputs), ctx)

    def _stamp(self) -> Any:
        return tuple((t.data_ptr(), t._version) for t in self._tensors)

    def _capture(self, inputs: tuple[Tensor, ...]) -> None:
        device = inputs[0].device
        self._static_in = [t.clone() for t in inputs]
        self._static = {
            (oid, context, key): v.detach().to(device, copy=True)
            for (oid, context), values in self._pending.items()
            for key, v in values.items()
            if isinstance(v, Tensor)
        }
        self._memo, self._seen = None, {}
        if self.invariant is not None:
            self._memo = B.InvariantMemo()
            for (oid, context, key), buf in self._static.items():
                if self.invariant(context, key):
                    B.InvariantMemo.mark(buf)
                    self._seen[(oid, context, key)] = _stamp_of(self._pending[(oid, context)][key])
        previous_memo = B.set_invariant_memo(self._memo)
        try:
            side = torch.cuda.Stream(device=device)
            side.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(side):  # warm-up off the capture: packs weights, sizes the allocator, fills the memo
                for _ in range(self.warmup):
                    self._apply_contexts()
                    self.chain(*self._static_in)
            torch.cuda.current_stream(device).wait_stream(side)
            self._apply_contexts()
            graph = torch.cuda.CUDAGraph()
            before = B.launch_count()
            with torch.cuda.graph(graph):
                self._static_out = self.chain(*self._static_in)
            self.launches_per_replay = B.launch_count() - before
        finally:
            B.set_invariant_memo(previous_memo)
        self.hoisted_ops = len(self._memo) if self._memo is not None else 0
        self._restore_contexts()
        self._graph = graph
        self._epochs = (structure_epoch(), value_epoch())
        self._tensors = [*self.chain.parameters(), *self.chain.buffers()]
        self._fingerprint = self._stamp()
        self.captures += 1

    def __call__(self, *inputs: Tensor) -> Any:
        sig = self._signature(inputs)
        stale = (
            self._graph is None
            or self._epochs != (structure_epoch(), value_epoch())
            or sig != self._sig
            or self._fingerprint != self._stamp()
        )
        if stale:
            self._capture(inputs)
            self._sig = sig
        else:
            for buf, new in zip(self._static_in, inputs):
                buf.copy_(new, non_blocking=True)
            changed = False
            for (oid, context), values in self._pending.items():
                for key, v in values.items():
                    if not isinstance(v, Tensor):
                        continue
                    slot = (oid, context, key)
                    if slot in self._seen:  # step-invariant: copied (and its dependants recomputed) only when it changed
                        stamp = _stamp_of(v)
                        if stamp == self._seen[slot]:
                            continue
                        self._seen[slot] = stamp
                        changed = True
                    self._static[slot].copy_(v, non_blocking=True)
            if changed and self._memo is not None:
                self._memo.refresh()
                self.refreshes += 1
        assert self._graph is not None
        self._graph.replay()
        self.replays += 1
        return self._static_out
