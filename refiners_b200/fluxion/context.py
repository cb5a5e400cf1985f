"""Context store shared between a Chain and its sub-chains.

Mirrors the behaviour of the reference's ``ContextProvider``
(/root/reference/src/refiners/fluxion/context.py:9-72): a provider owns a mapping
``name -> dict``; when a parent pushes its contexts into a child, *missing* entries
are adopted by reference (so parent and child mutate the very same dict) and
*existing* entries are updated key by key.
"""

from typing import Any

from torch import Tensor

Context = dict[str, Any]
Contexts = dict[str, Context]


class ContextProvider:
    __slots__ = ("contexts",)

    def __init__(self) -> None:
        self.contexts: Contexts = {}

    @staticmethod
    def create(contexts: Contexts) -> "ContextProvider":
        provider = ContextProvider()
        provider.update_contexts(contexts)
        return provider

    def set_context(self, key: str, value: Context) -> None:
        self.contexts[key] = value

    def get_context(self, key: str) -> Any:
        return self.contexts.get(key)

    def update_contexts(self, new_contexts: Contexts) -> None:
        mine = self.contexts
        for name, incoming in new_contexts.items():
            current = mine.get(name)
            if current is None:
                mine[name] = incoming  # adopt by reference: shared with the giver
            elif current is not incoming:
                current.update(incoming)

    def __repr__(self) -> str:
        def show(v: Any) -> str:
            if isinstance(v, Tensor):
                return f"Tensor(shape={v.shape}, dtype={v.dtype}, device={v.device})"
            return repr(v)

        body = {name: {k: show(v) for k, v in ctx.items()} for name, ctx in self.contexts.items()}
        return f"{type(self).__name__}(contexts={body})"
