"""Adapter protocol: wrap a target module in a Chain and splice it into the tree.

Behavioural contract from /root/reference/src/refiners/fluxion/adapters/adapter.py:10-127
(`setup_adapter` :34, `inject` :53, `eject` :86, `lookup_top_adapter` :107).
"""

from __future__ import annotations

import contextlib
from typing import Any, Generic, Iterator, TypeVar

import refiners_b200.fluxion.layers as fl

T = TypeVar("T", bound=fl.Module)
TAdapter = TypeVar("TAdapter", bound="Adapter[Any]")


class Adapter(Generic[T]):
    """Mixin for Chains that adapt (wrap) a ``target`` module."""

    # list-wrapped so torch does not register the target twice
    _target: "list[T]"

    def __init_subclass__(cls, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)
        assert issubclass(cls, fl.Chain), f"Adapter {cls.__name__} must be a Chain"

    @property
    def target(self) -> T:
        return self._target[0]

    @contextlib.contextmanager
    def setup_adapter(self, target: T) -> Iterator[None]:
        """Context in which the adapter's Chain constructor must be called: the target keeps
        its current parent while it is being wrapped."""
        assert isinstance(self, fl.Chain)
        assert not hasattr(self, "_modules") or len(self) == 0, (
            "Call the Chain constructor in the setup_adapter context."
        )
        self._target = [target]
        if isinstance(target, fl.ContextModule):
            with target.no_parent_refresh():
                yield
        else:
            yield

    def inject(self: TAdapter, parent: fl.Chain | None = None) -> TAdapter:
        """Put the adapter where its target currently sits in ``parent`` (or in the target's
        own parent when it is known)."""
        assert isinstance(self, fl.Chain)
        target = self.target
        if parent is None and isinstance(target, fl.ContextModule):
            parent = target.parent
            if parent is not None:
                assert isinstance(parent, fl.Chain), f"{target} has invalid parent {parent}"
        inner_parent = self.find_parent(target)
        if parent is None:
            if isinstance(target, fl.ContextModule):
                target._set_parent(inner_parent)
            return self
        holder = parent.ensure_find_parent(target)
        holder.replace(old_module=target, new_module=self, old_module_parent=inner_parent)
        return self

    def eject(self) -> None:
        """Undo ``inject``: the target (or the outermost adapter stacked on it) takes the
        adapter's place."""
        assert isinstance(self, fl.Chain)
        successor = lookup_top_adapter(self, self.target)
        parent = self.parent
        if parent is None:
            if isinstance(successor, fl.ContextModule):
                successor._set_parent(None)
        else:
            parent.replace(old_module=self, new_module=successor)

    def _pre_structural_copy(self) -> None:
        if isinstance(self.target, fl.Chain):
            raise RuntimeError(f"Chain adapters ({self}) typically cannot be copied, eject them first.")

    def _post_structural_copy(self: TAdapter, source: TAdapter) -> None:
        self._target = [source.target]


def lookup_top_adapter(top: fl.Chain, target: fl.Module) -> fl.Module:
    """Outermost Adapter between ``target`` and ``top`` (exclusive), else ``target``."""
    holder = top.find_parent(target)
    if holder is None or holder is top:
        return target
    best: fl.Module = target
    node = holder
    while node is not top:
        if isinstance(node, Adapter):
            best = node
        assert node.parent, f"parent tree of {top} is broken"
        node = node.parent
    return best
