"""Adapters: Chains that wrap a ``target`` module and take its place in a model tree.

Contract (names, signatures, effects on the tree, error messages) from
/root/reference/src/refiners/fluxion/adapters/adapter.py:10-127; the reference's own adapter tests run against this
module (tests/test_reference_own_tests.py).

How this implementation thinks about it: an adapter is a Chain that contains its target somewhere below it.  Injecting
means *swapping positions*: the node that currently holds the target in the model tree gets the adapter instead, and
the target's parent pointer moves to the node that holds it inside the adapter.  Ejecting swaps back - except that
whatever was adapted *inside* this adapter in the meantime (a second adapter stacked on the same target) is what goes
back into the tree, so stacked adapters can be removed in any order.
"""

from __future__ import annotations

import contextlib
from typing import Any, Generic, Iterator, TypeVar

import refiners_b200.fluxion.layers as fl

T = TypeVar("T", bound=fl.Module)
TAdapter = TypeVar("TAdapter", bound="Adapter[Any]")


def _ancestors_below(top: fl.Chain, start: fl.Chain) -> Iterator[fl.Chain]:
    """``start`` and its parents up to, excluding, ``top``."""
    node: fl.Chain | None = start
    while node is not top:
        assert node is not None, f"parent tree of {top} is broken"
        yield node
        node = node.parent


def lookup_top_adapter(top: fl.Chain, target: fl.Module) -> fl.Module:
    """The outermost adapter that sits strictly between ``top`` and ``target``; ``target`` itself if there is none
    (or if ``top`` holds the target directly)."""
    holder = top.find_parent(target)
    if holder is None or holder is top:
        return target
    outermost: fl.Module = target
    for node in _ancestors_below(top, holder):
        if isinstance(node, Adapter):
            outermost = node
    return outermost


class Adapter(Generic[T]):
    """Mixin for Chains that adapt (wrap) a ``target`` module."""

    # kept in a one-element list: a plain attribute would make torch register the target as a sub-module twice
    _target: "list[T]"

    def __init_subclass__(cls, **kwargs: Any) -> None:
        super().__init_subclass__(**kwargs)
        assert issubclass(cls, fl.Chain), f"Adapter {cls.__name__} must be a Chain"

    @property
    def target(self) -> T:
        return self._target[0]

    @contextlib.contextmanager
    def setup_adapter(self, target: T) -> Iterator[None]:
        """To be entered by the adapter's constructor around its ``Chain.__init__`` call: records the target and keeps
        the target's parent pointer frozen while the adapter's own tree is assembled around it (the target still
        belongs to the model until ``inject``)."""
        assert isinstance(self, fl.Chain)
        already_built = hasattr(self, "_modules") and len(self) > 0
        assert not already_built, "Call the Chain constructor in the setup_adapter context."
        self._target = [target]
        freeze = target.no_parent_refresh() if isinstance(target, fl.ContextModule) else contextlib.nullcontext()
        with freeze:
            yield

    def inject(self: TAdapter, parent: fl.Chain | None = None) -> TAdapter:
        """Put the adapter where the target sits.  ``parent`` names the tree to look in when the target does not know
        its parent (weighted leaves do not) or has none."""
        assert isinstance(self, fl.Chain)
        target = self.target
        knows_parent = isinstance(target, fl.ContextModule)
        if parent is None and knows_parent:
            parent = target.parent
            assert parent is None or isinstance(parent, fl.Chain), f"{target} has invalid parent {parent}"
        inside = self.find_parent(target)  # the node of THIS adapter that holds the target
        if parent is None:
            # nothing to splice into: the adapter simply becomes the root above the target
            if knows_parent:
                target._set_parent(inside)
            return self
        # the node holding the target may be deeper than ``parent`` (two adapters built before either is injected)
        parent.ensure_find_parent(target).replace(old_module=target, new_module=self, old_module_parent=inside)
        return self

    def eject(self) -> None:
        """Undo ``inject``; the tree ends up as it was, minus this adapter."""
        assert isinstance(self, fl.Chain)
        comes_back = lookup_top_adapter(self, self.target)
        holder = self.parent
        if holder is not None:
            holder.replace(old_module=self, new_module=comes_back)
        elif isinstance(comes_back, fl.ContextModule):
            comes_back._set_parent(None)

    # -- structural copies (Chain.structural_copy calls these hooks) ---------------------------------
    def _pre_structural_copy(self) -> None:
        if isinstance(self.target, fl.Chain):
            raise RuntimeError(f"Chain adapters ({self}) typically cannot be copied, eject them first.")

    def _post_structural_copy(self: TAdapter, source: TAdapter) -> None:
        self._target = [source.target]
