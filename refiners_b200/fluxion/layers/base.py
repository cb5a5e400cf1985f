"""Module base classes of the fluxion mirror.

API surface follows /root/reference/src/refiners/fluxion/layers/module.py
(`Module` :23, `ContextModule` :153, `WeightedModule` :238, `ModuleTree` :267); the
implementation is written from scratch. Class ``__name__``s and parameter names are part
of the state-dict key contract (chain.py:19-38 in the reference), so they are kept.
"""

from __future__ import annotations

import contextlib
import inspect
import sys
from pathlib import Path
from typing import TYPE_CHECKING, Any, Iterator, Sequence, TypeVar

import torch
from torch import Tensor

from refiners_b200.fluxion.context import Context, ContextProvider
from refiners_b200.fluxion.utils import load_from_safetensors

if TYPE_CHECKING:
    from refiners_b200.fluxion.layers.graph import Chain

T = TypeVar("T", bound="Module")
_BASIC = (str, float, int, bool)
_PLAIN = (float, int, bool, str, list, tuple)

# Bumped whenever a public plain-Python attribute (a scale, an eps, a flag ...) is assigned on any fluxion
# module.  Such values end up as kernel arguments and inside packed weights; a captured CUDA graph
# (refiners_b200.engine.graph) compares this counter to know that its baked-in copies are stale.
_value_epoch = 0
_UNSET = object()


def value_epoch() -> int:
    return _value_epoch


def _is_basic(value: Any) -> bool:
    if isinstance(value, _BASIC):
        return True
    return isinstance(value, Sequence) and not isinstance(value, str) and all(isinstance(v, _BASIC) for v in value)


class Module(torch.nn.Module):
    """torch.nn.Module with tree-style printing and a (device, dtype) ``to``."""

    _tag: str = ""

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)

    def __setattr__(self, name: str, value: Any) -> None:
        if isinstance(value, _PLAIN) and name[:1] != "_":
            held = self.__dict__.get(name, _UNSET)
            if held is _UNSET and isinstance(getattr(type(self), name, None), property):
                try:  # a property (adapter.scale forwarding to its layers): compare with what it reads now
                    held = getattr(self, name)
                except Exception:
                    held = _UNSET
            if type(held) is not type(value) or held != value:  # re-assigning the value already held changes nothing
                global _value_epoch
                _value_epoch += 1
        super().__setattr__(name, value)

    def load_from_safetensors(self: T, tensors_path: str | Path, strict: bool = True) -> T:
        self.load_state_dict(load_from_safetensors(tensors_path), strict=strict)
        return self

    def to(self: T, device: torch.device | str | None = None, dtype: torch.dtype | None = None) -> T:  # type: ignore[override]
        return super().to(device=device, dtype=dtype)

    # -- printing ---------------------------------------------------------------------------
    def basic_attributes(self, init_attrs_only: bool = False) -> dict[str, Any]:
        """Public scalar (or sequence-of-scalar) attributes; optionally only the ctor
        arguments that differ from their default."""
        params = inspect.signature(self.__init__).parameters
        defaults = {k: p.default for k, p in params.items() if p.default is not inspect.Parameter.empty}
        out: dict[str, Any] = {}
        for key, value in self.__dict__.items():
            if key.startswith("_") or not _is_basic(value):
                continue
            if init_attrs_only and (key not in params or value == defaults.get(key, inspect.Parameter.empty)):
                continue
            out[key] = value
        return out

    def __str__(self) -> str:
        attrs = ", ".join(f"{k}={v}" for k, v in self.basic_attributes(init_attrs_only=True).items())
        return f"{type(self).__name__}({attrs})"

    def __repr__(self) -> str:
        return repr(ModuleTree(self))

    def pretty_print(self, depth: int = -1) -> None:
        print(ModuleTree(self).render(depth=depth))

    def _show_only_tag(self) -> bool:
        return False

    def get_path(self, parent: "Chain | None" = None, top: "Module | None" = None) -> str:
        """Dotted path of this module, as it appears in state-dict keys."""
        if parent is None or self is top:
            return type(self).__name__
        for key, child in parent._modules.items():
            if child is self:
                return parent.get_path(parent=parent.parent, top=top) + "." + key
        raise ValueError(f"{self} not found in {parent}")


class ContextModule(Module):
    """A module that knows its parent Chain and can therefore reach the context store."""

    _can_refresh_parent: bool = True

    def __init__(self, *args: Any, **kwargs: Any) -> None:
        super().__init__(*args, **kwargs)
        # kept in a list so that torch does not register the parent as a sub-module
        self._parent: list[Chain] = []

    @property
    def parent(self) -> "Chain | None":
        return self._parent[0] if self._parent else None

    @property
    def ensure_parent(self) -> "Chain":
        assert self._parent, "module does not have a parent"
        return self._parent[0]

    def get_parents(self) -> "list[Chain]":
        chain: list[Chain] = []
        node = self.parent
        while node is not None:
            chain.append(node)
            node = node.parent
        return chain

    def _set_parent(self, parent: "Chain | None") -> None:
        if not self._can_refresh_parent:
            return
        if parent is None:
            self._parent = []
            return
        assert any(m is self for m in parent), f"{self} not in {parent}"
        self._parent = [parent]

    @property
    def provider(self) -> ContextProvider:
        return self.ensure_parent.provider

    def use_context(self, context_name: str) -> Context:
        context = self.provider.get_context(context_name)
        assert context is not None, f"Context {context_name} not found."
        return context

    def structural_copy(self: T) -> T:
        """New instance of the same class sharing every non-torch public attribute
        (weights are *not* duplicated: leaves are shared by the caller)."""
        clone = object.__new__(type(self))
        for key, value in self.__dict__.items():
            if key.startswith("_"):
                continue
            owner = sys.modules.get(type(value).__module__)
            if owner is not None and "torch" not in owner.__name__:
                object.__setattr__(clone, key, value)
        ContextModule.__init__(clone)
        return clone

    def get_path(self, parent: "Chain | None" = None, top: "Module | None" = None) -> str:
        return super().get_path(parent=parent or self.parent, top=top)

    @contextlib.contextmanager
    def no_parent_refresh(self) -> Iterator[None]:
        previous = self._can_refresh_parent
        self._can_refresh_parent = False
        try:
            yield
        finally:
            self._can_refresh_parent = previous


class WeightedModule(Module):
    """A module with a ``weight`` tensor; device and dtype are read from it."""

    weight: Tensor

    @property
    def device(self) -> torch.device:
        return self.weight.device

    @property
    def dtype(self) -> torch.dtype:
        return self.weight.dtype

    def __str__(self) -> str:
        head = Module.__str__(self).removesuffix(")")
        # the separator is unconditional, as in the reference (module.py:253-258)
        return f"{head}, device={self.device}, dtype={str(self.dtype).removeprefix('torch.')})"


class _Node:
    __slots__ = ("value", "class_name", "children")

    def __init__(self, value: str, class_name: str, children: "list[_Node]") -> None:
        self.value, self.class_name, self.children = value, class_name, children

    def same_as(self, other: "_Node") -> bool:
        return (
            self.value == other.value
            and self.class_name == other.class_name
            and len(self.children) == len(other.children)
            and all(a.same_as(b) for a, b in zip(self.children, other.children))
        )

    # dict-style access kept for callers that treat nodes like the reference's TreeNode
    def __getitem__(self, key: str) -> Any:
        return getattr(self, key)

    def __setitem__(self, key: str, value: Any) -> None:
        setattr(self, key, value)


class ModuleTree:
    """Text rendering of a module tree (same output format as the reference)."""

    def __init__(self, module: Module) -> None:
        self.root = self._build(module)
        self._fold(self.root)

    def __iter__(self) -> Iterator[_Node]:
        return iter(self.root.children)

    def __str__(self) -> str:
        return f"{type(self).__name__}(root={self.root.value})"

    def __repr__(self) -> str:
        return self.render(depth=7)

    @staticmethod
    def _build(module: torch.nn.Module) -> _Node:
        if not isinstance(module, Module):
            return _Node(str(module), type(module).__name__, [])
        tag = module._tag
        if not tag:
            value = str(module)
        elif module._show_only_tag():
            value = f"({tag})"
        else:
            value = f"({tag}) {module}"
        return _Node(value, type(module).__name__, [ModuleTree._build(c) for c in module.children()])

    @staticmethod
    def _fold(node: _Node) -> None:
        kept: list[_Node] = []
        i, n = 0, len(node.children)
        while i < n:
            j = i + 1
            while j < n and node.children[i].same_as(node.children[j]):
                j += 1
            head = node.children[i]
            if j - i > 1:
                head.value += f" (x{j - i})"
            ModuleTree._fold(head)
            kept.append(head)
            i = j
        node.children = kept

    def render(self, depth: int = -1) -> str:
        lines: list[str] = []
        self._emit(self.root, self.root.value, "", True, True, depth, lines)
        return "\n".join(lines)

    # kept under the reference's private name because Chain error reporting uses it
    def _generate_tree_repr(self, node: _Node, /, *, depth: int = -1, is_root: bool = True, **_: Any) -> str:
        lines: list[str] = []
        self._emit(node, node.value, "", True, is_root, depth, lines)
        return "\n".join(lines)

    def _emit(self, node: _Node, label: str, prefix: str, last: bool, root: bool, depth: int, out: list[str]) -> None:
        branch = "" if root else ("└── " if last else "├── ")
        if depth == 0 and node.children:
            out.append(f"{prefix}{'└── ' if last else '├── '}{label} ...")
            return
        out.append(f"{prefix}{branch}{label}")
        totals: dict[str, int] = {}
        for child in node.children:
            totals[child.class_name] = totals.get(child.class_name, 0) + 1
        seen: dict[str, int] = {}
        child_prefix = prefix + ("    " if last else "│   ")
        for idx, child in enumerate(node.children):
            seen[child.class_name] = seen.get(child.class_name, 0) + 1
            child_label = child.value
            if totals[child.class_name] > 1:
                child_label = f"{child.value} #{seen[child.class_name]}"
            self._emit(child, child_label, child_prefix, idx == len(node.children) - 1, False, depth - 1 if depth > 0 else depth, out)

    @classmethod
    def shorten_tree_repr(cls, tree_repr: str, /, line_index: int = 0, max_lines: int = 20) -> str:
        lines = tree_repr.split("\n")
        lo = max(0, line_index - max_lines // 2)
        hi = min(len(lines), line_index + max_lines // 2 + 1)
        return "\n".join(lines[lo:hi])
