"""Chain: the composable, editable module graph of the fluxion mirror.

Public behaviour follows /root/reference/src/refiners/fluxion/layers/chain.py
(`Chain` :53-642, `UseContext` :645, `SetContext` :684, `Lambda` :723, `Parallel` :756,
`Distribute` :793, `Passthrough` :834, `Sum` :867, `Residual` :901, `Concatenate` :930,
`Matmul` :967, `Return` :1001, `Breakpoint` :1019).  The walker stays pure Python, as the
north star asks; what is new is the *plan* slot: every structural edit goes through
``_regenerate_keys`` which drops the cached fusion plan of that chain
(see refiners_b200/engine/fusion.py), so kernel fusion is always derived from the current
tree and never rewrites it.
"""

from __future__ import annotations

import inspect
import re
import sys
import traceback
from typing import Any, Callable, Iterable, Iterator, Sequence, TypeVar, get_origin

import torch
from torch import Tensor

from refiners_b200.fluxion.context import ContextProvider, Contexts
from refiners_b200.fluxion.layers.base import ContextModule, Module, ModuleTree, WeightedModule
from refiners_b200.fluxion.utils import summarize_tensor

T = TypeVar("T", bound=Module)
TChain = TypeVar("TChain", bound="Chain")

# bumped on every structural edit anywhere; captured CUDA graphs compare against it
_structure_epoch = 0


def structure_epoch() -> int:
    return _structure_epoch


# observers of Chain.set_context (refiners_b200.engine.graph mirrors context tensors into the
# static buffers of a captured CUDA graph); a listener returning True has consumed the update
_context_listeners: list[Callable[["Chain", str, Any], bool]] = []


def generate_unique_names(modules: Sequence[Module]) -> dict[str, Module]:
    """``ClassName`` when the class appears once among the siblings, else ``ClassName_<i>``
    (1-based, in order).  These names are the state-dict keys."""
    names = [type(m).__name__ for m in modules]
    total: dict[str, int] = {}
    for n in names:
        total[n] = total.get(n, 0) + 1
    rank: dict[str, int] = {}
    keyed: dict[str, Module] = {}
    for n, m in zip(names, modules):
        rank[n] = rank.get(n, 0) + 1
        keyed[f"{n}_{rank[n]}" if total[n] > 1 else n] = m
    return keyed


def structural_copy(m: T) -> T:
    return m.structural_copy() if isinstance(m, ContextModule) else m


class ChainError(RuntimeError):
    def __init__(self, message: str, /) -> None:
        super().__init__(message)


_TRACE_SKIP = (
    (r"torch/nn/modules/", r"^_call_impl$"),
    (r"torch/nn/functional\.py", r""),
    (r"refiners_b200/fluxion/layers/", r"^_call_layer$"),
    (r"refiners_b200/fluxion/layers/", r"^forward$"),
    (r"refiners_b200/fluxion/layers/graph\.py", r""),
    (r"", r"^_"),
)


def _flatten(items: Any) -> list[Any]:
    if isinstance(items, tuple):
        flat: list[Any] = []
        for it in items:
            flat.extend(_flatten(it))
        return flat
    return [items]


class Chain(ContextModule):
    """Sequential composition with a context store and structural editing."""

    _modules: dict[str, Module]  # type: ignore[assignment]
    _tag = "CHAIN"

    def __init__(self, *args: Module | Iterable[Module]) -> None:
        super().__init__()
        self._provider = ContextProvider()
        self._plan: Any = None
        if len(args) == 1 and isinstance(args[0], Iterable) and not isinstance(args[0], Chain):
            children = tuple(args[0])
        else:
            children = tuple(args)  # type: ignore[arg-type]
        for child in children:
            if isinstance(child, ContextModule) and child._can_refresh_parent:
                assert child.parent is None or child.parent is self, (
                    f"{type(child).__name__} already has parent {type(child.parent).__name__}"
                )
        self._regenerate_keys(children)
        self._reset_context()
        for child in children:
            if isinstance(child, ContextModule) and child.parent is not self:
                child._set_parent(self)

    def __setattr__(self, name: str, value: Any) -> None:
        if isinstance(value, torch.nn.Module):
            raise ValueError(
                "Chain does not support setting modules by attribute. Instead, use a mutation method like `append` or"
                " wrap it within a single element list to prevent pytorch from registering it as a submodule."
            )
        super().__setattr__(name, value)

    # -- context plumbing ---------------------------------------------------------------------
    @property
    def provider(self) -> ContextProvider:
        return self._provider

    def init_context(self) -> Contexts:
        return {}

    def _register_provider(self, context: Contexts | None = None) -> None:
        if context:
            self._provider.update_contexts(context)
        shared = self._provider.contexts
        for child in self._modules.values():
            if isinstance(child, Chain):
                child._register_provider(shared)

    def _reset_context(self) -> None:
        self._register_provider(self.init_context())

    def set_context(self, context: str, value: Any) -> None:
        for listener in _context_listeners:
            if listener(self, context, value):
                return
        self._provider.set_context(context, value)
        self._register_provider()

    # -- execution ----------------------------------------------------------------------------
    def _call_layer(self, layer: Module, name: str, /, *args: Any) -> Any:
        try:
            return layer(*args)
        except Exception as exc:
            raise ChainError(self._describe_failure(exc, name, args)) from None

    def _call_fused(self, name: str, fn: Callable[..., Any], /, *args: Any) -> Any:
        """Run a fused kernel standing in for the child ``name`` (+ its successor) with the same
        error reporting as a plain child call."""
        try:
            return fn(*args)
        except Exception as exc:
            raise ChainError(self._describe_failure(exc, name, args)) from None

    def _steps(self) -> list[tuple[Any, ...]]:
        """Cached fusion plan over the current children (dropped by every structural edit)."""
        if self._plan is None:
            from refiners_b200.engine.fusion import build_plan

            self._plan = build_plan(self)
        return self._plan

    def _run_children(self, args: tuple[Any, ...], skip_last: bool = False) -> Any:
        """Walk the children (through the fusion plan); optionally stop before the last child and
        return the arguments it would receive."""
        from refiners_b200.engine.fusion import run_steps

        steps = self._steps()
        if skip_last:
            assert steps and steps[-1][0] == "call"
            result = run_steps(self, steps[:-1], args) if len(steps) > 1 else (args[0] if len(args) == 1 else args)
            return result
        return run_steps(self, steps, args)

    def forward(self, *args: Any) -> Any:
        result = self._run_children(args)
        self._reset_context()
        return result

    # -- error reporting ----------------------------------------------------------------------
    def _describe_failure(self, exc: Exception, name: str, args: tuple[Any, ...]) -> str:
        exc_type, _, tb = sys.exc_info()
        frames = [f for f in traceback.extract_tb(tb) if not self._skip_frame(f)]
        where = "".join(traceback.format_list(frames))
        what = re.sub(r"\n\s*\n", "\n", str(exc))
        shown_args = "\n".join(
            f"{i}: {summarize_tensor(a) if isinstance(a, Tensor) else a}" for i, a in enumerate(_flatten(args))
        )
        message = f"{where}\n{what}\n---------------\n{self._show_error_in_tree(name)}\n{shown_args}"
        if "Error" not in what and exc_type is not None:
            message = f"{exc_type.__name__}:\n {message}"
        return message

    @staticmethod
    def _skip_frame(frame: traceback.FrameSummary) -> bool:
        return any(re.search(fp, frame.filename) and re.search(np_, frame.name) for fp, np_ in _TRACE_SKIP)

    def _show_error_in_tree(self, name: str, /, max_lines: int = 20) -> str:
        tree = ModuleTree(self)
        wanted_class, _, wanted_rank = name.rpartition("_") if "_" in name else (name, "", "1")
        if not wanted_rank.isdigit():
            wanted_class, wanted_rank = name, "1"
        parents = self.get_parents()
        top: Module = parents[-1] if parents else self
        my_key = next((k for k, m in top.named_modules() if m is self), None)
        seen = 0
        for node in tree:
            if node.class_name == wanted_class:
                seen += 1
                if seen == int(wanted_rank):
                    full_key = None if my_key is None else ".".join((my_key, name))
                    node.value = f">>> {node.value} | {full_key}"
                    break
        text = tree._generate_tree_repr(tree.root, depth=3)
        lines = text.split("\n")
        hit = next((i for i, line in enumerate(lines) if line.startswith(">>>")), 0)
        return ModuleTree.shorten_tree_repr(text, line_index=hit, max_lines=max_lines)

    # -- container protocol -------------------------------------------------------------------
    def _regenerate_keys(self, modules: Iterable[Module]) -> None:
        global _structure_epoch
        self._modules = generate_unique_names(tuple(modules))  # type: ignore[assignment]
        self._plan = None
        _structure_epoch += 1

    def __getitem__(self, key: int | str | slice) -> Any:
        if isinstance(key, str):
            return self._modules[key]
        if isinstance(key, slice):
            clone = self.structural_copy()
            clone._regenerate_keys(list(clone)[key])
            return clone
        return list(self._modules.values())[key]

    def __iter__(self) -> Iterator[Module]:
        return iter(self._modules.values())

    def __len__(self) -> int:
        return len(self._modules)

    @property
    def device(self) -> torch.device | None:
        leaf = self.find(WeightedModule)
        return None if leaf is None else leaf.device

    @property
    def dtype(self) -> torch.dtype | None:
        leaf = self.find(WeightedModule)
        return None if leaf is None else leaf.dtype

    # -- search -------------------------------------------------------------------------------
    def walk(
        self,
        predicate: type[T] | Callable[[Module, "Chain"], bool] | None = None,
        recurse: bool = False,
    ) -> Iterator[tuple[Any, "Chain"]]:
        """Depth-first (module, parent) pairs matching ``predicate``.  Without ``recurse`` the
        search does not descend below a match.  A predicate may raise ``StopIteration`` to
        prune a whole sub-tree."""
        if get_origin(predicate) is not None:
            raise ValueError("subscripted generics cannot be used as predicates")
        if predicate is None:
            test: Callable[[Module, Chain], bool] = lambda _m, _p: True
        elif isinstance(predicate, type):
            wanted = predicate
            test = lambda m, _p: isinstance(m, wanted)
        else:
            test = predicate
        return self._walk(test, recurse)

    def _walk(self, test: Callable[[Module, "Chain"], bool], recurse: bool) -> Iterator[tuple[Module, "Chain"]]:
        for child in self._modules.values():
            try:
                matched = test(child, self)
            except StopIteration:
                continue
            if matched:
                yield child, self
                if not recurse:
                    continue
            if isinstance(child, Chain):
                yield from child._walk(test, recurse)

    def layers(self, layer_type: type[T], recurse: bool = False) -> Iterator[T]:
        for module, _ in self.walk(layer_type, recurse):
            yield module

    def find(self, layer_type: type[T]) -> T | None:
        return next(self.layers(layer_type), None)

    def ensure_find(self, layer_type: type[T]) -> T:
        found = self.find(layer_type)
        assert found is not None, f"could not find {layer_type} in {self}"
        return found

    def find_parent(self, module: Module) -> "Chain | None":
        if any(m is module for m in self._modules.values()):
            return self
        for _, parent in self.walk(lambda m, _p: m is module):
            return parent
        return None

    def ensure_find_parent(self, module: Module) -> "Chain":
        parent = self.find_parent(module)
        assert parent is not None, f"could not find {module} in {self}"
        return parent

    def layer(self, key: str | int | Sequence[str | int], layer_type: type[T] = Module) -> T:  # type: ignore[assignment]
        """Child (or descendant, for a path of keys) checked against ``layer_type``."""
        path: Sequence[str | int] = (key,) if isinstance(key, (str, int)) else key
        node: Module = self
        for i, step in enumerate(path):
            assert isinstance(node, Chain), f"layer {path[:i]} is {type(node)}, not a Chain"
            node = node[step]
        assert isinstance(node, layer_type), f"layer {key} is {type(node)}, not {layer_type}"
        return node

    # -- structural edits ---------------------------------------------------------------------
    def insert(self, index: int, module: Module) -> None:
        children = list(self)
        if index < 0:
            index = max(0, len(children) + index + 1)
        children.insert(index, module)
        self._regenerate_keys(children)
        if isinstance(module, ContextModule):
            module._set_parent(self)
        self._register_provider()

    def _index_of_type(self, module_type: type[Module]) -> int:
        for i, child in enumerate(self):
            if isinstance(child, module_type):
                return i
        raise ValueError(f"No module of type {module_type.__name__} found in the chain.")

    def insert_before_type(self, module_type: type[Module], new_module: Module) -> None:
        self.insert(self._index_of_type(module_type), new_module)

    def insert_after_type(self, module_type: type[Module], new_module: Module) -> None:
        self.insert(self._index_of_type(module_type) + 1, new_module)

    def append(self, module: Module) -> None:
        self.insert(-1, module)

    def pop(self, index: int = -1) -> Module:
        children = list(self)
        if index < 0:
            index += len(children)
        if not 0 <= index < len(children):
            raise IndexError("Index out of range.")
        gone = children.pop(index)
        if isinstance(gone, ContextModule):
            gone._set_parent(None)
        self._regenerate_keys(children)
        return gone

    def _position(self, module: Module) -> int:
        for i, child in enumerate(self):
            if child is module:
                return i
        raise ValueError(f"{module} is not in {self}")

    def remove(self, module: Module) -> None:
        children = list(self)
        del children[self._position(module)]
        self._regenerate_keys(children)
        if isinstance(module, ContextModule):
            module._set_parent(None)

    def replace(self, old_module: Module, new_module: Module, old_module_parent: "Chain | None" = None) -> None:
        children = list(self)
        children[self._position(old_module)] = new_module
        self._regenerate_keys(children)
        if isinstance(new_module, ContextModule):
            new_module._set_parent(self)
        if isinstance(old_module, ContextModule):
            old_module._set_parent(old_module_parent)
        self._register_provider()

    def structural_copy(self: TChain) -> TChain:
        """Copy the Chain skeleton; weighted leaves are shared with the original."""
        pre = getattr(self, "_pre_structural_copy", None)
        if callable(pre):
            pre()
        children = [structural_copy(m) for m in self]
        clone = super().structural_copy()
        object.__setattr__(clone, "_provider", ContextProvider.create(clone.init_context()))
        object.__setattr__(clone, "_plan", None)
        for child in children:
            clone.append(child)
        post = getattr(clone, "_post_structural_copy", None)
        if callable(post):
            post(self)
        return clone

    def _show_only_tag(self) -> bool:
        return type(self) is Chain


class UseContext(ContextModule):
    """Emit ``func(contexts[context][key])``, ignoring its positional inputs."""

    def __init__(self, context: str, key: str) -> None:
        super().__init__()
        self.context = context
        self.key = key
        self.func: Callable[[Any], Any] = lambda x: x

    def __call__(self, *args: Any) -> Any:
        store = self.use_context(self.context)
        assert store, f"context {self.context} is unset"
        value = store.get(self.key)
        assert value is not None, f"context entry {self.context}.{self.key} is unset"
        return self.func(value)

    def compose(self, func: Callable[[Any], Any]) -> "UseContext":
        self.func = func
        return self

    def __repr__(self) -> str:
        return f"{type(self).__name__}(context={self.context!r}, key={self.key!r})"


class SetContext(ContextModule):
    """Store (or feed to ``callback``) its input under ``contexts[context][key]``; pass it on."""

    def __init__(self, context: str, key: str, callback: Callable[[Any, Any], Any] | None = None) -> None:
        super().__init__()
        self.context = context
        self.key = key
        self.callback = callback

    def __call__(self, x: Tensor) -> Tensor:
        store = self.use_context(self.context)
        if store:
            if self.callback is None:
                store[self.key] = x
            else:
                self.callback(store[self.key], x)
        return x

    def __repr__(self) -> str:
        return f"{type(self).__name__}(context={self.context!r}, key={self.key!r})"


class Lambda(Module):
    def __init__(self, func: Callable[..., Any]) -> None:
        super().__init__()
        self.func = func

    def forward(self, *args: Any) -> Any:
        return self.func(*args)

    def __str__(self) -> str:
        name = getattr(self.func, "__name__", "partial_function")
        return f"Lambda({name}{inspect.signature(self.func)})"


class Parallel(Chain):
    """Every child sees the same inputs; outputs are gathered in a tuple."""

    _tag = "PAR"

    def forward(self, *args: Any) -> tuple[Any, ...]:
        return tuple(self._call_layer(layer, name, *args) for name, layer in self._modules.items())

    def _show_only_tag(self) -> bool:
        return type(self) is Parallel


class Distribute(Chain):
    """The i-th child receives the i-th input."""

    _tag = "DISTR"

    def forward(self, *args: Any) -> tuple[Any, ...]:
        n, m = len(args), len(self._modules)
        assert n == m, f"Number of positional arguments ({n}) must match number of sub-modules ({m})."
        if n > 1 and isinstance(args[0], Tensor) and args[0].is_cuda:
            from refiners_b200.engine import fusion

            fused = fusion.try_fuse_distribute(self, args)
            if fused is not NotImplemented:
                return fused
        return tuple(self._call_layer(layer, name, arg) for arg, (name, layer) in zip(args, self._modules.items()))

    def _show_only_tag(self) -> bool:
        return type(self) is Distribute


class Passthrough(Chain):
    """Run the children for their side effects, hand the inputs through unchanged."""

    _tag = "PASS"

    def forward(self, *inputs: Any) -> Any:
        Chain.forward(self, *inputs)
        return inputs

    def _show_only_tag(self) -> bool:
        return type(self) is Passthrough


class Sum(Chain):
    """Add up the outputs of all children (each fed the same inputs)."""

    _tag = "SUM"

    def forward(self, *inputs: Any) -> Any:
        on_gpu = bool(inputs) and isinstance(inputs[0], Tensor) and inputs[0].is_cuda
        if on_gpu:
            from refiners_b200.engine import fusion

            fused = fusion.try_fuse_sum(self, inputs)
            if fused is not NotImplemented:
                return fused
        total: Any = None
        for layer in self._modules.values():
            term = layer(*inputs)
            if isinstance(term, tuple):
                term = sum(term)
            if total is None:
                total = term
            elif on_gpu and isinstance(term, Tensor) and isinstance(total, Tensor) and term.shape == total.shape and term.dtype == total.dtype:
                from refiners_b200 import backend as B

                total = B.add(total, term)
            else:
                total = total + term
        return total

    def _show_only_tag(self) -> bool:
        return type(self) is Sum


class Residual(Chain):
    _tag = "RES"

    def forward(self, *inputs: Any) -> Any:
        assert len(inputs) == 1, "Residual connection can only be used with a single input."
        x = inputs[0]
        if isinstance(x, Tensor) and x.is_cuda:
            from refiners_b200.engine import fusion

            # the skip connection is added in the epilogue of the chain's final GEMM when the tail is
            # a Linear (possibly nested in plain Chains), else by one add kernel
            return fusion.forward_with_residual(self, inputs, x)
        return Chain.forward(self, *inputs) + x


class Concatenate(Chain):
    _tag = "CAT"

    def __init__(self, *modules: Module, dim: int = 0) -> None:
        super().__init__(*modules)
        self.dim = dim

    def forward(self, *args: Any) -> Tensor:
        parts = [p for p in (layer(*args) for layer in self._modules.values()) if p is not None]
        if self.dim == 1 and len(parts) > 1 and isinstance(parts[0], Tensor) and parts[0].is_cuda:
            from refiners_b200 import backend as B

            if B.concat_channels_supported(parts):
                return B.concat_channels(parts)  # skip connections: one vectorised channels-last pass
        return torch.cat(parts, dim=self.dim)

    def _show_only_tag(self) -> bool:
        return type(self) is Concatenate


class Matmul(Chain):
    _tag = "MATMUL"

    def __init__(self, input: Module, other: Module) -> None:
        super().__init__(input, other)

    def forward(self, *args: Tensor) -> Tensor:
        return torch.matmul(self[0](*args), self[1](*args))


class ReturnException(Exception):
    def __init__(self, value: Tensor) -> None:
        self.value = value


class Return(Module):
    def forward(self, x: Tensor) -> Any:
        raise ReturnException(x)


class Breakpoint(ContextModule):
    def __init__(self, vscode: bool = True) -> None:
        super().__init__()
        self.vscode = vscode

    def forward(self, *args: Any) -> Any:
        if self.vscode:
            import debugpy  # type: ignore

            debugpy.breakpoint()
        else:
            breakpoint()
        return args[0] if len(args) == 1 else args
