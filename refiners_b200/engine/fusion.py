"""Peephole kernel fusion for Chains - a derived, cached, invalidatable plan.

The tree is never rewritten (adapters splice modules anywhere and tests assert on ``repr``):
instead every Chain caches a *plan* over its current children, rebuilt whenever the children
change (``Chain._regenerate_keys`` drops it).  A plan is a list of steps:

  ("call", name, layer)                    run the child as the reference would
  ("gn_silu", name, gn, silu)              GroupNorm -> SiLU            => one rb200_group_norm launch
  ("linear_geglu", name, linear, glu)      Linear -> GLU(GeLU)          => GEMM with GEGLU epilogue
  ("linear_act", name, linear, act, epi)   Linear -> GeLU (erf) | SiLU  => GEMM with activation epilogue

and two tail fusions used by the containers that add a skip connection:

  Residual(..., Linear)                    => the last GEMM adds the residual in its epilogue
  ResidualBlock = Sum(Chain(..., Conv2d), shortcut) => the last conv adds shortcut(x) in its epilogue

A step falls back to plain calls when a participating leaf has forward hooks (so hook-based
tools such as the reference's ModelConverter still see every leaf's true output), when the
inputs are not CUDA tensors, or when fusion is switched off (``backend.set_fusion(False)`` /
``RB200_FUSION=0``).
"""

from __future__ import annotations

from typing import Any

from torch import Tensor

from refiners_b200 import backend as B


def _hooked(*modules: Any) -> bool:
    return any(m._forward_hooks or m._forward_pre_hooks for m in modules)


def build_plan(chain: Any) -> list[tuple[Any, ...]]:
    from refiners_b200.fluxion.layers.leaves import GLU, GeLU, GeLUApproximation, GroupNorm, Linear, SiLU

    items = list(chain._modules.items())
    plan: list[tuple[Any, ...]] = []
    i = 0
    while i < len(items):
        name, layer = items[i]
        nxt = items[i + 1][1] if i + 1 < len(items) else None
        if type(layer) is GroupNorm and type(nxt) is SiLU:
            plan.append(("gn_silu", name, layer, nxt))
            i += 2
        elif (
            type(layer) is Linear
            and type(nxt) is GLU
            and type(nxt.activation) is GeLU
            and nxt.activation.approximation is GeLUApproximation.NONE
        ):
            plan.append(("linear_geglu", name, layer, nxt))
            i += 2
        elif type(layer) is Linear and (
            type(nxt) is SiLU or (type(nxt) is GeLU and nxt.approximation is GeLUApproximation.NONE)
        ):
            plan.append(("linear_act", name, layer, nxt, B.EPI_SILU if type(nxt) is SiLU else B.EPI_GELU))
            i += 2
        else:
            plan.append(("call", name, layer))
            i += 1
    return plan


def _cuda_tensor(args: tuple[Any, ...]) -> bool:
    return len(args) == 1 and isinstance(args[0], Tensor) and args[0].is_cuda


def run_steps(chain: Any, steps: list[tuple[Any, ...]], args: tuple[Any, ...]) -> Any:
    """Execute plan steps in order, threading tuple results like Chain.forward does."""
    result: Any = None
    fuse = B.fusion_enabled()
    for step in steps:
        kind = step[0]
        if kind == "gn_silu" and fuse and _cuda_tensor(args) and not _hooked(step[2], step[3]):
            gn = step[2]
            result = chain._call_fused(step[1], lambda x, gn=gn: B.group_norm(x, gn.num_groups, gn.weight, gn.bias, gn.eps, silu=True), *args)
        elif (
            kind == "linear_geglu"
            and fuse
            and _cuda_tensor(args)
            and not _hooked(step[2], step[3], step[3].activation)
            and B.geglu_fusable(step[2].weight)
        ):
            lin = step[2]
            result = chain._call_fused(step[1], lambda x, lin=lin: B.linear_geglu(x, lin.weight, lin.bias), *args)
        elif kind == "linear_act" and fuse and _cuda_tensor(args) and not _hooked(step[2], step[3]):
            lin, epi = step[2], step[4]
            result = chain._call_fused(step[1], lambda x, lin=lin, epi=epi: B.linear(x, lin.weight, lin.bias, epilogue=epi), *args)
        elif kind == "call":
            result = chain._call_layer(step[2], step[1], *args)
        else:  # unfused pair
            result = chain._call_layer(step[2], step[1], *args)
            args = result if isinstance(result, tuple) else (result,)
            pair_name = next(k for k, m in chain._modules.items() if m is step[3])
            result = chain._call_layer(step[3], pair_name, *args)
        args = result if isinstance(result, tuple) else (result,)
    return result


def tail_linear(chain: Any) -> Any:
    """The final child if it is a plain, hook-free Linear whose epilogue may take a residual."""
    from refiners_b200.fluxion.layers.leaves import Linear

    if not chain._modules:
        return None
    last = next(reversed(chain._modules.values()))
    if type(last) is Linear and not _hooked(last):
        return last
    return None


def tail_conv(chain: Any) -> Any:
    from refiners_b200.fluxion.layers.leaves import Conv2d

    if not chain._modules:
        return None
    last = next(reversed(chain._modules.values()))
    if type(last) is Conv2d and not _hooked(last) and B.conv_supported(last):
        return last
    return None


# -- Sum fusers ---------------------------------------------------------------------------------
# callables (sum_chain, inputs) -> result | NotImplemented, tried in order by Sum.forward on CUDA
_sum_fusers: list[Any] = []


def register_sum_fuser(fn: Any) -> None:
    if fn not in _sum_fusers:
        _sum_fusers.append(fn)


def try_fuse_sum(chain: Any, inputs: tuple[Any, ...]) -> Any:
    if not B.fusion_enabled():
        return NotImplemented
    for fn in _sum_fusers:
        out = fn(chain, inputs)
        if out is not NotImplemented:
            return out
    return NotImplemented


# -- tail residual ------------------------------------------------------------------------------
def forward_with_residual(chain: Any, args: tuple[Any, ...], residual: Tensor) -> Any:
    """``chain(*args) + residual`` with the addition folded into the epilogue of the chain's final
    GEMM when the tail is (nested plain Chains ending in) a Linear or a fusable LoraAdapter.
    Always returns the sum (falls back to a separate add kernel)."""
    from refiners_b200.fluxion.layers.graph import Chain
    from refiners_b200.fluxion.layers.leaves import Linear

    steps = chain._steps()
    fused: Any = NotImplemented
    if steps and steps[-1][0] == "call" and B.fusion_enabled():
        name, last = steps[-1][1], steps[-1][2]
        h = run_steps(chain, steps[:-1], args) if len(steps) > 1 else (args[0] if len(args) == 1 else args)
        hargs = h if isinstance(h, tuple) else (h,)
        if type(last) is Linear and not _hooked(last) and len(hargs) == 1 and isinstance(hargs[0], Tensor):
            t = hargs[0]
            if t.is_cuda and t.shape[:-1] == residual.shape[:-1] and last.out_features == residual.shape[-1] and t.dtype == residual.dtype:
                fused = chain._call_fused(name, lambda u: B.linear(u, last.weight, last.bias, residual=residual), t)
        elif isinstance(last, Chain) and type(last).forward is Chain.forward and not _hooked(last):
            fused = chain._call_fused(name, lambda *u: forward_with_residual(last, u, residual), *hargs)
        elif hasattr(last, "_forward_with_residual") and not _hooked(last):
            fused = chain._call_fused(name, lambda *u: last._forward_with_residual(u, residual), *hargs)
        if fused is NotImplemented:
            out = chain._call_layer(last, name, *hargs)
        else:
            chain._reset_context()
            return fused
    else:
        out = run_steps(chain, steps, args)
    chain._reset_context()
    if isinstance(out, Tensor) and out.is_cuda and out.shape == residual.shape and out.dtype == residual.dtype:
        return B.add(out, residual)
    return out + residual


# -- Distribute: sibling Linears fed the same tensor become one GEMM -----------------------------------
def _same_tensor(a: Any, b: Any) -> bool:
    return (
        isinstance(a, Tensor)
        and isinstance(b, Tensor)
        and (a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape and a.stride() == b.stride() and a.dtype == b.dtype))
    )


def try_fuse_distribute(chain: Any, args: tuple[Any, ...]) -> Any:
    """Distribute(Linear, Linear, ...)(x, x, ...) -> one GEMM over the row-concatenated weights for
    every run of children that are plain Linears receiving the same tensor (self-attention q/k/v,
    cross-attention k/v).  Outputs are column views of the fused result."""
    from refiners_b200.fluxion.layers.leaves import Linear

    if not B.fusion_enabled():
        return NotImplemented
    items = list(chain._modules.items())
    if len(items) != len(args) or len(items) < 2:
        return NotImplemented
    n = len(items)
    runs: list[tuple[int, int]] = []
    i = 0
    while i < n:
        j = i + 1
        lin = items[i][1]
        if type(lin) is Linear and isinstance(args[i], Tensor) and args[i].is_cuda and not _hooked(lin):
            while (
                j < n
                and type(items[j][1]) is Linear
                and not _hooked(items[j][1])
                and _same_tensor(args[i], args[j])
                and items[j][1].in_features == lin.in_features
                and (items[j][1].bias is None) == (lin.bias is None)
                and items[j][1].weight.dtype == lin.weight.dtype
            ):
                j += 1
        runs.append((i, j))
        i = j
    if all(j - i == 1 for i, j in runs):
        return NotImplemented
    outs: list[Any] = [None] * n
    for i, j in runs:
        if j - i == 1:
            outs[i] = chain._call_layer(items[i][1], items[i][0], args[i])
            continue
        group = [items[k][1] for k in range(i, j)]
        w, b = B.concat_linear_weights([m.weight for m in group], [m.bias for m in group])
        y = chain._call_fused(items[i][0], lambda x, w=w, b=b: B.linear(x, w, b), args[i])
        off = 0
        for k, m in zip(range(i, j), group):
            outs[k] = y[..., off : off + m.out_features]
            off += m.out_features
    return tuple(outs)
