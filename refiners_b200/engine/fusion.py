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


def _linear_like(layer: Any) -> bool:
    """A plain Linear, or a LoraAdapter standing where a Linear stood (the adapter may be evaluated as one GEMM against
    the merged weight, see ``linear_params``)."""
    from refiners_b200.fluxion.adapters.lora import LoraAdapter
    from refiners_b200.fluxion.layers.leaves import Linear

    return type(layer) is Linear or type(layer) is LoraAdapter


def linear_params(layer: Any) -> tuple[Tensor, Tensor | None, int, int] | None:
    """(weight, bias, in_features, out_features) with which ``layer`` can be evaluated as ONE plain GEMM right now, or None.
    A hook-free Linear: its own parameters.  A hook-free LoraAdapter around a Linear whose LoRAs are all plain LinearLoras
    (and with LoRA merging enabled): the cached merged weight ``W + sum_i s_i B_i A_i`` (backend.merged_lora_weight)."""
    from refiners_b200.fluxion.adapters.lora import LoraAdapter
    from refiners_b200.fluxion.layers.leaves import Linear

    if type(layer) is Linear:
        return None if _hooked(layer) else (layer.weight, layer.bias, layer.in_features, layer.out_features)
    if type(layer) is LoraAdapter and B.lora_merge_enabled() and not _hooked(layer):
        triples = layer._fusable()  # None unless the children are [hook-free Linear, hook-free plain LinearLoras...]
        base = layer[0]
        if triples is not None and base.weight.is_cuda and all(d.is_cuda and u.is_cuda and d.dtype == base.weight.dtype for d, u, _ in triples):
            return B.merged_lora_weight(base.weight, triples), base.bias, base.in_features, base.out_features
    return None


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
            _linear_like(layer)
            and type(nxt) is GLU
            and type(nxt.activation) is GeLU
            and nxt.activation.approximation is GeLUApproximation.NONE
        ):
            plan.append(("linear_geglu", name, layer, nxt))
            i += 2
        elif _linear_like(layer) and (
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
            and not _hooked(step[3], step[3].activation)
            and (prm := linear_params(step[2])) is not None
            and B.geglu_fusable(prm[0])
        ):
            result = chain._call_fused(step[1], lambda x, prm=prm: B.linear_geglu(x, prm[0], prm[1]), *args)
        elif kind == "linear_act" and fuse and _cuda_tensor(args) and not _hooked(step[3]) and (prm := linear_params(step[2])) is not None:
            epi = step[4]
            result = chain._call_fused(step[1], lambda x, prm=prm, epi=epi: B.linear(x, prm[0], prm[1], epilogue=epi), *args)
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
        if _linear_like(last) and len(hargs) == 1 and isinstance(hargs[0], Tensor) and (prm := linear_params(last)) is not None:
            t = hargs[0]
            if t.is_cuda and t.shape[:-1] == residual.shape[:-1] and prm[3] == residual.shape[-1] and t.dtype == residual.dtype:
                fused = chain._call_fused(name, lambda u: B.linear(u, prm[0], prm[1], residual=residual), t)
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
    if isinstance(out, (int, float)) and out == 0:
        # ResidualAccumulator on a slot that still holds its initial 0.0 (unet.py:54-63 of the reference): x + 0.0 is x;
        # no kernel of this library writes in place, so handing the same tensor on is safe (10 full-size ATen adds per step)
        return residual
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
    if not any(_linear_like(m) for _, m in items):
        return NotImplemented
    # (weight, bias, in, out) of every child that is one plain GEMM right now (a Linear, or a LoraAdapter with merged weight)
    params = [linear_params(m) if _linear_like(m) and isinstance(a, Tensor) and a.is_cuda else None for (_, m), a in zip(items, args)]
    runs: list[tuple[int, int]] = []
    i = 0
    while i < n:
        j = i + 1
        first = params[i]
        if first is not None:
            while (
                j < n
                and params[j] is not None
                and _same_tensor(args[i], args[j])
                and params[j][2] == first[2]
                and (params[j][1] is None) == (first[1] is None)
                and params[j][0].dtype == first[0].dtype
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
        group = params[i:j]
        w, b = B.concat_linear_weights([g[0] for g in group], [g[1] for g in group])
        y = chain._call_fused(items[i][0], lambda x, w=w, b=b: B.linear(x, w, b), args[i])
        off = 0
        for k, g in zip(range(i, j), group):
            outs[k] = y[..., off : off + g[3]]
            off += g[3]
    return tuple(outs)
