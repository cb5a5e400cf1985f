"""Peephole kernel fusion for Chains - a derived, cached, invalidatable plan.

The tree is never rewritten (adapters splice modules anywhere and tests assert on ``repr``):
instead every Chain caches a *plan* over its current children, rebuilt whenever the children
change (``Chain._regenerate_keys`` drops it).  A plan is a list of steps:

  ("call", name, layer)                    run the child as the reference would
  ("gn_silu", name, gn, silu)              GroupNorm -> SiLU            => one rb200_group_norm launch
  ("linear_geglu", name, linear, glu)      Linear -> GLU(GeLU)          => GEMM with GEGLU epilogue

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
