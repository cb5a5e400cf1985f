"""Python side of the drop-in boundary: ctypes bindings of ``librefiners_b200.so`` (the C ABI
declared in include/refiners_b200.h) surfaced as ``torch.ops.refiners_b200.*`` custom ops.

Rules enforced here:
  * CUDA tensors only ever run on our kernels.  If the shared library is missing the first
    call raises - there is no ATen/CPU fallback behind these ops.
  * Inference only: ops refuse to run when autograd would need a graph.
  * Nothing here synchronises or allocates outside torch's caching allocator, so every op is
    CUDA-graph capturable.
  * Packed weights (conv KRSC layout, GEGLU interleave, concatenated LoRA factors) are cached
    per parameter storage + version, as the reference swaps ``.weight`` objects freely
    (fluxion/adapters/lora.py:168-178, image_prompt.py:344-347 in the reference).
"""

from __future__ import annotations

import ctypes
import os
from pathlib import Path
from typing import Any, Sequence

import torch
from torch import Tensor

_LIB_PATH = Path(__file__).resolve().parent.parent / "csrc" / "librefiners_b200.so"
_lib: ctypes.CDLL | None = None
_DT = {torch.bfloat16: 0, torch.float16: 1, torch.float32: 2}
_UNARY = {"silu": 0, "gelu": 1, "gelu_tanh": 2, "gelu_sigmoid": 3, "relu": 4, "sigmoid": 5}
EPI_NONE, EPI_GEGLU, EPI_GELU, EPI_SILU = 0, 1, 2, 3

_fusion = os.environ.get("RB200_FUSION", "1") != "0"
# LoRA-adapted Linears: 1 (default) = fold the rank terms into a cached effective weight W + sum_i s_i B_i A_i (one GEMM
# launch per adapted layer, re-merged when a factor or a scale changes); 0 = keep them separate (rank-space GEMM + base
# GEMM with the up-projection as extra k-blocks, two launches)
_lora_merge = os.environ.get("RB200_LORA_MERGE", "1") != "0"


class BackendError(RuntimeError):
    pass


def fusion_enabled() -> bool:
    return _fusion


def set_fusion(enabled: bool) -> bool:
    global _fusion
    previous, _fusion = _fusion, bool(enabled)
    return previous


def lora_merge_enabled() -> bool:
    return _lora_merge


def set_lora_merge(enabled: bool) -> bool:
    global _lora_merge
    previous, _lora_merge = _lora_merge, bool(enabled)
    return previous


def library_path() -> Path:
    return _LIB_PATH


def is_built() -> bool:
    return _LIB_PATH.exists()


class _LoraDesc(ctypes.Structure):
    _fields_ = [("down", ctypes.c_void_p), ("up", ctypes.c_void_p), ("scale", ctypes.c_float), ("rank", ctypes.c_int32)]


_P, _I, _L, _F, _Z = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_size_t
_SIGNATURES: dict[str, tuple[Any, list[Any]]] = {
    "rb200_abi_version": (_I, []),
    "rb200_last_error": (ctypes.c_char_p, []),
    "rb200_device_info": (_I, [ctypes.POINTER(_I)] * 3),
    "rb200_launch_count": (_L, []),
    "rb200_set_kernel_mode": (_I, [_I]),
    "rb200_linear_workspace_bytes": (_Z, [_L, _I, _I]),
    "rb200_linear": (_I, [_P, _I, _P, _L, _P, _L, _P, _P, _L, _L, _L, _L, _I, _P, _P, _P, _P, _L, _I, _P, _Z]),
    "rb200_lora_pack": (_I, [_P, _I, _I, ctypes.POINTER(_LoraDesc), _L, _L, _P, _P, _P, _I]),
    "rb200_geglu_pack": (_I, [_P, _I, _P, _P, _P, _P, _L, _L]),
    "rb200_conv2d_pack_weight": (_I, [_P, _I, _P, _P, _L, _L, _I, _I]),
    "rb200_conv2d": (_I, [_P, _I, _P, _P, _P, _P, _P, _P, _L, _L, _L, _L, _L, _I, _I, _I, _I, _I]),
    "rb200_group_norm_workspace_bytes": (_Z, [_L, _L, _I]),
    "rb200_group_norm": (_I, [_P, _I, _P, _P, _L, _L, _L, _I, _F, _P, _P, _I, _P, _Z]),
    "rb200_group_norm_fixed": (_I, [_P, _I, _P, _P, _L, _L, _L, _I, _F, _P, _P, _I, _P, _Z, _P, _I]),
    "rb200_layer_norm": (_I, [_P, _I, _P, _P, _L, _L, _F, _P, _P]),
    "rb200_unary": (_I, [_P, _I, _P, _P, _L, _I]),
    "rb200_geglu": (_I, [_P, _I, _P, _P, _L, _L]),
    "rb200_add": (_I, [_P, _I, _P, _P, _P, _L, _F]),
    "rb200_cfg_scale_input": (_I, [_P, _I, _P, _P, _L, _P, _I]),
    "rb200_cfg_euler": (_I, [_P, _I, _P, _P, _P, _L, _P, _F, _I]),
    "rb200_attention_probs": (_I, [_P, _I, _P, _P, _P, _L, _I, _L, _L, _I, _L, _L, _L, _L, _F]),
    "rb200_sdpa": (_I, [_P, _I, _P, _P, _P, _P, _L, _I, _L, _L, _I] + [_L] * 8 + [_F, _I, _P, _P, _L] + [_L] * 4 + [_F]),
    "rb200_sam_attention_workspace_bytes": (_Z, [_L, _I, _I, _I, _I]),
    "rb200_sam_attention": (_I, [_P, _I, _P, _P, _P, _P, _L, _I, _I, _I, _I, _P, _Z]),
    "rb200_patchify": (_I, [_P, _I, _P, _P, _L, _L, _L, _L, _I, _L, _L, _L, _L]),
    "rb200_pad_channels": (_I, [_P, _I, _P, _P, _L, _I, _I, _I, _I, _L, _L, _L, _L]),
    "rb200_concat_channels": (_I, [_P, _I, _I, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int), _P, _L]),
    "rb200_resize_nearest": (_I, [_P, _I, _P, _P, _L, _I, _I, _I, _I, _I]),
    "rb200_avg_pool2d": (_I, [_P, _I, _P, _P, _L, _I, _I, _I, _I]),
    "rb200_style_aligned_workspace_bytes": (_Z, [_L, _L]),
    "rb200_style_aligned": (_I, [_P, _I, _P, _P, _L, _L, _L, _L, _L, _I, _I, _F, _F, _P, _Z]),
    "rb200_window_partition": (_I, [_P, _I, _P, _P, _L, _I, _I, _I, _I, _I]),
}


def exported_symbols() -> list[str]:
    return list(_SIGNATURES)


def load_library() -> ctypes.CDLL:
    """Load (once) and type the C ABI.  Raises BackendError when the .so is absent or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise BackendError(
            f"{_LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  refiners_b200 has no fallback for CUDA tensors."
        )
    lib = ctypes.CDLL(str(_LIB_PATH))
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means the header and the .so disagree
        fn.restype, fn.argtypes = restype, argtypes
    if lib.rb200_abi_version() != 1:
        raise BackendError(f"ABI mismatch: library reports {lib.rb200_abi_version()}, bindings expect 1")
    _lib = lib
    return lib


def _check(rc: int) -> None:
    if rc != 0:
        message = load_library().rb200_last_error()
        raise BackendError(f"rb200 error {rc}: {message.decode() if message else '?'}")


def launch_count() -> int:
    return int(load_library().rb200_launch_count())


def set_kernel_mode(mode: int) -> int:
    return int(load_library().rb200_set_kernel_mode(int(mode)))


def device_info() -> tuple[int, int, int]:
    a, b, c = _I(), _I(), _I()
    _check(load_library().rb200_device_info(ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
    return a.value, b.value, c.value


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _dtype_code(t: Tensor) -> int:
    try:
        return _DT[t.dtype]
    except KeyError:
        raise BackendError(f"unsupported dtype {t.dtype} (bf16, fp16, fp32 only)") from None


def _inference_only(*tensors: Tensor | None) -> None:
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise BackendError(
            "refiners_b200 kernels are inference-only: wrap the call in refiners_b200.fluxion.no_grad() "
            "(training through the custom ops is out of scope)"
        )


def _same(x: Tensor, *others: Tensor | None) -> None:
    for o in others:
        if o is not None and (o.dtype != x.dtype or o.device != x.device):
            raise BackendError(f"operand mismatch: {o.dtype}@{o.device} vs {x.dtype}@{x.device}")


# ------------------------------------------------------------------------------------ caches
class _PackCache:
    """Derived tensors (packed weights) cached per SOURCE TENSOR OBJECT.

    Keyed on ``id()`` of the source tensors and validated with weak references plus
    (data_ptr, _version): a storage address alone is not an identity - the caching allocator
    hands a freed weight's address to the next model's weight of the same shape - and the
    reference swaps ``.weight`` objects (lora.py:168-178, image_prompt.py:344-347), moves them
    with ``.to()`` and edits them in place.  Entries die with their source tensor."""

    def __init__(self) -> None:
        self._items: dict[tuple[Any, ...], tuple[tuple[Any, ...], Any]] = {}

    @staticmethod
    def key(*tensors: Tensor | None) -> tuple[Any, ...]:
        return tuple(None if t is None else id(t) for t in tensors)

    @staticmethod
    def _stamp(tensors: Sequence[Tensor | None]) -> tuple[Any, ...]:
        return tuple(None if t is None else (t.data_ptr(), t._version, tuple(t.shape), t.dtype) for t in tensors)

    def get(self, key: tuple[Any, ...], tensors: Sequence[Tensor | None]) -> Any:
        hit = self._items.get(key)
        if hit is None:
            return None
        refs, stamp, value = hit
        if stamp != self._stamp(tensors) or any((r is None) != (t is None) or (r is not None and r() is not t) for r, t in zip(refs, tensors)):
            del self._items[key]
            return None
        return value

    def put(self, key: tuple[Any, ...], tensors: Sequence[Tensor | None], value: Any) -> Any:
        import weakref

        def drop(_ref: Any, key: tuple[Any, ...] = key, items: dict = self._items) -> None:
            items.pop(key, None)

        refs = tuple(None if t is None else weakref.ref(t, drop) for t in tensors)
        self._items[key] = (refs, self._stamp(tensors), value)
        return value

    def clear(self) -> None:
        self._items.clear()

    def __len__(self) -> int:
        return len(self._items)


_conv_cache = _PackCache()
_conv_pad_cache = _PackCache()
_patch_cache = _PackCache()
_geglu_cache = _PackCache()
_lora_cache = _PackCache()
_concat_cache = _PackCache()
_merge_cache = _PackCache()


def clear_caches() -> None:
    for cache in (_conv_cache, _conv_pad_cache, _patch_cache, _geglu_cache, _lora_cache, _concat_cache, _merge_cache):
        cache.clear()


# --------------------------------------------------------------- step-invariant hoisting (engine.graph)
class InvariantMemo:
    """Results of ops whose inputs do not change from one denoising step to the next.

    A captured CUDA graph (refiners_b200.engine.graph) marks the static buffers of step-invariant contexts - the text
    embedding, the IP-Adapter image embedding, a ControlNet / ControlLora condition image - with ``mark``.  While a memo
    is active, ``linear`` / ``conv2d`` / ``unary`` on marked tensors run ONCE: the result is kept (and marked in turn,
    so a chain of invariant ops such as ConditionEncoder's conv - SiLU stack is hoisted as a whole), later calls with
    the same operands return it without launching anything - in particular the call made under stream capture, so the
    graph replayed every step does not contain these kernels.  ``refresh`` recomputes every entry in place when the
    owner sees an invariant context change (a new prompt).  SURVEY.md section 8f rank 2:
    cross_attention.py:52-65 (text K / V projections), image_prompt.py:243-262, control_lora.py:193-201 of the reference."""

    def __init__(self) -> None:
        self.entries: dict[tuple[Any, ...], tuple[Tensor, Any]] = {}
        self.hits = 0

    @staticmethod
    def mark(t: Tensor) -> Tensor:
        t._rb200_invariant = True  # type: ignore[attr-defined]
        return t

    @staticmethod
    def marked(*tensors: Tensor | None) -> bool:
        return all(t is None or getattr(t, "_rb200_invariant", False) for t in tensors)

    def fetch(self, key: tuple[Any, ...], thunk: Any) -> Tensor:
        hit = self.entries.get(key)
        if hit is not None:
            self.hits += 1
            return hit[0]
        out = self.mark(thunk())
        self.entries[key] = (out, thunk)
        return out

    def refresh(self) -> None:
        for out, thunk in self.entries.values():  # insertion order = dependency order
            out.copy_(thunk())

    def __len__(self) -> int:
        return len(self.entries)


_memo: InvariantMemo | None = None


def set_invariant_memo(memo: InvariantMemo | None) -> InvariantMemo | None:
    """Activate (or, with None, deactivate) a memo; returns the previous one."""
    global _memo
    previous, _memo = _memo, memo
    return previous


# ------------------------------------------------------------------------- raw op kernels
def _linear_impl(
    x: Tensor,
    w: Tensor,
    bias: Tensor | None,
    residual: Tensor | None,
    lora_down: Tensor | None,
    lora_up: Tensor | None,
    lora_scale: Tensor | None,
    epilogue: int,
) -> Tensor:
    lib = load_library()
    _same(x, w, bias, residual, lora_down, lora_up)
    K = x.shape[-1]
    N = w.shape[0]
    if w.shape[1] != K:
        raise BackendError(f"linear: x[..., {K}] incompatible with weight {tuple(w.shape)}")
    x2 = x.reshape(-1, K)
    if x2.stride(-1) != 1:
        x2 = x2.contiguous()
    if w.stride(-1) != 1:
        w = w.contiguous()
    M = x2.shape[0]
    n_out = N // 2 if epilogue == EPI_GEGLU else N
    y = torch.empty((M, n_out), device=x.device, dtype=x.dtype)
    res2 = None
    if residual is not None:
        res2 = residual.reshape(-1, n_out)
        if res2.stride(-1) != 1:
            res2 = res2.contiguous()
        if res2.shape[0] != M:
            raise BackendError("linear: residual shape mismatch")
    r_pad = 0 if lora_down is None else lora_down.shape[0]
    ws = None
    ws_bytes = 0
    if r_pad:
        ws_bytes = lib.rb200_linear_workspace_bytes(M, r_pad, _dtype_code(x))
        ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    if M > 0:
        _check(
            lib.rb200_linear(
                _stream(), _dtype_code(x), x2.data_ptr(), x2.stride(0), w.data_ptr(), w.stride(0), _ptr(bias),
                y.data_ptr(), y.stride(0), M, N, K, r_pad, _ptr(lora_down), _ptr(lora_up), _ptr(lora_scale),
                _ptr(res2), 0 if res2 is None else res2.stride(0), epilogue, _ptr(ws), ws_bytes,
            )
        )
    return y.reshape(*x.shape[:-1], n_out)


def _conv2d_impl(
    x: Tensor,
    w_packed: Tensor,
    bias: Tensor | None,
    chan_bias: Tensor | None,
    residual: Tensor | None,
    R: int,
    S: int,
    stride: int,
    pad: int,
    epilogue: int,
) -> Tensor:
    lib = load_library()
    _same(x, w_packed, bias, chan_bias, residual)
    B, Cin, H, W = x.shape
    taps, Cout, cin_w = w_packed.shape
    if taps != R * S or cin_w != Cin:
        raise BackendError(f"conv2d: input {tuple(x.shape)} incompatible with packed weight {tuple(w_packed.shape)}")
    Ho = (H + 2 * pad - R) // stride + 1
    Wo = (W + 2 * pad - S) // stride + 1
    xc = x.contiguous(memory_format=torch.channels_last)
    y = torch.empty((B, Cout, Ho, Wo), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    if residual is not None:
        if residual.shape != y.shape:
            raise BackendError("conv2d: residual shape mismatch")
        residual = residual.contiguous(memory_format=torch.channels_last)
    if chan_bias is not None:
        chan_bias = chan_bias.reshape(B, Cout).contiguous()
    if y.numel():
        _check(
            lib.rb200_conv2d(
                _stream(), _dtype_code(x), xc.data_ptr(), w_packed.data_ptr(), _ptr(bias), _ptr(chan_bias),
                _ptr(residual), y.data_ptr(), B, H, W, Cin, Cout, R, S, stride, pad, epilogue,
            )
        )
    return y


def _group_norm_impl(x: Tensor, groups: int, gamma: Tensor, beta: Tensor, eps: float, silu: bool) -> Tensor:
    lib = load_library()
    _same(x, gamma, beta)
    if x.ndim < 3:
        raise BackendError("group_norm expects [B, C, *spatial]")
    B, C = x.shape[0], x.shape[1]
    if x.ndim == 4:
        xc = x.contiguous(memory_format=torch.channels_last)
        y = torch.empty_like(xc, memory_format=torch.channels_last)
        HW = x.shape[2] * x.shape[3]
        xk, yk = xc, y
    else:  # [B, C, L]: normalise a channels-last copy, hand back the original layout
        xk = x.flatten(2).transpose(1, 2).contiguous()
        yk = torch.empty_like(xk)
        HW = xk.shape[1]
        y = None
    ws_bytes = lib.rb200_group_norm_workspace_bytes(B, HW, groups)
    ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
    if xk.numel():
        _check(
            lib.rb200_group_norm(
                _stream(), _dtype_code(x), xk.data_ptr(), yk.data_ptr(), B, HW, C, groups, float(eps),
                gamma.data_ptr(), beta.data_ptr(), int(silu), ws.data_ptr(), ws_bytes,
            )
        )
    if y is None:
        return yk.transpose(1, 2).reshape(x.shape)
    return y


def _layer_norm_impl(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    lib = load_library()
    _same(x, gamma, beta)
    C = x.shape[-1]
    xc = x if x.is_contiguous() else x.contiguous()
    y = torch.empty_like(xc)
    rows = xc.numel() // C if C else 0
    if rows:
        _check(
            lib.rb200_layer_norm(
                _stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), rows, C, float(eps), gamma.data_ptr(),
                beta.data_ptr(),
            )
        )
    return y


def _dense_like(x: Tensor) -> Tensor:
    """x itself when its memory is a dense permutation (contiguous or channels-last), else a copy."""
    if x.is_contiguous() or (x.ndim == 4 and x.is_contiguous(memory_format=torch.channels_last)):
        return x
    return x.contiguous()


def _unary_impl(x: Tensor, op: int) -> Tensor:
    lib = load_library()
    xc = _dense_like(x)
    y = torch.empty_like(xc)
    if xc.numel():
        _check(lib.rb200_unary(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), xc.numel(), op))
    return y


def _geglu_impl(x: Tensor) -> Tensor:
    lib = load_library()
    F2 = x.shape[-1]
    xc = x if x.is_contiguous() else x.contiguous()
    y = torch.empty((*x.shape[:-1], F2 // 2), device=x.device, dtype=x.dtype)
    rows = xc.numel() // F2 if F2 else 0
    if rows:
        _check(lib.rb200_geglu(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), rows, F2 // 2))
    return y


def _add_impl(a: Tensor, b: Tensor, alpha: float) -> Tensor:
    lib = load_library()
    _same(a, b)
    if a.shape != b.shape:
        raise BackendError("add: shapes differ (broadcasting stays in ATen)")
    ac = _dense_like(a)
    if b.stride() != ac.stride():
        b = b.contiguous(memory_format=torch.channels_last) if (
            ac.ndim == 4 and not ac.is_contiguous()
        ) else b.contiguous()
        if b.stride() != ac.stride():
            ac = ac.contiguous()
            b = b.contiguous()
    y = torch.empty_like(ac)
    if ac.numel():
        _check(lib.rb200_add(_stream(), _dtype_code(a), ac.data_ptr(), b.data_ptr(), y.data_ptr(), ac.numel(), alpha))
    return y


def _rows(t: Tensor) -> Tensor:
    """[B, S, C] view with unit channel stride (copy only if needed)."""
    return t if t.stride(-1) == 1 else t.contiguous()


def _sdpa_impl(
    q: Tensor, k: Tensor, v: Tensor, k2: Tensor | None, v2: Tensor | None, heads: int, causal: bool, scale2: float
) -> Tensor:
    lib = load_library()
    _same(q, k, v, k2, v2)
    B, Sq, C = q.shape
    Sk = k.shape[1]
    if C % heads or k.shape[2] != C or v.shape[2] != C or k.shape[0] != B or v.shape[0] != B or v.shape[1] != Sk:
        raise BackendError(f"sdpa: bad shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)} heads={heads}")
    D = C // heads
    q, k, v = _rows(q), _rows(k), _rows(v)
    o = torch.empty((B, Sq, C), device=q.device, dtype=q.dtype)
    Sk2 = 0
    if k2 is not None:
        assert v2 is not None
        k2, v2 = _rows(k2), _rows(v2)
        Sk2 = k2.shape[1]
    if o.numel():
        _check(
            lib.rb200_sdpa(
                _stream(), _dtype_code(q), q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), B, heads, Sq, Sk, D,
                q.stride(0), q.stride(1), k.stride(0), k.stride(1), v.stride(0), v.stride(1), o.stride(0), o.stride(1),
                float(D) ** -0.5, int(causal), _ptr(k2), _ptr(v2), Sk2,
                0 if k2 is None else k2.stride(0), 0 if k2 is None else k2.stride(1),
                0 if v2 is None else v2.stride(0), 0 if v2 is None else v2.stride(1), float(scale2),
            )
        )
    return o


def _sam_attention_impl(qkv: Tensor, rel_h: Tensor, rel_w: Tensor, heads: int) -> Tensor:
    lib = load_library()
    _same(qkv, rel_h, rel_w)
    Bw, Hh, Ww, C3 = qkv.shape
    C = C3 // 3
    d = C // heads
    qc = qkv if qkv.is_contiguous() else qkv.contiguous()
    o = torch.empty((Bw, Hh, Ww, C), device=qkv.device, dtype=qkv.dtype)
    ws_bytes = lib.rb200_sam_attention_workspace_bytes(Bw, Hh, Ww, heads, d)
    ws = torch.empty(ws_bytes, device=qkv.device, dtype=torch.uint8)
    if o.numel():
        _check(
            lib.rb200_sam_attention(
                _stream(), _dtype_code(qkv), qc.data_ptr(), rel_h.contiguous().data_ptr(),
                rel_w.contiguous().data_ptr(), o.data_ptr(), Bw, Hh, Ww, heads, d, ws.data_ptr(), ws_bytes,
            )
        )
    return o


def _patchify_impl(x: Tensor, patch: int) -> Tensor:
    lib = load_library()
    B, C, H, W = x.shape
    if H % patch or W % patch:
        raise BackendError(f"patchify: {H}x{W} is not a multiple of the patch size {patch}")
    y = torch.empty((B * (H // patch) * (W // patch), patch * patch * C), device=x.device, dtype=x.dtype)
    if y.numel():
        _check(
            lib.rb200_patchify(
                _stream(), _dtype_code(x), x.data_ptr(), y.data_ptr(), B, H, W, C, patch,
                x.stride(0), x.stride(1), x.stride(2), x.stride(3),
            )
        )
    return y


def _pad_channels_impl(x: Tensor, cp: int) -> Tensor:
    lib = load_library()
    B, C, H, W = x.shape
    y = torch.empty((B, cp, H, W), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    if y.numel():
        _check(
            lib.rb200_pad_channels(
                _stream(), _dtype_code(x), x.data_ptr(), y.data_ptr(), B, H, W, C, cp,
                x.stride(0), x.stride(1), x.stride(2), x.stride(3),
            )
        )
    return y


def _concat_channels_impl(parts: list[Tensor]) -> Tensor:
    lib = load_library()
    _same(*parts)
    B, _, H, W = parts[0].shape
    srcs = [p.contiguous(memory_format=torch.channels_last) for p in parts]
    for p in srcs:
        if p.shape[0] != B or p.shape[2] != H or p.shape[3] != W:
            raise BackendError(f"concat_channels: shapes differ outside the channel axis: {[tuple(t.shape) for t in parts]}")
    ct = sum(p.shape[1] for p in srcs)
    y = torch.empty((B, ct, H, W), device=parts[0].device, dtype=parts[0].dtype, memory_format=torch.channels_last)
    if y.numel():
        ptrs = (ctypes.c_void_p * len(srcs))(*[p.data_ptr() for p in srcs])
        chans = (ctypes.c_int * len(srcs))(*[p.shape[1] for p in srcs])
        _check(lib.rb200_concat_channels(_stream(), _dtype_code(y), len(srcs), ptrs, chans, y.data_ptr(), B * H * W))
    return y


def _resize_nearest_impl(x: Tensor, height: int, width: int) -> Tensor:
    lib = load_library()
    B, C, H, W = x.shape
    xc = x.contiguous(memory_format=torch.channels_last)
    y = torch.empty((B, C, height, width), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    if y.numel():
        _check(lib.rb200_resize_nearest(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), B, H, W, C, height, width))
    return y


def _window_partition_impl(x: Tensor, window: int) -> Tensor:
    lib = load_library()
    B, H, W, C = x.shape
    xc = x if x.is_contiguous() else x.contiguous()
    nh, nw = -(-H // window), -(-W // window)
    y = torch.empty((B * nh * nw, window, window, C), device=x.device, dtype=x.dtype)
    if y.numel():
        _check(lib.rb200_window_partition(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), B, H, W, C, window, 0))
    return y


def _window_merge_impl(x: Tensor, window: int, height: int, width: int) -> Tensor:
    lib = load_library()
    nh, nw = -(-height // window), -(-width // window)
    C = x.shape[-1]
    if x.shape[0] % (nh * nw) or x.shape[1] != window or x.shape[2] != window:
        raise BackendError(f"window_merge: {tuple(x.shape)} does not tile a {height}x{width} map with window {window}")
    B = x.shape[0] // (nh * nw)
    xc = x if x.is_contiguous() else x.contiguous()
    y = torch.empty((B, height, width, C), device=x.device, dtype=x.dtype)
    if y.numel():
        _check(lib.rb200_window_partition(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), B, height, width, C, window, 1))
    return y


def _on_operand_device(fn: Any) -> Any:
    """Run an op with the first tensor operand's device current: ``_stream()``, the allocations and the
    library's per-device state (shared-memory opt-in, SM count) all follow the operands, not whatever
    device happened to be current (a model on cuda:1, two GPUs in one process)."""

    def run(*args: Any) -> Any:
        first = args[0][0] if isinstance(args[0], (list, tuple)) else args[0]
        if first.device.index == torch.cuda.current_device():
            return fn(*args)
        with torch.cuda.device(first.device):
            return fn(*args)

    run.__name__ = fn.__name__
    return run


# --------------------------------------------------------------- torch custom-op registration
_torch_lib = torch.library.Library("refiners_b200", "DEF")
_torch_lib.define(
    "linear(Tensor x, Tensor w, Tensor? bias, Tensor? residual, Tensor? lora_down, Tensor? lora_up, "
    "Tensor? lora_scale, int epilogue) -> Tensor"
)
_torch_lib.define(
    "conv2d(Tensor x, Tensor w_packed, Tensor? bias, Tensor? chan_bias, Tensor? residual, int R, int S, "
    "int stride, int pad, int epilogue) -> Tensor"
)
_torch_lib.define("group_norm(Tensor x, int groups, Tensor gamma, Tensor beta, float eps, bool silu) -> Tensor")
_torch_lib.define("layer_norm(Tensor x, Tensor gamma, Tensor beta, float eps) -> Tensor")
_torch_lib.define("unary(Tensor x, int op) -> Tensor")
_torch_lib.define("geglu(Tensor x) -> Tensor")
_torch_lib.define("add(Tensor a, Tensor b, float alpha) -> Tensor")
_torch_lib.define(
    "sdpa(Tensor q, Tensor k, Tensor v, Tensor? k2, Tensor? v2, int heads, bool causal, float scale2) -> Tensor"
)
_torch_lib.define("sam_attention(Tensor qkv, Tensor rel_h, Tensor rel_w, int heads) -> Tensor")
_torch_lib.define("patchify(Tensor x, int patch) -> Tensor")
_torch_lib.define("pad_channels(Tensor x, int cp) -> Tensor")
_torch_lib.define("concat_channels(Tensor[] parts) -> Tensor")
_torch_lib.define("resize_nearest(Tensor x, int height, int width) -> Tensor")
_torch_lib.define("window_partition(Tensor x, int window) -> Tensor")
_torch_lib.define("window_merge(Tensor x, int window, int height, int width) -> Tensor")

for _name, _fn in (
    ("linear", _linear_impl),
    ("conv2d", _conv2d_impl),
    ("group_norm", _group_norm_impl),
    ("layer_norm", _layer_norm_impl),
    ("unary", _unary_impl),
    ("geglu", _geglu_impl),
    ("add", _add_impl),
    ("sdpa", _sdpa_impl),
    ("sam_attention", _sam_attention_impl),
    ("patchify", _patchify_impl),
    ("pad_channels", _pad_channels_impl),
    ("concat_channels", _concat_channels_impl),
    ("resize_nearest", _resize_nearest_impl),
    ("window_partition", _window_partition_impl),
    ("window_merge", _window_merge_impl),
):
    _torch_lib.impl(_name, _on_operand_device(_fn), "CUDA")


@torch.library.register_fake("refiners_b200::linear")
def _linear_fake(x, w, bias, residual, lora_down, lora_up, lora_scale, epilogue):  # type: ignore[no-untyped-def]
    n = w.shape[0] // 2 if epilogue == EPI_GEGLU else w.shape[0]
    return x.new_empty((*x.shape[:-1], n))


@torch.library.register_fake("refiners_b200::conv2d")
def _conv2d_fake(x, w_packed, bias, chan_bias, residual, R, S, stride, pad, epilogue):  # type: ignore[no-untyped-def]
    B, _, H, W = x.shape
    return x.new_empty((B, w_packed.shape[1], (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1))


@torch.library.register_fake("refiners_b200::group_norm")
def _group_norm_fake(x, groups, gamma, beta, eps, silu):  # type: ignore[no-untyped-def]
    return torch.empty_like(x)


@torch.library.register_fake("refiners_b200::layer_norm")
def _layer_norm_fake(x, gamma, beta, eps):  # type: ignore[no-untyped-def]
    return torch.empty_like(x)


@torch.library.register_fake("refiners_b200::unary")
def _unary_fake(x, op):  # type: ignore[no-untyped-def]
    return torch.empty_like(x)


@torch.library.register_fake("refiners_b200::geglu")
def _geglu_fake(x):  # type: ignore[no-untyped-def]
    return x.new_empty((*x.shape[:-1], x.shape[-1] // 2))


@torch.library.register_fake("refiners_b200::add")
def _add_fake(a, b, alpha):  # type: ignore[no-untyped-def]
    return torch.empty_like(a)


@torch.library.register_fake("refiners_b200::sdpa")
def _sdpa_fake(q, k, v, k2, v2, heads, causal, scale2):  # type: ignore[no-untyped-def]
    return torch.empty_like(q)


@torch.library.register_fake("refiners_b200::sam_attention")
def _sam_attention_fake(qkv, rel_h, rel_w, heads):  # type: ignore[no-untyped-def]
    return qkv.new_empty((*qkv.shape[:-1], qkv.shape[-1] // 3))


@torch.library.register_fake("refiners_b200::patchify")
def _patchify_fake(x, patch):  # type: ignore[no-untyped-def]
    B, C, H, W = x.shape
    return x.new_empty((B * (H // patch) * (W // patch), patch * patch * C))


@torch.library.register_fake("refiners_b200::pad_channels")
def _pad_channels_fake(x, cp):  # type: ignore[no-untyped-def]
    return x.new_empty((x.shape[0], cp, x.shape[2], x.shape[3]))


@torch.library.register_fake("refiners_b200::concat_channels")
def _concat_channels_fake(parts):  # type: ignore[no-untyped-def]
    p0 = parts[0]
    return p0.new_empty((p0.shape[0], sum(p.shape[1] for p in parts), p0.shape[2], p0.shape[3]))


@torch.library.register_fake("refiners_b200::resize_nearest")
def _resize_nearest_fake(x, height, width):  # type: ignore[no-untyped-def]
    return x.new_empty((x.shape[0], x.shape[1], height, width))


@torch.library.register_fake("refiners_b200::window_partition")
def _window_partition_fake(x, window):  # type: ignore[no-untyped-def]
    B, H, W, C = x.shape
    return x.new_empty((B * -(-H // window) * -(-W // window), window, window, C))


@torch.library.register_fake("refiners_b200::window_merge")
def _window_merge_fake(x, window, height, width):  # type: ignore[no-untyped-def]
    n = -(-height // window) * -(-width // window)
    return x.new_empty((x.shape[0] // n, height, width, x.shape[-1]))


_ops = torch.ops.refiners_b200


# ----------------------------------------------------------------------------- public helpers
def pack_loras(x: Tensor, w: Tensor, loras: Sequence[tuple[Tensor, Tensor, float]]) -> tuple[Tensor, Tensor, Tensor]:
    """(down_cat[r_pad, K], up_cat[N, r_pad], colscale[r_pad]) for rb200_linear, cached."""
    sources = [t for d, u, _ in loras for t in (d, u)]
    key = _PackCache.key(*sources) + tuple(float(s) for _, _, s in loras)
    hit = _lora_cache.get(key, sources)
    if hit is not None:
        return hit
    lib = load_library()
    N, K = w.shape
    total = sum(d.shape[0] for d, _, _ in loras)
    r_pad = (total + 63) // 64 * 64
    descs = (_LoraDesc * len(loras))()
    keep = []
    for i, (d, u, s) in enumerate(loras):
        d, u = d.contiguous(), u.contiguous()
        keep += [d, u]
        if d.shape[1] != K or u.shape[0] != N or u.shape[1] != d.shape[0]:
            raise BackendError(f"LoRA {i}: down{tuple(d.shape)} / up{tuple(u.shape)} do not fit Linear({K}->{N})")
        _same(w, d, u)
        descs[i] = _LoraDesc(d.data_ptr(), u.data_ptr(), float(s), d.shape[0])
    down_cat = torch.empty((r_pad, K), device=w.device, dtype=w.dtype)
    up_cat = torch.empty((N, r_pad), device=w.device, dtype=w.dtype)
    colscale = torch.empty((r_pad,), device=w.device, dtype=torch.float32)
    _check(
        lib.rb200_lora_pack(
            _stream(), _dtype_code(w), len(loras), descs, N, K, down_cat.data_ptr(), up_cat.data_ptr(),
            colscale.data_ptr(), r_pad,
        )
    )
    return _lora_cache.put(key, sources, (down_cat, up_cat, colscale))


def merged_lora_weight(weight: Tensor, loras: Sequence[tuple[Tensor, Tensor, float]]) -> Tensor:
    """``W + sum_i s_i B_i A_i`` ([N, K], the dtype of W), cached per (W, factors, scales).

    The scales and factors of a LoRA are step-invariant, so the adapted layer
    ``x W^T + sum_i s_i (x A_i^T) B_i^T`` (fluxion/adapters/lora.py:383-448 of the reference) is evaluated as ONE GEMM
    against the merged weight.  The merge itself is one launch of the library's GEMM: rows of ``[s_i B_i]`` times
    ``[A_i]^T`` with W as the epilogue residual, accumulated in fp32 and rounded once to the weight dtype - the same
    single rounding every weight of the model has already been through."""
    sources = [weight] + [t for d, u, _ in loras for t in (d, u)]
    key = _PackCache.key(*sources) + tuple(float(s) for _, _, s in loras)
    hit = _merge_cache.get(key, sources)
    if hit is not None:
        return hit
    N, K = weight.shape
    for i, (d, u, _) in enumerate(loras):
        if d.shape[1] != K or u.shape[0] != N or u.shape[1] != d.shape[0]:
            raise BackendError(f"LoRA {i}: down{tuple(d.shape)} / up{tuple(u.shape)} do not fit Linear({K}->{N})")
        _same(weight, d, u)
    total = sum(d.shape[0] for d, _, _ in loras)
    r_pad = (total + 15) // 16 * 16
    up_scaled = weight.new_zeros((N, r_pad))
    down_t = weight.new_zeros((K, r_pad))
    at = 0
    for d, u, s in loras:  # one-time packing glue (a few MB); the product below is the library's own GEMM
        r = d.shape[0]
        up_scaled[:, at : at + r] = u.detach() * float(s)
        down_t[:, at : at + r] = d.detach().t()
        at += r
    w = weight.detach()
    merged = _ops.linear(up_scaled, down_t, None, w if w.stride(-1) == 1 else w.contiguous(), None, None, None, EPI_NONE)
    return _merge_cache.put(key, sources, merged)


def lora_fusable(x: Tensor, loras: Sequence[tuple[Tensor, Tensor, float]]) -> bool:
    return len(loras) > 0 and all(d.is_cuda and u.is_cuda and d.dtype == x.dtype for d, u, _ in loras)


def linear(
    x: Tensor,
    weight: Tensor,
    bias: Tensor | None = None,
    *,
    residual: Tensor | None = None,
    loras: Sequence[tuple[Tensor, Tensor, float]] = (),
    epilogue: int = EPI_NONE,
) -> Tensor:
    _inference_only(x, weight, bias)
    down = up = scale = None
    if loras:
        if _lora_merge:
            weight = merged_lora_weight(weight, loras)
        else:
            down, up, scale = pack_loras(x, weight, loras)
    if _memo is not None and residual is None and InvariantMemo.marked(x):
        return _memo.fetch(("linear", id(x), id(weight), id(bias), id(down), epilogue),
                           lambda: _ops.linear(x, weight, bias, None, down, up, scale, epilogue))
    return _ops.linear(x, weight, bias, residual, down, up, scale, epilogue)


def linear_geglu(x: Tensor, weight: Tensor, bias: Tensor | None) -> Tensor:
    """``GLU(GeLU)(Linear(x))`` in one launch; the value/gate interleave of W is cached."""
    _inference_only(x, weight, bias)
    key = _PackCache.key(weight, bias)
    packed = _geglu_cache.get(key, (weight, bias))
    if packed is None:
        lib = load_library()
        w = weight.contiguous()
        wp = torch.empty_like(w)
        bp = None if bias is None else torch.empty_like(bias)
        _check(
            lib.rb200_geglu_pack(
                _stream(), _dtype_code(w), w.data_ptr(), _ptr(bias), wp.data_ptr(), _ptr(bp), w.shape[0] // 2, w.shape[1]
            )
        )
        packed = _geglu_cache.put(key, (weight, bias), (wp, bp))
    wp, bp = packed
    return _ops.linear(x, wp, bp, None, None, None, None, EPI_GEGLU)


def concat_linear_weights(weights: Sequence[Tensor], biases: Sequence[Tensor | None]) -> tuple[Tensor, Tensor | None]:
    """Row-concatenation of sibling Linear weights (fused q/k/v projection), cached per source tensors."""
    sources = list(weights) + list(biases)
    key = _PackCache.key(*sources)
    hit = _concat_cache.get(key, sources)
    if hit is not None:
        return hit
    w = torch.cat([t.detach() for t in weights], dim=0).contiguous()
    b = None if biases[0] is None else torch.cat([t.detach() for t in biases if t is not None], dim=0).contiguous()
    return _concat_cache.put(key, sources, (w, b))


def geglu_fusable(weight: Tensor) -> bool:
    return weight.dtype in (torch.bfloat16, torch.float16) and (weight.shape[0] // 2) % 16 == 0 and weight.shape[1] % 8 == 0


def packed_conv_weight(weight: Tensor, cin_padded: int | None = None) -> Tensor:
    """[Cout, Cin, R, S] -> [R*S, Cout, Cin] (cached per weight object); with ``cin_padded`` the input-channel
    axis is zero-extended to that width (pairs with ``pad_channels`` on the activation)."""
    cache = _conv_cache if cin_padded is None else _conv_pad_cache
    key = _PackCache.key(weight)
    packed = cache.get(key, (weight,))
    if packed is None:
        lib = load_library()
        w = weight.contiguous()
        Cout, Cin, R, S = w.shape
        packed = torch.empty((R * S, Cout, Cin), device=w.device, dtype=w.dtype)
        _check(lib.rb200_conv2d_pack_weight(_stream(), _dtype_code(w), w.data_ptr(), packed.data_ptr(), Cout, Cin, R, S))
        if cin_padded is not None:
            wide = torch.zeros((R * S, Cout, cin_padded), device=w.device, dtype=w.dtype)
            wide[:, :, :Cin] = packed
            packed = wide
        cache.put(key, (weight,), packed)
    return packed


def conv2d(
    x: Tensor,
    weight: Tensor,
    bias: Tensor | None,
    stride: int,
    padding: int,
    *,
    chan_bias: Tensor | None = None,
    residual: Tensor | None = None,
    epilogue: int = EPI_NONE,
) -> Tensor:
    _inference_only(x, weight, bias)
    if _memo is not None and chan_bias is None and residual is None and InvariantMemo.marked(x):
        memo = _memo
        key = ("conv2d", id(x), id(weight), id(bias), stride, padding, epilogue)
        if key in memo.entries:
            return memo.fetch(key, None)

        def run() -> Tensor:
            prev = set_invariant_memo(None)  # the body below must launch, not look itself up
            try:
                return conv2d(x, weight, bias, stride, padding, epilogue=epilogue)
            finally:
                set_invariant_memo(prev)

        return memo.fetch(key, run)
    R, S = weight.shape[2], weight.shape[3]
    if (
        R == S == stride
        and R > 1
        and padding == 0
        and chan_bias is None
        and residual is None
        and x.shape[2] % R == 0
        and x.shape[3] % R == 0
    ):
        # kernel = stride (ViT patch embedding): a plain GEMM over non-overlapping patches
        Bn, _, H, W = x.shape
        y = _ops.linear(_ops.patchify(x, R), patch_gemm_weight(weight), bias, None, None, None, None, epilogue)
        return y.view(Bn, H // R, W // R, weight.shape[0]).permute(0, 3, 1, 2)
    cin = weight.shape[1]
    if cin % 8 != 0 and x.dtype in (torch.bfloat16, torch.float16):
        # narrow input convs (4 latent channels, 3 image channels): zero-extend to 8 channels for the tcgen05 path
        cp = (cin + 7) // 8 * 8
        return _ops.conv2d(_ops.pad_channels(x, cp), packed_conv_weight(weight, cp), bias, chan_bias, residual, R, S, stride, padding, epilogue)
    return _ops.conv2d(x, packed_conv_weight(weight), bias, chan_bias, residual, R, S, stride, padding, epilogue)


def patch_gemm_weight(weight: Tensor) -> Tensor:
    """[Cout, Cin, P, P] -> [Cout, (r, s, c)] matching rb200_patchify's row layout (cached)."""
    key = _PackCache.key(weight)
    packed = _patch_cache.get(key, (weight,))
    if packed is None:
        packed = weight.detach().permute(0, 2, 3, 1).reshape(weight.shape[0], -1).contiguous()
        _patch_cache.put(key, (weight,), packed)
    return packed


def concat_channels_supported(parts: Sequence[Any]) -> bool:
    """dim-1 concatenation of 2..4 same-dtype 4-D CUDA maps whose channel counts fill whole 16-byte vectors."""
    if not 2 <= len(parts) <= 4 or not all(isinstance(p, Tensor) and p.is_cuda and p.ndim == 4 for p in parts):
        return False
    p0 = parts[0]
    if p0.dtype not in _DT or any(p.dtype != p0.dtype or p.device != p0.device for p in parts):
        return False
    vec = 16 // p0.element_size()
    return all(p.shape[1] % vec == 0 and p.shape[0] == p0.shape[0] and p.shape[2:] == p0.shape[2:] for p in parts)


def concat_channels(parts: Sequence[Tensor]) -> Tensor:
    _inference_only(*parts)
    return _ops.concat_channels(list(parts))


def resize_nearest_supported(x: Tensor) -> bool:
    return x.is_cuda and x.ndim == 4 and x.dtype in _DT and x.shape[1] % (16 // x.element_size()) == 0


def style_aligned(x: Tensor, *, adain: bool, concatenate: bool, scale: float, epsilon: float) -> Tensor:
    """One `StyleAligned` chain on q, k or v ``[B, S, C]`` of a guidance batch -> ``[B, S or 2S, C]``; see rb200_style_aligned."""
    _inference_only(x)
    lib = load_library()
    B, S, C = x.shape
    if x.stride(2) != 1:
        x = x.contiguous()
    y = torch.empty((B, 2 * S if concatenate else S, C), device=x.device, dtype=x.dtype)
    with torch.cuda.device(x.device):
        ws_bytes = lib.rb200_style_aligned_workspace_bytes(B, C) if adain else 0
        ws = torch.empty(max(ws_bytes, 1), device=x.device, dtype=torch.uint8)
        if y.numel():
            _check(lib.rb200_style_aligned(_stream(), _dtype_code(x), x.data_ptr(), y.data_ptr(), B, S, C, x.stride(0), x.stride(1),
                                           int(adain), int(concatenate), float(scale), float(epsilon), ws.data_ptr(), ws_bytes))
    return y


def avg_pool2d(x: Tensor, k: int) -> Tensor:
    """k x k average pooling with stride k of an NCHW map (result channels-last); see rb200_avg_pool2d."""
    _inference_only(x)
    lib = load_library()
    B, C, H, W = x.shape
    if C % (16 // x.element_size()):
        raise BackendError(f"avg_pool2d: {C} channels are not a whole number of 16-byte vectors")
    xc = x.contiguous(memory_format=torch.channels_last)
    y = torch.empty((B, C, H // k, W // k), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    if y.numel():
        with torch.cuda.device(x.device):
            _check(lib.rb200_avg_pool2d(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), B, H, W, C, int(k)))
    return y


def resize_nearest(x: Tensor, height: int, width: int) -> Tensor:
    _inference_only(x)
    return _ops.resize_nearest(x, int(height), int(width))


def window_partition(x: Tensor, window: int) -> Tensor:
    _inference_only(x)
    return _ops.window_partition(x, window)


def window_merge(x: Tensor, window: int, height: int, width: int) -> Tensor:
    _inference_only(x)
    return _ops.window_merge(x, window, height, width)


def conv_supported(module: Any) -> bool:
    s, p, d = module.stride, module.padding, module.dilation
    return (
        module.groups == 1
        and module.padding_mode == "zeros"
        and not isinstance(p, str)
        and s[0] == s[1]
        and p[0] == p[1]
        and tuple(d) == (1, 1)
    )


def conv2d_module(x: Tensor, module: Any, **fused: Any) -> Tensor:
    """Run a fluxion ``Conv2d`` leaf.  Unsupported conv flavours (groups, dilation, string or
    asymmetric padding) are not on the hot path and are refused rather than silently rerouted."""
    if not conv_supported(module):
        raise BackendError(
            f"conv2d: unsupported configuration on CUDA (groups={module.groups}, dilation={module.dilation}, "
            f"padding={module.padding}, stride={module.stride}, padding_mode={module.padding_mode})"
        )
    return conv2d(x, module.weight, module.bias, int(module.stride[0]), int(module.padding[0]), **fused)


def group_norm(x: Tensor, groups: int, gamma: Tensor, beta: Tensor, eps: float, silu: bool = False) -> Tensor:
    _inference_only(x, gamma, beta)
    return _ops.group_norm(x, groups, gamma, beta, eps, silu)


def group_norm_fixed(
    x: Tensor, groups: int, gamma: Tensor, beta: Tensor, eps: float, stats: Tensor | None = None
) -> tuple[Tensor, Tensor]:
    """GroupNorm of an NCHW map with statistics that are captured once and then frozen (tiled VAE inference; see
    rb200_group_norm_fixed).  ``stats`` is None on the capturing pass; the fp32 ``[B, groups, 2]`` (mean, rstd) buffer
    this returns is handed back on every later call."""
    _inference_only(x, gamma, beta)
    _same(x, gamma, beta)
    if x.ndim != 4:
        raise BackendError("group_norm_fixed expects [B, C, H, W]")
    lib = load_library()
    B, C, H, W = x.shape
    frozen = stats is not None
    if stats is None:
        stats = torch.empty((B, groups, 2), device=x.device, dtype=torch.float32)
    elif stats.shape != (B, groups, 2) or stats.dtype != torch.float32 or stats.device != x.device or not stats.is_contiguous():
        raise BackendError(f"group_norm_fixed: statistics {tuple(stats.shape)} do not fit a batch of {B} with {groups} groups")
    xc = x.contiguous(memory_format=torch.channels_last)
    y = torch.empty_like(xc, memory_format=torch.channels_last)
    with torch.cuda.device(x.device):
        ws_bytes = lib.rb200_group_norm_workspace_bytes(B, H * W, groups)
        ws = torch.empty(ws_bytes, device=x.device, dtype=torch.uint8)
        if xc.numel():
            _check(lib.rb200_group_norm_fixed(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), B, H * W, C, groups, float(eps),
                                              gamma.data_ptr(), beta.data_ptr(), 0, ws.data_ptr(), ws_bytes, stats.data_ptr(), int(frozen)))
    return y, stats


def layer_norm(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    _inference_only(x, gamma, beta)
    return _ops.layer_norm(x, gamma, beta, eps)


def layer_norm_2d(x: Tensor, gamma: Tensor, beta: Tensor, eps: float) -> Tensor:
    """Channel-wise LN of an NCHW map = row LN over the NHWC pixels."""
    _inference_only(x, gamma, beta)
    nhwc = x.permute(0, 2, 3, 1)
    return _ops.layer_norm(nhwc, gamma, beta, eps).permute(0, 3, 1, 2)


def unary(x: Tensor, name: str) -> Tensor:
    _inference_only(x)
    if _memo is not None and InvariantMemo.marked(x):
        return _memo.fetch(("unary", id(x), name), lambda: _ops.unary(x, _UNARY[name]))
    return _ops.unary(x, _UNARY[name])


def geglu(x: Tensor) -> Tensor:
    _inference_only(x)
    return _ops.geglu(x)


def add(a: Tensor, b: Tensor, alpha: float = 1.0) -> Tensor:
    _inference_only(a, b)
    return _ops.add(a, b, alpha)


def cfg_scale_input(x: Tensor, sigmas: Tensor, step: int, twice: bool) -> Tensor:
    """``cat((x, x)) / ((sigmas[step]^2 + 1)^0.5)`` (or without the doubling) in one launch; see rb200_cfg_scale_input."""
    _inference_only(x)
    _same(x, sigmas)
    lib = load_library()
    xc = x.contiguous()
    y = torch.empty(((2 if twice else 1) * xc.shape[0], *xc.shape[1:]), device=x.device, dtype=x.dtype)
    sigma = sigmas[step : step + 1]
    with torch.cuda.device(x.device):
        _check(lib.rb200_cfg_scale_input(_stream(), _dtype_code(x), xc.data_ptr(), y.data_ptr(), xc.numel(), sigma.data_ptr(), int(twice)))
    return y


def cfg_euler(x: Tensor, eps: Tensor, sigmas: Tensor, step: int, condition_scale: float, guided: bool) -> Tensor:
    """``x + (u + s (c - u)) * (sigmas[step + 1] - sigmas[step])`` in one launch; see rb200_cfg_euler."""
    _inference_only(x, eps)
    _same(x, eps, sigmas)
    lib = load_library()
    xc, ec = x.contiguous(), eps.contiguous()
    if ec.numel() != (2 if guided else 1) * xc.numel():
        raise BackendError(f"cfg_euler: noise prediction {tuple(eps.shape)} does not match latents {tuple(x.shape)} (guided={guided})")
    y = torch.empty_like(xc)
    pair = sigmas[step : step + 2]
    with torch.cuda.device(x.device):
        _check(lib.rb200_cfg_euler(_stream(), _dtype_code(x), xc.data_ptr(), ec.data_ptr(), y.data_ptr(), xc.numel(), pair.data_ptr(),
                                   float(condition_scale), int(guided)))
    return y


def sdpa(
    q: Tensor,
    k: Tensor,
    v: Tensor,
    num_heads: int,
    is_causal: bool = False,
    *,
    k2: Tensor | None = None,
    v2: Tensor | None = None,
    scale2: float = 0.0,
) -> Tensor:
    _inference_only(q, k, v)
    return _ops.sdpa(q, k, v, k2, v2, num_heads, is_causal, scale2)


def attention_probs(q: Tensor, k: Tensor, num_heads: int) -> Tensor:
    """``softmax(q k^T / sqrt(d))`` as a ``[B, heads, Sq, Sk]`` tensor (self-attention guidance); see rb200_attention_probs."""
    _inference_only(q, k)
    _same(q, k)
    lib = load_library()
    B, Sq, C = q.shape
    if C % num_heads or k.shape[0] != B or k.shape[2] != C:
        raise BackendError(f"attention_probs: bad shapes q{tuple(q.shape)} k{tuple(k.shape)} heads={num_heads}")
    D, Sk = C // num_heads, k.shape[1]
    q, k = _rows(q), _rows(k)
    probs = torch.empty((B, num_heads, Sq, Sk), device=q.device, dtype=q.dtype)
    if probs.numel():
        with torch.cuda.device(q.device):
            _check(lib.rb200_attention_probs(_stream(), _dtype_code(q), q.data_ptr(), k.data_ptr(), probs.data_ptr(), B, num_heads, Sq, Sk, D,
                                             q.stride(0), q.stride(1), k.stride(0), k.stride(1), float(D) ** -0.5))
    return probs


def sam_attention(qkv: Tensor, rel_h: Tensor, rel_w: Tensor, num_heads: int) -> Tensor:
    _inference_only(qkv, rel_h, rel_w)
    return _ops.sam_attention(qkv, rel_h, rel_w, num_heads)
