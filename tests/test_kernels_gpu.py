"""GPU parity of every C-ABI kernel against a plain PyTorch fp32 evaluation of the same op on
the CPU (inputs pre-rounded to the kernel dtype).  Both kernel families are exercised:
mode 0 = auto (tcgen05 wherever the shape allows) and mode 1 = CUDA-core kernels only.

Tolerance: eps(dtype) * max|ref| (tests/helpers.py) - i.e. output rounding + fp32 accumulation
order; stated per test where a longer op chain needs a small multiple.
"""

import math

import pytest
import torch
import torch.nn.functional as F

from tests.helpers import assert_close, rounded

pytestmark = pytest.mark.gpu

DTYPES = [torch.bfloat16, torch.float16, torch.float32]


@pytest.fixture(params=[0, 1], ids=["auto", "simt"])
def kernel_mode(request, cuda_device):
    from refiners_b200 import backend as B

    prev = B.set_kernel_mode(request.param)
    yield request.param
    B.set_kernel_mode(prev)


def _gen(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


LINEAR_SHAPES = [
    # (M, K, N)
    (256, 128, 256),
    (1000, 320, 1280),   # ragged M
    (16, 1280, 1280),    # timestep MLP: M smaller than a tile
    (1232, 2048, 640),   # text K/V projection (77 * 16 rows)
    (384, 64, 96),       # N not a multiple of 32
    (130, 72, 40),       # K not a multiple of 64, N tail
    (64, 36, 24),        # K % 8 != 0 -> CUDA-core path even in auto mode
    (2048, 1280, 2560),  # several waves of 128x256 tiles
]


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("shape", LINEAR_SHAPES, ids=str)
def test_linear_bias(cuda_device, kernel_mode, dtype, shape):
    from refiners_b200 import backend as B

    M, K, N = shape
    x, w, b = _gen((M, K), 1), _gen((N, K), 2, K**-0.5), _gen((N,), 3)
    ref = F.linear(rounded(x, dtype), rounded(w, dtype), rounded(b, dtype))
    with torch.no_grad():
        y = B.linear(x.to(cuda_device, dtype), w.to(cuda_device, dtype), b.to(cuda_device, dtype))
    assert_close(y, ref, dtype, what=f"linear{shape}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=str)
def test_linear_3d_no_bias_residual(cuda_device, kernel_mode, dtype):
    from refiners_b200 import backend as B

    x, w, r = _gen((3, 100, 256), 4), _gen((512, 256), 5, 1 / 16), _gen((3, 100, 512), 6)
    ref = F.linear(rounded(x, dtype), rounded(w, dtype)) + rounded(r, dtype)
    with torch.no_grad():
        y = B.linear(x.to(cuda_device, dtype), w.to(cuda_device, dtype), None, residual=r.to(cuda_device, dtype))
    assert y.shape == (3, 100, 512)
    assert_close(y, ref, dtype, what="linear+residual")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("epi", ["gelu", "silu"])
def test_linear_activation_epilogue(cuda_device, kernel_mode, dtype, epi):
    from refiners_b200 import backend as B

    x, w, b = _gen((300, 256), 7), _gen((192, 256), 8, 1 / 16), _gen((192,), 9)
    pre = F.linear(rounded(x, dtype), rounded(w, dtype), rounded(b, dtype))
    ref = F.gelu(pre) if epi == "gelu" else F.silu(pre)
    code = B.EPI_GELU if epi == "gelu" else B.EPI_SILU
    with torch.no_grad():
        y = B.linear(x.to(cuda_device, dtype), w.to(cuda_device, dtype), b.to(cuda_device, dtype), epilogue=code)
    assert_close(y, ref, dtype, what=f"linear+{epi}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("shape", [(200, 128, 64), (1024, 1280, 5120), (77, 320, 1280)], ids=str)
def test_linear_geglu(cuda_device, kernel_mode, dtype, shape):
    """Linear(K -> 2F) + GLU(GeLU) fused (cross_attention.py:69-71 in the reference)."""
    from refiners_b200 import backend as B

    M, K, Fh = shape
    x, w, b = _gen((M, K), 10), _gen((2 * Fh, K), 11, K**-0.5), _gen((2 * Fh,), 12)
    pre = F.linear(rounded(x, dtype), rounded(w, dtype), rounded(b, dtype))
    a, g = pre.chunk(2, dim=-1)
    ref = a * F.gelu(g)
    with torch.no_grad():
        y = B.linear_geglu(x.to(cuda_device, dtype), w.to(cuda_device, dtype), b.to(cuda_device, dtype))
    assert y.shape == (M, Fh)
    assert_close(y, ref, dtype, scale=2.0, what=f"geglu{shape}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=str)
@pytest.mark.parametrize("ranks", [(16,), (16, 16), (4, 8, 32)], ids=str)
@pytest.mark.parametrize("merge", [True, False], ids=["merged", "two-launch"])
def test_linear_lora(cuda_device, kernel_mode, dtype, ranks, merge):
    """y = x W^T + b + sum_i s_i (x A_i^T) B_i^T (lora.py:383-448 in the reference), fp32 math.  Both evaluations:
    one GEMM against the cached merged weight W + sum_i s_i B_i A_i (default), and rank-space GEMM + base GEMM with the
    up-projection as extra k-blocks."""
    from refiners_b200 import backend as B

    M, K, N = 500, 320, 640
    x, w, b = _gen((M, K), 13), _gen((N, K), 14, K**-0.5), _gen((N,), 15)
    ref = F.linear(rounded(x, dtype), rounded(w, dtype), rounded(b, dtype))
    loras = []
    for i, r in enumerate(ranks):
        down, up, s = _gen((r, K), 20 + i, 1 / r), _gen((N, r), 30 + i, 0.05), 0.5 + 0.7 * i
        ref = ref + s * F.linear(F.linear(rounded(x, dtype), rounded(down, dtype)), rounded(up, dtype))
        loras.append((down.to(cuda_device, dtype), up.to(cuda_device, dtype), s))
    prev = B.set_lora_merge(merge)
    try:
        with torch.no_grad():
            wd = w.to(cuda_device, dtype)
            before = B.launch_count()
            y = B.linear(x.to(cuda_device, dtype), wd, b.to(cuda_device, dtype), loras=loras)
            first = B.launch_count() - before
            before = B.launch_count()
            y2 = B.linear(x.to(cuda_device, dtype), wd, b.to(cuda_device, dtype), loras=loras)
            again = B.launch_count() - before
    finally:
        B.set_lora_merge(prev)
    # the rank-space intermediate (two-launch) / the merged weight (merged) is rounded to the operand dtype once: 3 eps
    assert_close(y, ref, dtype, scale=3.0, what=f"lora{ranks}")
    assert torch.equal(y, y2)
    assert again == (1 if merge else 2), f"{again} launches per adapted Linear (first call: {first})"


CONV_CASES = [
    # (B, Cin, H, W, Cout, k, stride, pad)
    (2, 64, 16, 16, 128, 3, 1, 1),
    (1, 320, 32, 32, 320, 3, 1, 1),
    (2, 320, 32, 32, 640, 1, 1, 0),
    (2, 64, 32, 32, 64, 3, 2, 1),     # Downsample (sampling.py:41-98)
    (2, 4, 32, 32, 320, 3, 1, 1),     # UNet input conv: Cin = 4 -> channels zero-extended to 8 for the tcgen05 path
    (1, 3, 40, 24, 16, 3, 1, 1),      # ConditionEncoder stem: Cin = 3, Cout = 16
    (2, 12, 16, 16, 32, 3, 2, 1),     # Cin = 12 -> 16, stride 2
    (2, 320, 16, 16, 4, 3, 1, 1),     # UNet output conv: Cout = 4
    (4, 128, 8, 8, 128, 3, 1, 1),     # 8x8 map: one tile spans two images
    (1, 96, 24, 40, 80, 3, 1, 1),     # odd geometry
    (1, 8, 64, 64, 32, 16, 16, 0),    # patch embedding style (kernel = stride = 16): patchify + GEMM
    (2, 3, 64, 96, 160, 16, 16, 0),   # SAM PatchEncoder: Cin = 3
    (1, 5, 12, 12, 7, 4, 4, 0),       # odd channel counts -> CUDA-core GEMM
]


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("case", CONV_CASES, ids=str)
def test_conv2d(cuda_device, kernel_mode, dtype, case):
    from refiners_b200 import backend as B

    Bn, Cin, H, W, Cout, k, stride, pad = case
    x = _gen((Bn, Cin, H, W), 40)
    w = _gen((Cout, Cin, k, k), 41, (Cin * k * k) ** -0.5)
    b = _gen((Cout,), 42)
    ref = F.conv2d(rounded(x, dtype), rounded(w, dtype), rounded(b, dtype), stride=stride, padding=pad)
    with torch.no_grad():
        y = B.conv2d(x.to(cuda_device, dtype), w.to(cuda_device, dtype), b.to(cuda_device, dtype), stride, pad)
    assert y.shape == ref.shape
    assert_close(y, ref, dtype, what=f"conv{case}")


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("geom", [(2, 64, 64, 32, 14), (1, 9, 20, 16, 7), (3, 8, 8, 8, 8), (1, 5, 3, 8, 4)], ids=str)
def test_window_partition_merge(cuda_device, dtype, geom):
    """WindowPartition / WindowMerge of the SAM encoder (image_encoder.py:202-237 in the reference)."""
    from refiners_b200 import backend as B

    Bn, H, W, C, ws = geom
    x = _gen((Bn, H, W, C), 48).to(dtype)
    ph, pw = (ws - H % ws) % ws, (ws - W % ws) % ws
    xp = F.pad(x, (0, 0, 0, pw, 0, ph))
    hp, wp = H + ph, W + pw
    ref = xp.view(Bn, hp // ws, ws, wp // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws, ws, C)
    with torch.no_grad():
        y = B.window_partition(x.to(cuda_device), ws)
        back = B.window_merge(y, ws, H, W)
    assert torch.equal(y.cpu(), ref)
    assert torch.equal(back.cpu(), x)


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
def test_concat_channels_and_resize_nearest(cuda_device, dtype):
    """Skip-connection plumbing of the UNet: Concatenate(dim=1) (chain.py:930-964 in the reference) and the
    nearest-neighbour Interpolate (sampling.py:13-38) - pure data movement, bit exact."""
    from refiners_b200 import backend as B
    import refiners_b200.fluxion.layers as fl

    a, b, c = _gen((2, 64, 9, 7), 50).to(dtype), _gen((2, 32, 9, 7), 51).to(dtype), _gen((2, 8, 9, 7), 52).to(dtype)
    dev = lambda t: t.to(cuda_device).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        y = B.concat_channels([dev(a), dev(b), dev(c)])
        z = fl.Concatenate(fl.Identity(), fl.Identity(), dim=1)(dev(a))
    assert torch.equal(y.cpu(), torch.cat([a, b, c], dim=1)) and torch.equal(z.cpu(), torch.cat([a, a], dim=1))
    for size in ((18, 14), (13, 20), (9, 7), (4, 3)):
        with torch.no_grad():
            r = fl.Interpolate()(dev(a), torch.Size(size))
        assert torch.equal(r.cpu(), F.interpolate(a.float(), size=size, mode="nearest").to(dtype)), size


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=str)
def test_conv2d_fused_terms(cuda_device, kernel_mode, dtype):
    """conv + per-sample channel bias (RangeAdapter2d) + residual (ResidualBlock shortcut)."""
    from refiners_b200 import backend as B

    x, w, b = _gen((2, 64, 16, 16), 43), _gen((128, 64, 3, 3), 44, 1 / 24), _gen((128,), 45)
    cb, r = _gen((2, 128), 46), _gen((2, 128, 16, 16), 47)
    ref = F.conv2d(rounded(x, dtype), rounded(w, dtype), rounded(b, dtype), padding=1)
    ref = ref + rounded(cb, dtype)[:, :, None, None] + rounded(r, dtype)
    dev = lambda t: t.to(cuda_device, dtype)
    with torch.no_grad():
        y = B.conv2d(dev(x), dev(w), dev(b), 1, 1, chan_bias=dev(cb), residual=dev(r))
    assert_close(y, ref, dtype, what="conv+chan_bias+residual")


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("silu", [False, True])
@pytest.mark.parametrize("shape", [(2, 320, 16, 16), (1, 64, 7, 9), (2, 1920, 8, 8), (3, 32, 5, 5)], ids=str)
def test_group_norm(cuda_device, dtype, silu, shape):
    from refiners_b200 import backend as B

    x = _gen(shape, 50) * 2 + 0.5
    C = shape[1]
    g, b = _gen((C,), 51) * 0.2 + 1, _gen((C,), 52) * 0.2
    ref = F.group_norm(rounded(x, dtype), 32, rounded(g, dtype), rounded(b, dtype), eps=1e-5)
    if silu:
        ref = F.silu(ref)
    with torch.no_grad():
        y = B.group_norm(x.to(cuda_device, dtype), 32, g.to(cuda_device, dtype), b.to(cuda_device, dtype), 1e-5, silu=silu)
    assert y.shape == ref.shape
    assert_close(y, ref, dtype, scale=2.0, what=f"group_norm{shape}")


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("shape", [(2, 77, 1280), (5, 640), (3, 10, 100)], ids=str)
def test_layer_norm(cuda_device, dtype, shape):
    from refiners_b200 import backend as B

    x = _gen(shape, 53) * 3 + 1
    C = shape[-1]
    g, b = _gen((C,), 54) * 0.2 + 1, _gen((C,), 55) * 0.2
    ref = F.layer_norm(rounded(x, dtype), (C,), rounded(g, dtype), rounded(b, dtype), eps=1e-5)
    with torch.no_grad():
        y = B.layer_norm(x.to(cuda_device, dtype), g.to(cuda_device, dtype), b.to(cuda_device, dtype), 1e-5)
    assert_close(y, ref, dtype, scale=2.0, what=f"layer_norm{shape}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=str)
def test_layer_norm_2d(cuda_device, dtype):
    from refiners_b200 import backend as B

    x = _gen((2, 256, 6, 5), 56)
    g, b = _gen((256,), 57) * 0.2 + 1, _gen((256,), 58) * 0.2
    xr = rounded(x, dtype)
    mu = xr.mean(1, keepdim=True)
    var = (xr - mu).pow(2).mean(1, keepdim=True)
    ref = rounded(g, dtype)[:, None, None] * ((xr - mu) / torch.sqrt(var + 1e-6)) + rounded(b, dtype)[:, None, None]
    with torch.no_grad():
        y = B.layer_norm_2d(x.to(cuda_device, dtype), g.to(cuda_device, dtype), b.to(cuda_device, dtype), 1e-6)
    assert y.shape == x.shape
    assert_close(y, ref, dtype, scale=2.0, what="layer_norm_2d")


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("op", ["silu", "gelu", "gelu_tanh", "gelu_sigmoid", "relu", "sigmoid"])
def test_unary(cuda_device, dtype, op):
    from refiners_b200 import backend as B

    x = _gen((3, 1001), 59) * 3
    xr = rounded(x, dtype)
    ref = {
        "silu": F.silu(xr), "gelu": F.gelu(xr), "gelu_tanh": F.gelu(xr, approximate="tanh"),
        "gelu_sigmoid": xr * torch.sigmoid(1.702 * xr), "relu": F.relu(xr), "sigmoid": torch.sigmoid(xr),
    }[op]
    with torch.no_grad():
        y = B.unary(x.to(cuda_device, dtype), op)
    assert_close(y, ref, dtype, scale=2.0, what=op)


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
def test_geglu_and_add(cuda_device, dtype):
    from refiners_b200 import backend as B

    x = _gen((4, 33, 2 * 120), 60)
    a, g = rounded(x, dtype).chunk(2, dim=-1)
    with torch.no_grad():
        y = B.geglu(x.to(cuda_device, dtype))
    assert_close(y, a * F.gelu(g), dtype, scale=2.0, what="geglu")
    p, q = _gen((2, 64, 9, 9), 61), _gen((2, 64, 9, 9), 62)
    with torch.no_grad():
        s = B.add(p.to(cuda_device, dtype), q.to(cuda_device, dtype), 0.5)
    assert_close(s, rounded(p, dtype) + 0.5 * rounded(q, dtype), dtype, what="add")


SDPA_CASES = [
    # (B, H, Sq, Sk, D)
    (2, 4, 64, 64, 64),
    (1, 10, 256, 256, 64),
    (2, 5, 200, 77, 64),     # cross-attention on 77 text tokens, ragged Sq
    (1, 8, 96, 50, 40),      # SD1 head dims
    (1, 8, 40, 40, 80),
    (1, 2, 33, 65, 160),
    (1, 3, 300, 333, 80),    # two 64-column slabs, zero-filled past 80, ragged tiles
    (2, 2, 130, 200, 128),
    (1, 2, 70, 70, 8),
    (1, 5, 260, 77, 40),
    (2, 3, 128, 4, 64),      # IP-Adapter token count
    (1, 1, 1, 1, 64),
    # second-generation kernel (tc_attention2.cu: Sq > 128, d <= 64): ragged query pairs, partial key tiles
    (1, 2, 257, 129, 64),    # 3 query tiles -> 2 pairs (phantom fourth tile), 2 key tiles with 1 valid key in the last
    (2, 3, 300, 255, 64),
    (1, 4, 512, 640, 40),    # zero-filled head columns, 5 key tiles
    (3, 60, 384, 300, 64),   # 360 work items: several per persistent CTA (Q double buffer, O hand-over between items)
    (1, 2, 130, 1, 8),
]


def _sdpa_ref(q, k, v, H, causal=False):
    B_, Sq, C = q.shape
    split = lambda t: t.reshape(t.shape[0], t.shape[1], H, C // H).transpose(1, 2)
    o = F.scaled_dot_product_attention(split(q), split(k), split(v), is_causal=causal)
    return o.transpose(1, 2).reshape(B_, Sq, C)


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("case", SDPA_CASES, ids=str)
def test_sdpa(cuda_device, kernel_mode, dtype, case):
    from refiners_b200 import backend as B

    Bn, H, Sq, Sk, D = case
    q, k, v = _gen((Bn, Sq, H * D), 70), _gen((Bn, Sk, H * D), 71), _gen((Bn, Sk, H * D), 72)
    ref = _sdpa_ref(rounded(q, dtype), rounded(k, dtype), rounded(v, dtype), H)
    dev = lambda t: t.to(cuda_device, dtype)
    with torch.no_grad():
        y = B.sdpa(dev(q), dev(k), dev(v), H)
    # P is rounded to the operand dtype before the PV product in the tensor-core kernel: 4 eps
    assert_close(y, ref, dtype, scale=4.0, what=f"sdpa{case}")


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize(
    "geom",
    [(2, 14, 14, 4, 80), (1, 9, 5, 2, 80), (1, 16, 16, 3, 64), (1, 20, 20, 2, 32),
     # SAM's 14 x 14 windows on tc_attention_win.cu: more (window, head) items than SMs, head dim 72, one single item
     (50, 14, 14, 16, 80), (3, 14, 14, 2, 72), (1, 14, 14, 1, 80),
     # the global blocks (64-wide maps) on the same kernel: the full 64 x 64 map, a short one with two query-tile pairs
     (1, 64, 64, 2, 80), (2, 8, 64, 3, 72)],
    ids=str,
)
def test_sam_attention(cuda_device, kernel_mode, dtype, geom):
    """Decomposed relative-position attention (segment_anything/image_encoder.py:87-143 in the
    reference): logits = q k^T d^-1/2 + rel_h[q, kh] + rel_w[q, kw]."""
    from refiners_b200 import backend as B

    Bw, Hh, Ww, heads, d = geom
    C = heads * d
    qkv = _gen((Bw, Hh, Ww, 3 * C), 90)
    rel_h, rel_w = _gen((2 * Hh - 1, d), 91), _gen((2 * Ww - 1, d), 92)
    qkv_r, rh, rw = rounded(qkv, dtype), rounded(rel_h, dtype), rounded(rel_w, dtype)
    t = qkv_r.reshape(Bw, Hh * Ww, 3, heads, d).permute(2, 0, 3, 1, 4)  # [3, Bw, heads, HW, d]
    q, k, v = t[0], t[1], t[2]
    ih = torch.arange(Hh)[:, None] - torch.arange(Hh)[None, :] + Hh - 1
    iw = torch.arange(Ww)[:, None] - torch.arange(Ww)[None, :] + Ww - 1
    q5 = q.reshape(Bw, heads, Hh, Ww, d)
    bias_h = torch.einsum("bnhwc,hkc->bnhwk", q5, rh[ih])
    bias_w = torch.einsum("bnhwc,wkc->bnhwk", q5, rw[iw])
    logits = (q * d**-0.5) @ k.transpose(-1, -2)
    logits = logits.reshape(Bw, heads, Hh, Ww, Hh, Ww) + bias_h[..., :, None] + bias_w[..., None, :]
    attn = logits.reshape(Bw, heads, Hh * Ww, Hh * Ww).softmax(-1)
    ref = (attn @ v).transpose(1, 2).reshape(Bw, Hh, Ww, C)
    dev = lambda x: x.to(cuda_device, dtype)
    with torch.no_grad():
        y = B.sam_attention(dev(qkv), dev(rel_h), dev(rel_w), heads)
    assert_close(y, ref, dtype, scale=4.0, what=f"sam_attention{geom}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32], ids=str)
def test_sdpa_causal(cuda_device, kernel_mode, dtype):
    from refiners_b200 import backend as B

    q, k, v = _gen((2, 77, 128), 73), _gen((2, 77, 128), 74), _gen((2, 77, 128), 75)
    ref = _sdpa_ref(rounded(q, dtype), rounded(k, dtype), rounded(v, dtype), 2, causal=True)
    dev = lambda t: t.to(cuda_device, dtype)
    with torch.no_grad():
        y = B.sdpa(dev(q), dev(k), dev(v), 2, True)
    assert_close(y, ref, dtype, scale=4.0, what="sdpa causal")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32], ids=str)
@pytest.mark.parametrize("tokens", [4, 16])
def test_sdpa_dual_kv(cuda_device, kernel_mode, dtype, tokens):
    """IP-Adapter: SDPA(q, k_text, v_text) + s * SDPA(q, k_img, v_img), two softmaxes
    (image_prompt.py:237-309 in the reference)."""
    from refiners_b200 import backend as B

    H, D = 5, 64
    q, k, v = _gen((2, 192, H * D), 76), _gen((2, 77, H * D), 77), _gen((2, 77, H * D), 78)
    k2, v2 = _gen((2, tokens, H * D), 79), _gen((2, tokens, H * D), 80)
    r = lambda t: rounded(t, dtype)
    ref = _sdpa_ref(r(q), r(k), r(v), H) + 0.6 * _sdpa_ref(r(q), r(k2), r(v2), H)
    dev = lambda t: t.to(cuda_device, dtype)
    with torch.no_grad():
        y = B.sdpa(dev(q), dev(k), dev(v), H, k2=dev(k2), v2=dev(v2), scale2=0.6)
    assert_close(y, ref, dtype, scale=4.0, what="sdpa dual kv")


def test_sdpa_strided_views(cuda_device, kernel_mode):
    """q/k/v as column slices of one fused projection (no copies made)."""
    from refiners_b200 import backend as B

    dtype = torch.bfloat16
    qkv = _gen((2, 100, 3 * 128), 81)
    q, k, v = rounded(qkv, dtype).chunk(3, dim=-1)
    ref = _sdpa_ref(q, k, v, 2)
    dq, dk, dv = qkv.to(cuda_device, dtype).chunk(3, dim=-1)
    with torch.no_grad():
        y = B.sdpa(dq, dk, dv, 2)
    assert_close(y, ref, dtype, scale=4.0, what="sdpa strided")


def test_requires_no_grad(cuda_device):
    from refiners_b200 import backend as B

    x = torch.randn(4, 64, device=cuda_device, requires_grad=True)
    w = torch.randn(64, 64, device=cuda_device)
    with pytest.raises(B.BackendError):
        B.linear(x, w)


def test_error_reporting(cuda_device):
    from refiners_b200 import backend as B

    x = torch.randn(4, 64, device=cuda_device)
    w = torch.randn(64, 32, device=cuda_device)
    with torch.no_grad(), pytest.raises(B.BackendError):
        B.linear(x, w)


def test_launch_counter_and_determinism(cuda_device):
    from refiners_b200 import backend as B

    x = torch.randn(512, 256, device=cuda_device, dtype=torch.bfloat16)
    w = torch.randn(256, 256, device=cuda_device, dtype=torch.bfloat16)
    before = B.launch_count()
    with torch.no_grad():
        a = B.linear(x, w)
        b = B.linear(x, w)
    assert B.launch_count() - before == 2
    assert torch.equal(a, b)


def test_cuda_graph_capture(cuda_device):
    """Every op is capture-safe (no sync, no foreign allocation): replay reproduces eager."""
    from refiners_b200 import backend as B

    x = torch.randn(256, 320, device=cuda_device, dtype=torch.bfloat16)
    w = torch.randn(640, 320, device=cuda_device, dtype=torch.bfloat16) * 0.05
    g, b = torch.ones(640, device=cuda_device, dtype=torch.bfloat16), torch.zeros(640, device=cuda_device, dtype=torch.bfloat16)
    with torch.no_grad():
        eager = B.layer_norm(B.linear(x, w), g, b, 1e-5)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            out = B.layer_norm(B.linear(x, w), g, b, 1e-5)
        x.copy_(x * 1.0)
        graph.replay()
        torch.cuda.synchronize()
    assert torch.equal(out, eager)


def test_packed_weight_cache_is_not_fooled_by_address_reuse(cuda_device):
    """The allocator hands a freed weight's address to the next tensor of the same shape; the
    packed-weight cache must key on tensor identity, not on the address."""
    from refiners_b200 import backend as B

    x = torch.randn(1, 64, 8, 8, device=cuda_device, dtype=torch.bfloat16)
    outs, refs, ptrs = [], [], []
    for seed in (1, 2, 3):
        w = (_gen((64, 64, 3, 3), seed) * 0.05).to(cuda_device, torch.bfloat16)
        ptrs.append(w.data_ptr())
        with torch.no_grad():
            outs.append(B.conv2d(x, w, None, 1, 1).float().cpu())
        refs.append(F.conv2d(x.float().cpu(), w.float().cpu(), None, padding=1))
        del w
    for o, r in zip(outs, refs):
        assert_close(o, r, torch.bfloat16, what="conv after weight replacement")


# ------------------------------------------------------------------------------ production shapes
# The shapes the benchmark actually runs (SURVEY.md section 8a at UNet batch 16): 64 key tiles of online softmax,
# pair-mode GEMMs with multi-wave tails, 128^2 / 32^2 convolutions.  The reference is the same op in fp32 torch
# evaluated on the GPU (TF32 off) from the rounded operands - the CPU would need minutes for these.
def _fp32_reference_mode():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


BIG_SDPA = [
    # (B, H, Sq, Sk, D)
    (16, 20, 1024, 1024, 64),   # SDXL self-attention, 1280-wide blocks
    (4, 10, 4096, 4096, 64),    # SDXL self-attention, 640-wide blocks (64 key tiles)
    (16, 20, 1024, 77, 64),     # text cross-attention
    (4, 10, 4096, 77, 64),
    (2, 16, 4096, 4096, 80),    # SAM global attention geometry without the bias term
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("case", BIG_SDPA, ids=str)
def test_sdpa_production_shapes(cuda_device, dtype, case):
    from refiners_b200 import backend as B

    _fp32_reference_mode()
    Bn, H, Sq, Sk, D = case
    dev = lambda t: t.to(cuda_device, dtype)
    q, k, v = dev(_gen((Bn, Sq, H * D), 170)), dev(_gen((Bn, Sk, H * D), 171)), dev(_gen((Bn, Sk, H * D), 172))
    with torch.no_grad():
        y = B.sdpa(q, k, v, H)
        ref = torch.cat([_sdpa_ref(q[i : i + 1].float(), k[i : i + 1].float(), v[i : i + 1].float(), H) for i in range(Bn)])
    assert_close(y, ref, dtype, scale=4.0, what=f"sdpa{case}")


BIG_LINEAR = [
    # (M, K, N, residual)
    (16384, 1280, 1280, True),    # attention out-projection + residual: 4.3 waves of 256-wide tiles
    (16384, 1280, 3840, False),   # fused q/k/v
    (16384, 5120, 1280, True),    # GEGLU down-projection
    (65536, 640, 640, True),
    (65536, 640, 1920, False),
    (1232, 2048, 2560, False),    # text K/V projection
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("case", BIG_LINEAR, ids=str)
def test_linear_production_shapes(cuda_device, dtype, case):
    from refiners_b200 import backend as B

    _fp32_reference_mode()
    M, K, N, with_res = case
    dev = lambda t: t.to(cuda_device, dtype)
    x, w, b = dev(_gen((M, K), 180)), dev(_gen((N, K), 181, K**-0.5)), dev(_gen((N,), 182))
    r = dev(_gen((M, N), 183)) if with_res else None
    with torch.no_grad():
        y = B.linear(x, w, b, residual=r)
        ref = F.linear(x.float(), w.float(), b.float())
        if r is not None:
            ref = ref + r.float()
    assert_close(y, ref, dtype, what=f"linear{case}")


@pytest.mark.parametrize("dtype", [torch.bfloat16], ids=str)
def test_linear_geglu_production_shape(cuda_device, dtype):
    from refiners_b200 import backend as B

    _fp32_reference_mode()
    M, K, Fh = 16384, 1280, 5120
    dev = lambda t: t.to(cuda_device, dtype)
    x, w, b = dev(_gen((M, K), 184)), dev(_gen((2 * Fh, K), 185, K**-0.5)), dev(_gen((2 * Fh,), 186))
    with torch.no_grad():
        y = B.linear_geglu(x, w, b)
        a, g = F.linear(x.float(), w.float(), b.float()).chunk(2, dim=-1)
        ref = a * F.gelu(g)
    assert_close(y, ref, dtype, scale=2.0, what="geglu 16384x1280->5120")


BIG_CONV = [
    # (B, Cin, Cout, H, W, k, stride, pad)
    (16, 1280, 1280, 32, 32, 3, 1, 1),
    (16, 320, 320, 128, 128, 3, 1, 1),
    (16, 640, 640, 64, 64, 3, 2, 1),     # Downsample
    (16, 1920, 640, 64, 64, 1, 1, 0),    # shortcut of an up block
    (4, 4, 320, 128, 128, 3, 1, 1),      # latent input conv (zero-extended channels)
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("case", BIG_CONV, ids=str)
def test_conv2d_production_shapes(cuda_device, dtype, case):
    from refiners_b200 import backend as B

    _fp32_reference_mode()
    Bn, Cin, Cout, H, W, k, stride, pad = case
    dev = lambda t: t.to(cuda_device, dtype)
    x = dev(_gen((Bn, Cin, H, W), 190)).contiguous(memory_format=torch.channels_last)
    w, b = dev(_gen((Cout, Cin, k, k), 191, (Cin * k * k) ** -0.5)), dev(_gen((Cout,), 192))
    with torch.no_grad():
        y = B.conv2d(x, w, b, stride, pad)
        ref = F.conv2d(x.float(), w.float(), b.float(), stride=stride, padding=pad)
    assert_close(y, ref, dtype, what=f"conv{case}")


@pytest.mark.parametrize("shape", [(16, 320, 128, 128), (16, 1280, 32, 32), (16, 2560, 32, 32)], ids=str)
def test_group_norm_silu_production_shapes(cuda_device, shape):
    from refiners_b200 import backend as B

    dtype = torch.bfloat16
    dev = lambda t: t.to(cuda_device, dtype)
    x = dev(_gen(shape, 193, 2.0) + 0.5).contiguous(memory_format=torch.channels_last)
    g, b = dev(1 + 0.1 * _gen((shape[1],), 194)), dev(0.1 * _gen((shape[1],), 195))
    with torch.no_grad():
        y = B.group_norm(x, 32, g, b, 1e-5, silu=True)
        ref = F.silu(F.group_norm(x.float(), 32, g.float(), b.float(), 1e-5))
    assert_close(y, ref, dtype, scale=2.0, what=f"gn+silu{shape}")


@pytest.mark.parametrize("shape", [(16, 1024, 1280), (16, 4096, 640)], ids=str)
def test_layer_norm_production_shapes(cuda_device, shape):
    from refiners_b200 import backend as B

    dtype = torch.bfloat16
    dev = lambda t: t.to(cuda_device, dtype)
    x, g, b = dev(_gen(shape, 196, 3.0)), dev(1 + 0.1 * _gen((shape[-1],), 197)), dev(0.1 * _gen((shape[-1],), 198))
    with torch.no_grad():
        y = B.layer_norm(x, g, b, 1e-5)
        ref = F.layer_norm(x.float(), (shape[-1],), g.float(), b.float(), 1e-5)
    assert_close(y, ref, dtype, scale=2.0, what=f"ln{shape}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("shape", [(2, 4, 512, 1024), (1, 2, 300, 700)], ids=str)
def test_sdpa_growing_maxima(cuda_device, dtype, shape):
    """The second-generation attention kernel keeps a STALE running maximum and rescales the TMEM-resident output only
    when a row's maximum grows by more than 2^8: keys whose magnitude ramps up along the sequence make every later key
    tile raise the maximum by a wide margin (the rescale path), while a flat first half exercises the stale path."""
    from refiners_b200 import backend as B

    _fp32_reference_mode()
    Bn, H, Sq, Sk = shape
    D = 64
    ramp = torch.linspace(0.5, 6.0, Sk).reshape(1, Sk, 1)
    dev = lambda t: t.to(cuda_device, dtype)
    q, k, v = dev(_gen((Bn, Sq, H * D), 270) * 2.0), dev(_gen((Bn, Sk, H * D), 271) * ramp), dev(_gen((Bn, Sk, H * D), 272))
    with torch.no_grad():
        y = B.sdpa(q, k, v, H)
        ref = _sdpa_ref(q.float(), k.float(), v.float(), H)
    assert_close(y, ref, dtype, scale=4.0, what=f"sdpa growing maxima {shape}")


SHORT_SDPA = [
    # (B, H, Sq, Sk, D): the single-pass short-key kernel (Sk <= 128, D <= 64) at its chunk boundaries and with ragged tiles
    (1, 1, 1, 1, 8), (2, 3, 127, 16, 64), (2, 3, 129, 17, 40), (1, 2, 300, 32, 64), (1, 2, 256, 33, 64), (3, 2, 130, 64, 16),
    (1, 4, 128, 65, 64), (2, 2, 500, 77, 64), (1, 2, 384, 96, 64), (1, 2, 384, 97, 64), (2, 1, 200, 127, 48), (1, 3, 260, 128, 64),
    (8, 20, 1024, 77, 64),  # many items per CTA: the three-stage ring and both softmax groups wrap around several times
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("case", SHORT_SDPA, ids=str)
def test_sdpa_short_keys(cuda_device, dtype, case):
    """tc_sdpa_short_kernel against fp32 attention on the same (rounded) operands, plus: strided q / k / v views of fused
    projections are read in place, the result is deterministic, and it agrees with the first-generation flash kernel."""
    import os

    from refiners_b200 import backend as B

    _fp32_reference_mode()
    Bn, H, Sq, Sk, D = case
    dev = lambda t: t.to(cuda_device, dtype)
    q, k, v = dev(_gen((Bn, Sq, H * D), 370)), dev(_gen((Bn, Sk, H * D), 371) * 1.5), dev(_gen((Bn, Sk, H * D), 372))
    with torch.no_grad():
        y = B.sdpa(q, k, v, H)
        ref = _sdpa_ref(q.float(), k.float(), v.float(), H)
        assert torch.equal(y, B.sdpa(q, k, v, H))
        kv = torch.cat((k, v), dim=-1)  # the fused K / V projection of a cross-attention
        assert torch.equal(y, B.sdpa(q, kv[..., : H * D], kv[..., H * D :], H))
    assert_close(y, ref, dtype, scale=4.0, what=f"short sdpa{case}")


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("case", [(8, 20, 1024, 77, 4, 64), (2, 10, 4096, 77, 16, 64), (2, 3, 200, 96, 32, 40), (1, 2, 130, 17, 1, 64), (2, 2, 256, 64, 20, 64)], ids=str)
def test_sdpa_short_keys_dual_kv(cuda_device, dtype, case):
    """The IP-Adapter form on the single-pass kernel: both softmaxes share one accumulator (the image tokens' probabilities
    are pre-scaled by scale2 * l / l2); against fp32 on the same operands, and against the first-generation dual kernel."""
    import os

    from refiners_b200 import backend as B

    _fp32_reference_mode()
    Bn, H, Sq, Sk, Sk2, D = case
    dev = lambda t: t.to(cuda_device, dtype)
    q, k, v = dev(_gen((Bn, Sq, H * D), 470)), dev(_gen((Bn, Sk, H * D), 471)), dev(_gen((Bn, Sk, H * D), 472))
    k2, v2 = dev(_gen((Bn, Sk2, H * D), 473) * 2.0), dev(_gen((Bn, Sk2, H * D), 474))
    with torch.no_grad():
        y = B.sdpa(q, k, v, H, k2=k2, v2=v2, scale2=0.6)
        ref = torch.cat([
            _sdpa_ref(q[i : i + 1].float(), k[i : i + 1].float(), v[i : i + 1].float(), H)
            + 0.6 * _sdpa_ref(q[i : i + 1].float(), k2[i : i + 1].float(), v2[i : i + 1].float(), H)
            for i in range(Bn)
        ])
        assert torch.equal(y, B.sdpa(q, k, v, H, k2=k2, v2=v2, scale2=0.6))
    assert_close(y, ref, dtype, scale=4.0, what=f"short dual sdpa{case}")


@pytest.mark.parametrize("dtype", DTYPES, ids=str)
@pytest.mark.parametrize("guided", [True, False], ids=["cfg", "plain"])
def test_cfg_euler_glue_is_bit_identical_to_the_operator_sequence(cuda_device, dtype, guided):
    """rb200_cfg_scale_input / rb200_cfg_euler against the reference's operator-by-operator evaluation
    (model.py:137-159, solvers/euler.py:63-100) on the same device and dtype: the fused launches round every intermediate
    where the ATen sequence does, so the results are the same bits."""
    from refiners_b200 import backend as B
    from refiners_b200.foundationals.latent_diffusion import Euler

    solver = Euler(num_inference_steps=30).to(device=cuda_device, dtype=dtype)
    x = (_gen((3, 4, 40, 24), 300) * 7.0).to(cuda_device, dtype)
    eps = _gen((6 if guided else 3, 4, 40, 24), 301).to(cuda_device, dtype)
    for step, scale in ((0, 5.0), (11, 7.5), (29, 1.3)):
        with torch.no_grad():
            want_in = solver.scale_model_input(torch.cat((x, x)) if guided else x, step=step)
            got_in = B.cfg_scale_input(x, solver.sigmas, step, twice=guided)
            if guided:
                u, c = eps.chunk(2)
                noise = u + scale * (c - u)
            else:
                noise = eps
            want = solver(x, predicted_noise=noise, step=step)
            got = B.cfg_euler(x, eps, solver.sigmas, step, scale, guided)
        assert torch.equal(got_in, want_in), f"scale_model_input differs at step {step}: max {(got_in.float() - want_in.float()).abs().max().item():.3e}"
        assert torch.equal(got, want), f"CFG + Euler update differs at step {step}: max {(got.float() - want.float()).abs().max().item():.3e}"
