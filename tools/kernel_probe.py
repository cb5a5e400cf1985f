"""Launch one hot kernel a few times in isolation (for `ncu --set full -k regex:<name>`).

    python tools/kernel_probe.py gemm        # [16384,1280] x [1280,1280]^T  (dominant SDXL Linear)
    python tools/kernel_probe.py conv        # 3x3 1280->1280 @ 32x32, batch 16
    python tools/kernel_probe.py attn        # self-attention B=16 H=20 S=1024 d=64
    python tools/kernel_probe.py attn4096    # self-attention B=16 H=10 S=4096 d=64
    python tools/kernel_probe.py gemm_res | gemm640_res | attn77 | attn77_4096 | attn_sam_win | attn_sam_global
Also prints CUDA-event timings (L2 flushed between launches)."""

import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from refiners_b200 import backend as B  # noqa: E402

dev = torch.device("cuda")
which = sys.argv[1] if len(sys.argv) > 1 else "gemm"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
bf = torch.bfloat16
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)

if which == "gemm":
    M, K, N = 16384, 1280, 1280
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    fn, flops = (lambda: B.linear(x, w)), 2.0 * M * N * K
elif which == "gemm_ff":
    M, K, N = 16384, 1280, 10240
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    fn, flops = (lambda: B.linear(x, w)), 2.0 * M * N * K
elif which == "gemm_geglu":  # Linear 1280 -> 10240 + GLU(GeLU): the largest single item of the SDXL step
    M, K, N = 16384, 1280, 10240
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    b = torch.randn(N, device=dev, dtype=bf)
    fn, flops = (lambda: B.linear_geglu(x, w, b)), 2.0 * M * N * K
elif which in ("gemm_mlp", "gemm_mlp_gelu"):  # SAM ViT-H MLP up-projection at batch 4: [16384,1280] x [1280,5120]^T + bias (+ GeLU)
    M, K, N = 16384, 1280, 5120
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    b = torch.randn(N, device=dev, dtype=bf)
    epi = B.EPI_GELU if which == "gemm_mlp_gelu" else B.EPI_NONE
    fn, flops = (lambda: B.linear(x, w, b, epilogue=epi)), 2.0 * M * N * K
elif which == "gemm640":
    M, K, N = 65536, 640, 640
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    fn, flops = (lambda: B.linear(x, w)), 2.0 * M * N * K
elif which == "conv":
    x = torch.randn(16, 1280, 32, 32, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
    w = torch.randn(1280, 1280, 3, 3, device=dev, dtype=bf) * 0.01
    fn, flops = (lambda: B.conv2d(x, w, None, 1, 1)), 2.0 * 16 * 1024 * 1280 * 1280 * 9
elif which == "conv320":
    x = torch.randn(16, 320, 128, 128, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
    w = torch.randn(320, 320, 3, 3, device=dev, dtype=bf) * 0.02
    fn, flops = (lambda: B.conv2d(x, w, None, 1, 1)), 2.0 * 16 * 16384 * 320 * 320 * 9
elif which == "gemm_kv":  # cross-attention K/V projection of the 77 text tokens, UNet batch 16
    M, K, N = 1232, 2048, 2560
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    fn, flops = (lambda: B.linear(x, w)), 2.0 * M * N * K
elif which == "gemm_res":
    M, K, N = 16384, 1280, 1280
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    b, r = torch.randn(N, device=dev, dtype=bf), torch.randn(M, N, device=dev, dtype=bf)
    fn, flops = (lambda: B.linear(x, w, b, residual=r)), 2.0 * M * N * K
elif which == "gemm640_res":
    M, K, N = 65536, 640, 640
    x, w = torch.randn(M, K, device=dev, dtype=bf), torch.randn(N, K, device=dev, dtype=bf) * 0.03
    b, r = torch.randn(N, device=dev, dtype=bf), torch.randn(M, N, device=dev, dtype=bf)
    fn, flops = (lambda: B.linear(x, w, b, residual=r)), 2.0 * M * N * K
elif which in ("attn77", "attn77_4096"):
    Bn, H, S = (16, 20, 1024) if which == "attn77" else (16, 10, 4096)
    q = torch.randn(Bn, S, H * 64, device=dev, dtype=bf)
    k, v = torch.randn(Bn, 77, H * 64, device=dev, dtype=bf), torch.randn(Bn, 77, H * 64, device=dev, dtype=bf)
    fn, flops = (lambda: B.sdpa(q, k, v, H)), 4.0 * Bn * H * S * 77 * 64
elif which == "attn77_dual":  # IP-Adapter: text cross-attention + 4 image tokens in one launch
    Bn, H, S = 16, 20, 1024
    q = torch.randn(Bn, S, H * 64, device=dev, dtype=bf)
    k, v = torch.randn(Bn, 77, H * 64, device=dev, dtype=bf), torch.randn(Bn, 77, H * 64, device=dev, dtype=bf)
    k2, v2 = torch.randn(Bn, 4, H * 64, device=dev, dtype=bf), torch.randn(Bn, 4, H * 64, device=dev, dtype=bf)
    fn, flops = (lambda: B.sdpa(q, k, v, H, k2=k2, v2=v2, scale2=0.6)), 4.0 * Bn * H * S * 81 * 64
elif which in ("attn_sam_win", "attn_sam_win4"):  # the windowed blocks of SAM ViT-H at batch 1 / batch 4 (rel-pos tables + attention)
    nwin = 25 if which == "attn_sam_win" else 100
    qkv = torch.randn(nwin, 14, 14, 3 * 1280, device=dev, dtype=bf)
    rh, rw = torch.randn(27, 80, device=dev, dtype=bf), torch.randn(27, 80, device=dev, dtype=bf)
    fn, flops = (lambda: B.sam_attention(qkv, rh, rw, 16)), 4.0 * nwin * 16 * 196 * 196 * 80
elif which in ("attn_sam_global", "attn_sam_global4"):
    qkv = torch.randn(1 if which == "attn_sam_global" else 4, 64, 64, 3 * 1280, device=dev, dtype=bf)
    rh, rw = torch.randn(127, 80, device=dev, dtype=bf), torch.randn(127, 80, device=dev, dtype=bf)
    fn, flops = (lambda: B.sam_attention(qkv, rh, rw, 16)), 4.0 * qkv.shape[0] * 16 * 4096 * 4096 * 80
elif which in ("attn", "attn4096"):
    Bn, H, S = (16, 20, 1024) if which == "attn" else (16, 10, 4096)
    q = torch.randn(Bn, S, H * 64, device=dev, dtype=bf)
    k, v = torch.randn_like(q), torch.randn_like(q)
    fn, flops = (lambda: B.sdpa(q, k, v, H)), 4.0 * Bn * H * S * S * 64
elif which == "gn":  # GroupNorm + SiLU on the largest UNet map: 16 x 320 x 128 x 128
    x = torch.randn(16, 320, 128, 128, device=dev, dtype=bf).contiguous(memory_format=torch.channels_last)
    g, b = torch.ones(320, device=dev, dtype=bf), torch.zeros(320, device=dev, dtype=bf)
    fn, flops = (lambda: B.group_norm(x, 32, g, b, 1e-5, silu=True)), 0.0
    nbytes = x.numel() * 2 * 3  # read twice (statistics, apply) + write once
elif which == "ln":  # LayerNorm of the 1280-wide transformer blocks: [16, 1024, 1280]
    x = torch.randn(16, 1024, 1280, device=dev, dtype=bf)
    g, b = torch.ones(1280, device=dev, dtype=bf), torch.zeros(1280, device=dev, dtype=bf)
    fn, flops = (lambda: B.layer_norm(x, g, b, 1e-5)), 0.0
    nbytes = x.numel() * 2 * 2
else:
    raise SystemExit(f"unknown probe {which}")

with torch.no_grad():
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
ms = sorted(times)[len(times) // 2]
if flops:
    print(f"{which}: median {ms * 1e3:.1f} us, {flops / ms / 1e9:.1f} TFLOP/s ({flops / 1e9:.1f} GFLOP per launch)")
else:
    print(f"{which}: median {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s ({nbytes / 1e6:.1f} MB algorithmic per launch)")
