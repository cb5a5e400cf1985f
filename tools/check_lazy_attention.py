"""Parity + timing of the experimental conditional-rescaling attention (RB200_ATTN_LAZY=1) against the default kernel.

    RB200_ATTN_LAZY=1 python tools/check_lazy_attention.py
The variant was written after round 1's GPU budget was spent: it compiles for sm_100a but has NOT yet run on a GPU, which
is why it is off by default and has no entry in the -m gpu test suite yet.  First thing to do in round 2."""

import os
import sys
from pathlib import Path

import torch
import torch.nn.functional as F

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from refiners_b200 import backend as B  # noqa: E402

dev = torch.device("cuda")
print("RB200_ATTN_LAZY =", os.environ.get("RB200_ATTN_LAZY", "0"))
torch.manual_seed(0)
for (Bn, H, Sq, Sk, scale) in [(2, 4, 256, 256, 1.0), (1, 10, 1024, 1024, 1.0), (2, 5, 200, 333, 1.0), (1, 2, 128, 4096, 8.0), (16, 20, 1024, 1024, 1.0)]:
    q = (torch.randn(Bn, Sq, H * 64, device=dev) * scale).bfloat16()
    k = (torch.randn(Bn, Sk, H * 64, device=dev) * scale).bfloat16()  # scale 8: maxima that keep growing -> rescales
    v = torch.randn(Bn, Sk, H * 64, device=dev).bfloat16()
    split = lambda t: t.float().reshape(t.shape[0], t.shape[1], H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(split(q), split(k), split(v)).transpose(1, 2).reshape(Bn, Sq, H * 64)
    with torch.no_grad():
        y = B.sdpa(q, k, v, H)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            B.sdpa(q, k, v, H)
        e1.record()
        torch.cuda.synchronize()
    err = (y.float() - ref).abs().max().item()
    us = e0.elapsed_time(e1) * 100
    print(f"B={Bn} H={H} Sq={Sq} Sk={Sk} scale={scale}: max abs err {err:.3e} (ref max {ref.abs().max().item():.2f}), "
          f"{us:.1f} us, {4.0 * Bn * H * Sq * Sk * 64 / us / 1e6:.1f} TFLOP/s")
