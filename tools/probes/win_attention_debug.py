"""SAM window attention (tc_attention_win.cu) against fp32 torch, error split by the part of the kernel that produced it:
output columns 0..63 come from the 128-byte-swizzle V slab, 64..79 from the 32-byte-swizzle slab; zeroing q[..., 64:]
removes the 32-byte-swizzle contribution to the logits.  usage: python tools/probes/win_attention_debug.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from refiners_b200 import backend as B

torch.manual_seed(0)
dev = torch.device("cuda:0")


def ref(qkv, rel_h, rel_w, heads):
    Bw, Hh, Ww, C3 = qkv.shape
    d = C3 // 3 // heads
    t = qkv.float().reshape(Bw, Hh * Ww, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    ih = torch.arange(Hh, device=dev)[:, None] - torch.arange(Hh, device=dev)[None, :] + Hh - 1
    iw = torch.arange(Ww, device=dev)[:, None] - torch.arange(Ww, device=dev)[None, :] + Ww - 1
    q5 = q.reshape(Bw, heads, Hh, Ww, d)
    bh = torch.einsum("bnhwc,hkc->bnhwk", q5, rel_h.float()[ih])
    bw = torch.einsum("bnhwc,wkc->bnhwk", q5, rel_w.float()[iw])
    lg = (q * d**-0.5) @ k.transpose(-1, -2)
    lg = lg.reshape(Bw, heads, Hh, Ww, Hh, Ww) + bh[..., :, None] + bw[..., None, :]
    a = lg.reshape(Bw, heads, Hh * Ww, Hh * Ww).softmax(-1)
    return (a @ v).transpose(1, 2).reshape(Bw, Hh, Ww, heads * d)


def run(tag, Bw, heads, d, zero_q_tail=False, zero_bias=False, Hh=14, Ww=14):
    qkv = torch.randn(Bw, Hh, Ww, 3 * heads * d, device=dev).to(torch.bfloat16)
    if zero_q_tail:
        v = qkv.view(Bw, Hh, Ww, 3, heads, d)
        v[:, :, :, 0, :, 64:] = 0
    rh = (torch.randn(2 * Hh - 1, d, device=dev) * (0 if zero_bias else 1)).to(torch.bfloat16)
    rw = (torch.randn(2 * Ww - 1, d, device=dev) * (0 if zero_bias else 1)).to(torch.bfloat16)
    y = B.sam_attention(qkv, rh, rw, heads).float()
    r = ref(qkv, rh, rw, heads)
    e = (y - r).abs().reshape(Bw, Hh * Ww, heads, d)
    print(f"{tag:34s} max|ref| {r.abs().max():.3f}  err cols[0:64] {e[..., :64].max():.4f}  cols[64:] {e[..., 64:].max():.4f}  "
          f"rows[0:128] {e[:, :128].max():.4f}  rows[128:] {e[:, 128:].max():.4f}  nan {int(torch.isnan(y).sum())}", flush=True)


print("RB200_ATTN_WIN =", os.environ.get("RB200_ATTN_WIN", "(unset: 1)"))
run("full", 2, 4, 80)
run("no bias", 2, 4, 80, zero_bias=True)
run("q[64:]=0", 2, 4, 80, zero_q_tail=True)
run("q[64:]=0, no bias", 2, 4, 80, zero_q_tail=True, zero_bias=True)
run("d=72", 3, 2, 72)
run("800 windows-heads", 50, 16, 80)
run("global 64x64", 1, 2, 80, Hh=64, Ww=64)
run("global 64x64 no bias", 1, 2, 80, zero_bias=True, Hh=64, Ww=64)
run("global 8x64 d=72 (2 pairs)", 2, 3, 72, Hh=8, Ww=64)
run("global 64x64 16 heads", 1, 16, 80, Hh=64, Ww=64)
