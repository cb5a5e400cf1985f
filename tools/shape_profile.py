"""Per-shape time breakdown of one eager SDXL UNet forward (UNet batch 16, bf16): every backend op
is timed with CUDA events (synchronising after each call - the absolute total is therefore slower
than a graph replay; compare shares) and aggregated by (op, problem shape)."""

import collections
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from refiners_b200 import backend as B  # noqa: E402
from refiners_b200.fluxion.utils import manual_seed, no_grad  # noqa: E402
from refiners_b200.foundationals.latent_diffusion import SDXLUNet  # noqa: E402

dev = torch.device("cuda")
stats: dict = collections.defaultdict(lambda: [0, 0.0, 0.0])
enabled = False


def wrap(name, fn, describe):
    def timed(*args):
        if not enabled:
            return fn(*args)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*args)
        e1.record()
        torch.cuda.synchronize()
        key, flops = describe(*args)
        s = stats[(name, key)]
        s[0] += 1
        s[1] += e0.elapsed_time(e1)
        s[2] += flops
        return out

    return timed


def d_linear(x, w, bias, residual, ld, lu, ls, epi):
    M = x.numel() // x.shape[-1]
    N, K = w.shape
    extra = ("+lora" if ld is not None else "") + ("+res" if residual is not None else "") + (f"+epi{epi}" if epi else "")
    return f"M={M} N={N} K={K}{extra}", 2.0 * M * N * K


def d_conv(x, wp, bias, cb, res, R, S, stride, pad, epi):
    Bn, C, H, W = x.shape
    Co = wp.shape[1]
    Ho, Wo = (H + 2 * pad - R) // stride + 1, (W + 2 * pad - S) // stride + 1
    extra = ("+cb" if cb is not None else "") + ("+res" if res is not None else "")
    return f"B={Bn} {C}->{Co} {R}x{S}s{stride} @{H}x{W}{extra}", 2.0 * Bn * Ho * Wo * Co * C * R * S


def d_sdpa(q, k, v, k2, v2, heads, causal, s2):
    Bn, Sq, C = q.shape
    Sk = k.shape[1]
    return f"B={Bn} H={heads} Sq={Sq} Sk={Sk} d={C // heads}", 4.0 * Bn * Sq * Sk * C


def d_elem(x, *rest):
    return f"{tuple(x.shape)}", 0.0


ops = B._ops
patched = {
    "linear": wrap("linear", ops.linear, d_linear),
    "conv2d": wrap("conv2d", ops.conv2d, d_conv),
    "sdpa": wrap("sdpa", ops.sdpa, d_sdpa),
    "group_norm": wrap("group_norm", ops.group_norm, d_elem),
    "layer_norm": wrap("layer_norm", ops.layer_norm, d_elem),
    "add": wrap("add", ops.add, d_elem),
    "unary": wrap("unary", ops.unary, d_elem),
    "geglu": wrap("geglu", ops.geglu, d_elem),
}


class Proxy:
    def __getattr__(self, name):
        return patched.get(name) or getattr(ops, name)


B._ops = Proxy()

manual_seed(0)
unet = SDXLUNet(4, device=dev, dtype=torch.bfloat16)
x = torch.randn(16, 4, 128, 128, device=dev, dtype=torch.bfloat16)


def run():
    unet.set_timestep(torch.tensor([981.0], device=dev))
    unet.set_clip_text_embedding(torch.randn(16, 77, 2048, device=dev, dtype=torch.bfloat16))
    unet.set_pooled_text_embedding(torch.randn(16, 1280, device=dev, dtype=torch.bfloat16))
    unet.set_time_ids(torch.tensor([[1024, 1024, 0, 0, 1024, 1024]], device=dev).repeat(16, 1))
    return unet(x)


with no_grad():
    run()
    run()
    enabled = True
    run()
total = sum(v[1] for v in stats.values())
print(f"total timed {total:.2f} ms over {sum(v[0] for v in stats.values())} calls")
print(f"{'ms':>9} {'share':>6} {'n':>4} {'TFLOP/s':>8}  op / shape")
for (name, key), (n, ms, fl) in sorted(stats.items(), key=lambda kv: -kv[1][1])[:45]:
    rate = f"{fl / ms / 1e9:8.0f}" if fl else "       -"
    print(f"{ms:9.3f} {100 * ms / total:5.1f}% {n:4d} {rate}  {name} {key}")
