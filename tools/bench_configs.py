"""BASELINE configs 3 and 4 on one GPU, same step and timing rules as bench.py:

    python tools/bench_configs.py --config 3   # SDXL + 2 rank-16 LoRAs on every Linear under CrossAttentionBlock
                                               #   (700 adapters) + SDXL IP-Adapter (4 image tokens), latent batch 8
    python tools/bench_configs.py --config 4   # SDXL + ControlLora('canny') with rank-R LoRAs on the control copy's
                                               #   Linear layers, non-zero zero-convs, latent batch 4 (B = 8)
Random-init weights (LoRA `up` re-drawn ~ N(0, 0.02) so the LoRA work is not a multiply by zero), synthetic inputs,
bf16, CUDA-graph replay of the UNet, CUDA-event timing over --steps steps after 3 warm-up steps."""

import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import refiners_b200.fluxion.layers as fl  # noqa: E402
from refiners_b200 import backend as B  # noqa: E402
from refiners_b200.fluxion.adapters import LinearLora, LoraAdapter  # noqa: E402
from refiners_b200.fluxion.utils import manual_seed, no_grad  # noqa: E402
from refiners_b200.foundationals.latent_diffusion import Euler, SDXLUNet, StableDiffusion_XL  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--config", type=int, choices=[2, 3, 4], default=3)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--rank", type=int, default=64, help="config 4: LoRA rank inside the control copy")
ap.add_argument("--no-graph", action="store_true")
args = ap.parse_args()

dev, dtype = torch.device("cuda"), torch.bfloat16
manual_seed(0)
unet = SDXLUNet(in_channels=4, device=dev, dtype=dtype)
lb = 4 if args.config == 4 else 8
extra = {}

if args.config == 3:
    from refiners_b200.foundationals.latent_diffusion.image_prompt import SDXLIPAdapter

    targets = [
        (lin, parent)
        for lin, parent in unet.walk(fl.Linear, recurse=True)
        if "CrossAttentionBlock" in {type(a).__name__ for a in parent.get_parents() + [parent]}
    ]
    for lin, parent in targets:
        loras = []
        for j, scale in enumerate((1.0, 1.4)):
            lora = LinearLora(f"lora{j}", in_features=lin.in_features, out_features=lin.out_features, rank=16, scale=scale,
                              device=dev, dtype=dtype)
            lora.up.weight.data.normal_(0, 0.02)
            loras.append(lora)
        LoraAdapter(lin, *loras).inject(parent)
    ip = SDXLIPAdapter(unet, scale=0.6)
    ip.inject()
    ip.set_clip_image_embedding(torch.randn(2 * lb, 4, 2048, device=dev, dtype=dtype))
    extra = {"lora_adapters": len(targets), "ip_sub_adapters": len(ip.sub_adapters)}
elif args.config == 4:
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import ControlLoraAdapter, ZeroConvolution

    adapter = ControlLoraAdapter("canny", unet, scale=1.0).inject()
    cl = adapter.control_lora
    sd = {}
    g = torch.Generator().manual_seed(7)
    for lin, parent in cl.walk(fl.Linear, recurse=True):
        path = lin.get_path(parent=parent, top=cl).split(".", 1)[1]
        sd[f"ControlLora.{path}.down"] = torch.randn(args.rank, lin.in_features, generator=g) / args.rank
        sd[f"ControlLora.{path}.up"] = torch.randn(lin.out_features, args.rank, generator=g) * 0.02
    if sd:
        ControlLoraAdapter.load_lora_layers("canny", sd, cl)
    for zc in cl.layers(ZeroConvolution):
        for prm in zc.parameters():
            prm.data.normal_(0, 0.02)
    adapter.set_condition(torch.rand(2 * lb, 3, 1024, 1024, device=dev, dtype=dtype))
    extra = {"control_lora_loras": len(sd) // 2, "rank": args.rank}

sdxl = StableDiffusion_XL(unet=unet, solver=Euler(num_inference_steps=30), device=dev, dtype=dtype)
if not args.no_graph:
    sdxl.enable_cuda_graph()
g = torch.Generator().manual_seed(1000)
x = (torch.randn(lb, 4, 128, 128, generator=g) * float(sdxl.solver.init_noise_sigma)).to(dev, dtype)
clip = torch.randn(2 * lb, 77, 2048, generator=g).to(dev, dtype)
pooled = torch.randn(2 * lb, 1280, generator=g).to(dev, dtype)
ids = torch.tensor([[1024, 1024, 0, 0, 1024, 1024]]).repeat(2 * lb, 1).to(dev)


def step(s: int) -> torch.Tensor:
    return sdxl(x, step=s % 30, clip_text_embedding=clip, pooled_text_embedding=pooled, time_ids=ids)


with no_grad():
    for s in range(3):
        y = step(s)
    torch.cuda.synchronize()
    n0 = B.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in range(args.steps):
        y = step(s)
    e1.record()
    torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / args.steps
runner = sdxl._graphed_unet[0] if sdxl._graphed_unet else None
print(json.dumps({
    "config": args.config, "metric": "SDXL 1024^2 bf16 denoising steps/s", "value": 1000.0 / ms, "unit": "steps/s",
    "ms_per_step": ms, "latent_batch": lb, "unet_batch": 2 * lb, "latents_steps_per_s": lb * 1000.0 / ms,
    "graph": runner is not None, "launches_per_step": runner.launches_per_replay if runner else (B.launch_count() - n0) // args.steps,
    "finite": bool(torch.isfinite(y.float()).all()), **extra,
}))
