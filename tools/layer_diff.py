"""Per-leaf comparison of a model between the host path (CPU fp32) and the kernels (CUDA).

    python tools/layer_diff.py sd1 float32        # or: sdxl bfloat16

Forward hooks record every leaf's output on both devices (hooks disable fusion across the hooked
leaves, so each leaf's true output is seen) and the first divergences are printed in call order.
"""

from __future__ import annotations

import sys
from pathlib import Path

import torch
from safetensors.torch import load_file

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import refiners_b200.fluxion.layers as fl  # noqa: E402
from refiners_b200.fluxion.utils import no_grad  # noqa: E402
from tests.test_models_golden import load_unet, run_sd1, run_sdxl  # noqa: E402
from refiners_b200.foundationals.latent_diffusion import SD1UNet, SDXLUNet  # noqa: E402


def record(model: fl.Chain) -> tuple[list[tuple[str, torch.Tensor]], list]:
    log: list[tuple[str, torch.Tensor]] = []
    handles = []
    for name, module in model.named_modules():
        if isinstance(module, fl.Chain) or not isinstance(module, fl.Module):
            continue
        if isinstance(module, (fl.UseContext, fl.SetContext, fl.Identity)):
            continue

        def hook(mod, args, out, name=name):
            if isinstance(out, torch.Tensor):
                log.append((f"{name} [{type(mod).__name__}]", out.detach().float().cpu()))

        handles.append(module.register_forward_hook(hook))
    return log, handles


def main() -> None:
    which = sys.argv[1] if len(sys.argv) > 1 else "sd1"
    dtype = getattr(torch, sys.argv[2]) if len(sys.argv) > 2 else torch.float32
    f = load_file(str(ROOT / "tests" / "golden" / "unets.safetensors"))
    cls, seed, run = (SD1UNet, 1, run_sd1) if which == "sd1" else (SDXLUNet, 2, run_sdxl)
    cpu = load_unet(cls, seed)
    log_c, _ = record(cpu)
    with no_grad():
        y_c = run(cpu, f, "cpu", torch.float32)
    del cpu
    gpu = load_unet(cls, seed, device="cuda", dtype=dtype)
    log_g, _ = record(gpu)
    with no_grad():
        y_g = run(gpu, f, torch.device("cuda"), dtype)
    print("final max abs diff", (y_g.float().cpu() - y_c).abs().max().item(), "ref max", y_c.abs().max().item())
    tol = 1e-3 if dtype == torch.float32 else 5e-2
    shown = 0
    assert len(log_c) == len(log_g), (len(log_c), len(log_g))
    for (name, a), (_, b) in zip(log_c, log_g):
        scale = max(a.abs().max().item(), 1e-3)
        err = (a - b).abs().max().item() / scale
        if err > tol:
            print(f"{err:9.3e}  shape {tuple(a.shape)}  {name}")
            shown += 1
            if shown >= 12:
                break
    if not shown:
        print("no leaf diverges beyond", tol)


if __name__ == "__main__":
    main()
