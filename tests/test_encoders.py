"""DINOv2 and CLIP vision towers against reference-recorded fixtures
(tests/golden/dinov2.safetensors, clip.safetensors; oracle/pin_against_reference.py --only-dinov2 / --only-clip):
host path (fp32, <= 1e-5 relative) and the same graphs on the CUDA kernels - fp32 within 2e-4 relative, bf16 under the
criterion of tests/test_full_size_gpu.py (error vs the reference's fp32 output no larger than that of torch-eager bf16 on
the same GPU + 1e-3 max|ref|)."""

from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file

import refiners_b200.fluxion.layers as fl
from oracle import clip as oclip
from oracle import dinov2 as odino
from oracle import ops as oops
from oracle.cases import keyed_input
from oracle.weights import keyed_state_dict
from refiners_b200.fluxion.utils import no_grad

GOLDEN = Path(__file__).parent / "golden"


def keyed(model, seed, device="cpu", dtype=torch.float32):
    sd = keyed_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=seed)
    model.load_state_dict({k: v.to(device, dtype) for k, v in sd.items()}, assign=True)
    return model


def rel_err(got, want):
    got, want = got.float().cpu(), want.float().cpu()
    scale = max(want.abs().max().item(), 1e-3)
    d = got - want
    return d.abs().max().item() / scale, d.pow(2).mean().sqrt().item() / scale


def build(case: str, device="cpu", dtype=torch.float32):
    """(model, input, oracle evaluation on a state dict, fixture key)"""
    if case == "clip_h":
        from refiners_b200.foundationals.clip import CLIPImageEncoderH

        return (keyed(CLIPImageEncoderH(device="meta"), 9, device, dtype), keyed_input("clip.h.image", (2, 3, 224, 224)),
                lambda sd, x: oclip.image_encoder(sd, x, patch_size=14, num_layers=32, num_heads=16), ("clip", "h.y"))
    if case == "clip_tiny":
        from refiners_b200.foundationals.clip import CLIPImageEncoder

        m = CLIPImageEncoder(image_size=32, embedding_dim=32, output_dim=16, patch_size=8, num_layers=2, num_attention_heads=2,
                             feedforward_dim=64, device="meta")
        return (keyed(m, 10, device, dtype), keyed_input("clip.tiny.image", (3, 3, 32, 32)),
                lambda sd, x: oclip.image_encoder(sd, x, patch_size=8, num_layers=2, num_heads=2), ("clip", "tiny.y"))
    if case == "dinov2_small":
        from refiners_b200.foundationals.dinov2 import DINOv2_small

        return (keyed(DINOv2_small(device="meta"), 6, device, dtype), None,
                lambda sd, x: odino.vit(sd, x, patch_size=14, num_layers=12, num_heads=6), ("dinov2", "small"))
    if case == "dinov2_tiny":
        from refiners_b200.foundationals.dinov2 import ViT

        m = ViT(embedding_dim=64, patch_size=4, image_size=16, num_layers=2, num_heads=2, num_registers=3, feedforward_dim=96,
                interpolate_antialias=True, activation=fl.GLU(fl.SiLU()), device="meta")
        return (keyed(m, 7, device, dtype), None,
                lambda sd, x: odino.vit(sd, x, patch_size=4, num_layers=2, num_heads=2, num_registers=3, swiglu=True, interpolate_antialias=True),
                ("dinov2", "tiny"))
    raise KeyError(case)


def fixture(case_key, x):
    file, key = case_key
    f = load_file(str(GOLDEN / f"{file}.safetensors"))
    if file == "dinov2":
        return f[f"{key}.x"], f[f"{key}.y"]
    return x, f[key]


@pytest.mark.parametrize("case", ["clip_tiny", "clip_h"])
def test_clip_image_encoder_host(case):
    model, x, _, key = build(case)
    x, want = fixture(key, x)
    with no_grad():
        e_max, _ = rel_err(model(x), want)
    assert e_max <= 2e-5, f"{case}: {e_max:.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
@pytest.mark.parametrize("case", ["clip_tiny", "clip_h", "dinov2_tiny", "dinov2_small"])
def test_vision_tower_gpu(cuda_device, case, dtype):
    from refiners_b200 import backend as B

    model, x, oracle_eval, key = build(case, cuda_device, dtype)
    x, want = fixture(key, x)
    xd = x.to(cuda_device, dtype)
    before = B.launch_count()
    with no_grad():
        y = model(xd)
        assert torch.equal(y, model(xd)), "two identical forwards must be bit-identical"
    assert B.launch_count() > before
    e_max, e_rms = rel_err(y, want)
    if dtype == torch.float32:
        print(f"\n[{case} fp32] max-abs {e_max:.3e} of max|ref|")
        assert e_max <= 2e-4
        return
    prev, oops.FAST = oops.FAST, True
    try:
        with torch.no_grad():
            eager = oracle_eval(dict(model.state_dict()), xd)
    finally:
        oops.FAST = prev
    t_max, t_rms = rel_err(eager, want)
    print(f"\n[{case} bf16] engine max-abs {e_max:.3e} rms {e_rms:.3e} | torch-eager bf16 max-abs {t_max:.3e} rms {t_rms:.3e} (relative to max|ref|)")
    assert e_rms <= t_rms + 1e-3 and e_max <= 1.25 * t_max + 1e-3
