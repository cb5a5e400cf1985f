"""The oracle restatement against the committed golden vectors (generated from the real
reference by oracle/pin_against_reference.py).  CPU only, fp32, tolerance 1e-5 * max|ref|."""

from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file

from oracle import euler as oeuler
from oracle import ops, sam as osam, unet as ounet
from oracle.weights import keyed_state_dict

GOLDEN = Path(__file__).parent / "golden"


def close(got, want, rel=1e-5):
    err = (got - want).abs().max().item()
    assert err <= rel * max(want.abs().max().item(), 1e-3), f"max abs diff {err:.3e}"


def sub(fx, prefix):
    return {k[len(prefix):]: v for k, v in fx.items() if k.startswith(prefix)}


@pytest.fixture(scope="module")
def fx_ops():
    return load_file(str(GOLDEN / "ops.safetensors"))


def test_leaf_ops(fx_ops):
    f = fx_ops
    close(ops.linear(f["linear.x"], f["linear.w"], f["linear.b"]), f["linear.y"])
    for tag, (s, p) in {"3x3": (1, 1), "3x3s2": (2, 1), "1x1": (1, 0)}.items():
        close(ops.conv2d(f[f"conv{tag}.x"], f[f"conv{tag}.w"], f[f"conv{tag}.b"], s, p), f[f"conv{tag}.y"])
    close(ops.group_norm(f["gn.x"], 32, f["gn.w"], f["gn.b"], 1e-6), f["gn.y"])
    close(ops.silu(ops.group_norm(f["gn.x"], 32, f["gn.w"], f["gn.b"], 1e-6)), f["gn_silu.y"])
    close(ops.layer_norm(f["ln.x"], f["ln.w"], f["ln.b"], 1e-5), f["ln.y"])
    close(ops.layer_norm_2d(f["ln2d.x"], f["ln2d.w"], f["ln2d.b"], 1e-6), f["ln2d.y"])
    close(ops.silu(f["act.x"]), f["silu.y"])
    close(ops.gelu(f["act.x"]), f["gelu.y"])
    close(ops.glu_gelu(f["act.x"]), f["glu.y"])
    close(ops.sdpa(f["sdpa.q"], f["sdpa.k"], f["sdpa.v"], 4), f["sdpa.y"])
    q = f["sdpa_causal.q"]
    close(ops.sdpa(q, q, q, 2, True), f["sdpa_causal.y"])
    loras = [(f["lora.down1"], f["lora.up1"], 1.0), (f["lora.down2"], f["lora.up2"], 1.4)]
    close(ops.lora_linear(f["lora.x"], f["lora.w"], f["lora.b"], loras), f["lora.y"])


def test_blocks():
    f = load_file(str(GOLDEN / "blocks.safetensors"))
    for tag in ("res_same", "res_proj"):
        close(ounet.residual_block(sub(f, f"{tag}.sd."), "ResidualBlock", f[f"{tag}.x"], f[f"{tag}.temb"]), f[f"{tag}.y"])
    for tag, lin in (("xattn_linear", True), ("xattn_conv", False)):
        sd = {"X." + k: v for k, v in sub(f, f"{tag}.sd.").items()}
        close(ounet.cross_attention_2d(sd, "X", f[f"{tag}.x"], f[f"{tag}.ctx"], 2, 2, lin), f[f"{tag}.y"])


def test_euler_schedule():
    f = load_file(str(GOLDEN / "euler.safetensors"))
    s = oeuler.EulerSchedule(30)
    close(s.sigmas, f["euler.sigmas"])
    close(s.timesteps, f["euler.timesteps"])
    close(s.scale_model_input(f["euler.x"], -1), f["euler.scaled_init"])
    close(s.scale_model_input(f["euler.x"], 7), f["euler.scaled_7"])
    close(s.update(f["euler.x"], f["euler.eps"], 7), f["euler.step_7"])
    close(s.update(f["euler.x"], f["euler.eps"], 29), f["euler.step_29"])
    assert torch.equal(oeuler.EulerSchedule(30, torch.bfloat16).sigmas, f["euler.sigmas_bf16"])


def test_sam_blocks():
    f = load_file(str(GOLDEN / "sam.safetensors"))
    sd = {"A." + k: v for k, v in sub(f, "fsa.sd.").items()}
    close(osam.fused_self_attention(sd, "A", f["fsa.x"], 2), f["fsa.y"])
    for tag, window in (("layer_win", 4), ("layer_global", None)):
        sd = {"L." + k: v for k, v in sub(f, f"{tag}.sd.").items()}
        close(osam.transformer_layer(sd, "L", f[f"{tag}.x"], 2, window), f[f"{tag}.y"])
    close(osam.patch_encoder({"P." + k: v for k, v in sub(f, "patch.sd.").items()}, "P", f["patch.x"]), f["patch.y"])
    close(osam.neck({"N." + k: v for k, v in sub(f, "neck.sd.").items()}, "N", f["neck.x"]), f["neck.y"])


def test_sd1_unet_full():
    """Whole SD1UNet (859.5 M keyed synthetic weights) against the reference's recorded output."""
    from refiners_b200.foundationals.latent_diffusion import SD1UNet

    f = load_file(str(GOLDEN / "unets.safetensors"))
    shapes = {k: tuple(v.shape) for k, v in SD1UNet(4, device="meta").state_dict().items()}
    sd = keyed_state_dict(shapes, seed=1)
    close(ounet.sd1_unet(sd, f["sd1.x"], f["sd1.timestep"], f["sd1.ctx"]), f["sd1.y"], rel=2e-5)


def test_sd1_controlnet_full():
    """Oracle restatement of SD 1.5 + ControlNet (13 scaled taps into the residual slots) against the
    reference's recorded output, and the plain UNet after eject."""
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1 import SD1ControlnetAdapter, SD1UNet

    f = load_file(str(GOLDEN / "controlnet.safetensors"))
    unet = SD1UNet(4, device="meta")
    SD1ControlnetAdapter(unet, name="canny").inject()
    sd = keyed_state_dict({k: tuple(v.shape) for k, v in unet.state_dict().items()}, seed=4)
    deltas = ounet.sd1_controlnet(sd, f["cn.x"], f["cn.timestep"], f["cn.ctx"], f["cn.cond"], scale=0.9, scale_decay=0.825)
    assert len(deltas) == 13
    close(ounet.sd1_unet(sd, f["cn.x"], f["cn.timestep"], f["cn.ctx"], residuals=deltas), f["cn.y"], rel=2e-5)
    close(ounet.sd1_unet(sd, f["cn.x"], f["cn.timestep"], f["cn.ctx"]), f["cn.y_plain"], rel=2e-5)


def test_denoise_step_full():
    """oracle.euler.denoise_step over the oracle SD1UNet against the reference's StableDiffusion_1 step."""
    from oracle import euler as oeuler
    from refiners_b200.foundationals.latent_diffusion import SD1UNet

    f = load_file(str(GOLDEN / "step.safetensors"))
    shapes = {k: tuple(v.shape) for k, v in SD1UNet(4, device="meta").state_dict().items()}
    sd = keyed_state_dict(shapes, seed=1)
    schedule = oeuler.EulerSchedule(30)
    unet = lambda lat, ts: ounet.sd1_unet(sd, lat, ts, f["step.ctx"])
    close(oeuler.denoise_step(unet, schedule, f["step.x"], 7, 5.0), f["step.y_7"], rel=2e-5)


def test_vae_full():
    """Oracle VAE (encode 72x56 image, decode 12x16 latents) against the reference's recorded outputs."""
    from oracle import vae as ovae
    from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder

    f = load_file(str(GOLDEN / "vae.safetensors"))
    shapes = {k: tuple(v.shape) for k, v in LatentDiffusionAutoencoder(device="meta").state_dict().items()}
    sd = keyed_state_dict(shapes, seed=5)
    close(ovae.decode(sd, f["vae.z"]), f["vae.decoded"], rel=2e-5)
    close(ovae.encode(sd, f["vae.image"]), f["vae.encoded"], rel=2e-5)


def test_dinov2_full():
    """Oracle ViT against the reference: DINOv2_small at 224x224 and a tiny registers + SwiGLU model on 24x20."""
    from oracle import dinov2 as odino
    from refiners_b200.foundationals.dinov2 import DINOv2_small

    f = load_file(str(GOLDEN / "dinov2.safetensors"))
    sd = keyed_state_dict({k: tuple(v.shape) for k, v in DINOv2_small(device="meta").state_dict().items()}, seed=6)
    close(odino.vit(sd, f["small.x"], patch_size=14, num_layers=12, num_heads=6), f["small.y"], rel=2e-5)


def test_fast_mode_matches_golden():
    """oracle.ops.FAST (the fused ATen CPU calls the reference itself makes; used only for the timed
    CPU baseline) is pinned to the same golden vectors."""
    f = load_file(str(GOLDEN / "blocks.safetensors"))
    ops.FAST = True
    try:
        for tag in ("res_same", "res_proj"):
            close(ounet.residual_block(sub(f, f"{tag}.sd."), "ResidualBlock", f[f"{tag}.x"], f[f"{tag}.temb"]), f[f"{tag}.y"])
        for tag, lin in (("xattn_linear", True), ("xattn_conv", False)):
            sd = {"X." + k: v for k, v in sub(f, f"{tag}.sd.").items()}
            close(ounet.cross_attention_2d(sd, "X", f[f"{tag}.x"], f[f"{tag}.ctx"], 2, 2, lin), f[f"{tag}.y"])
    finally:
        ops.FAST = False
