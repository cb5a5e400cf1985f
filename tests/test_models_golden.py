"""refiners_b200 models against golden vectors recorded from the real reference
(tests/golden/*.safetensors, see oracle/pin_against_reference.py).

CPU tests exercise the host path (BASELINE config 1: plumbing, fp32, <= 1e-5 relative).
GPU tests run the same graphs through the C-ABI kernels: fp32 within 2e-4 relative (CUDA-core
path, fp32 accumulate, different summation order), bf16 within 5 % of max|ref| max-abs and 1 %
mean-abs - the reference's own bf16-vs-fp32 noise on a UNet forward is 1.4 % max-abs
(SURVEY.md section 6), so bf16 bit-parity with fp32 is not a meaningful bar."""

from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file

import refiners_b200.fluxion.layers as fl
from oracle.weights import keyed_state_dict
from refiners_b200.fluxion.utils import no_grad
from refiners_b200.foundationals.latent_diffusion import (
    CrossAttentionBlock2d,
    Euler,
    RangeAdapter2d,
    ResidualBlock,
    SD1UNet,
    SDXLUNet,
)

GOLDEN = Path(__file__).parent / "golden"


def sub(fx, prefix):
    return {k[len(prefix):]: v for k, v in fx.items() if k.startswith(prefix)}


def check(got, want, dtype):
    got, want = got.float().cpu(), want.float()
    scale = max(want.abs().max().item(), 1e-3)
    err = (got - want).abs()
    if dtype == "host":
        assert err.max().item() <= 1e-5 * scale, f"max abs {err.max().item():.3e}"
    elif dtype == torch.float32:
        assert err.max().item() <= 2e-4 * scale, f"max abs {err.max().item():.3e} (scale {scale:.3f})"
    else:
        assert err.max().item() <= 5e-2 * scale, f"max abs {err.max().item():.3e} (scale {scale:.3f})"
        assert err.mean().item() <= 1e-2 * scale, f"mean abs {err.mean().item():.3e} (scale {scale:.3f})"


@pytest.fixture(scope="module")
def blocks():
    return load_file(str(GOLDEN / "blocks.safetensors"))


def build_residual(blocks, tag):
    cin, cout = (64, 64) if tag == "res_same" else (64, 96)
    rb = ResidualBlock(cin, cout)
    body = rb.layer("Chain", fl.Chain)
    RangeAdapter2d(target=body.layer("Conv2d_1", fl.Conv2d), channels=cout, embedding_dim=32, context_key="timestep_embedding").inject(body)
    top = fl.Chain(rb)
    top.load_state_dict(sub(blocks, f"{tag}.sd."))  # reference keys load unchanged
    return top


def build_xattn(blocks, tag):
    ca = CrossAttentionBlock2d(64, context_embedding_dim=48, context_key="ctx", num_attention_heads=2, num_attention_layers=2,
                               use_bias=False, use_linear_projection=(tag == "xattn_linear"))
    ca.load_state_dict(sub(blocks, f"{tag}.sd."))
    return ca


def run_residual(blocks, tag, device, dtype):
    top = build_residual(blocks, tag).to(device, dtype)
    top.set_context("range_adapter", {"timestep_embedding": blocks[f"{tag}.temb"].to(device, dtype)})
    with no_grad():
        return top(blocks[f"{tag}.x"].to(device, dtype))


def run_xattn(blocks, tag, device, dtype):
    ca = build_xattn(blocks, tag).to(device, dtype)
    ca.set_context("cross_attention_block", {"ctx": blocks[f"{tag}.ctx"].to(device, dtype)})
    with no_grad():
        return ca(blocks[f"{tag}.x"].to(device, dtype))


@pytest.mark.parametrize("tag", ["res_same", "res_proj"])
def test_residual_block_host(blocks, tag):
    check(run_residual(blocks, tag, "cpu", torch.float32), blocks[f"{tag}.y"], "host")


@pytest.mark.parametrize("tag", ["xattn_linear", "xattn_conv"])
def test_cross_attention_2d_host(blocks, tag):
    check(run_xattn(blocks, tag, "cpu", torch.float32), blocks[f"{tag}.y"], "host")


def test_euler_host():
    f = load_file(str(GOLDEN / "euler.safetensors"))
    s = Euler(num_inference_steps=30)
    assert torch.equal(s.sigmas, f["euler.sigmas"]) and torch.equal(s.timesteps, f["euler.timesteps"])
    assert torch.equal(s.scale_model_input(f["euler.x"], -1), f["euler.scaled_init"])
    assert torch.equal(s.scale_model_input(f["euler.x"], 7), f["euler.scaled_7"])
    assert torch.equal(s(f["euler.x"], predicted_noise=f["euler.eps"], step=7), f["euler.step_7"])
    assert torch.equal(s(f["euler.x"], predicted_noise=f["euler.eps"], step=29), f["euler.step_29"])
    assert torch.equal(Euler(num_inference_steps=30).to(dtype=torch.bfloat16).sigmas, f["euler.sigmas_bf16"])


def load_unet(cls, seed, device="cpu", dtype=torch.float32):
    shapes = {k: tuple(v.shape) for k, v in cls(4, device="meta").state_dict().items()}
    sd = keyed_state_dict(shapes, seed=seed)
    unet = cls(4, device="meta")
    unet.load_state_dict({k: v.to(device, dtype) for k, v in sd.items()}, assign=True)
    return unet


def run_sd1(unet, f, device, dtype):
    unet.set_timestep(f["sd1.timestep"].to(device))
    unet.set_clip_text_embedding(f["sd1.ctx"].to(device, dtype))
    with no_grad():
        return unet(f["sd1.x"].to(device, dtype))


def run_sdxl(unet, f, device, dtype):
    unet.set_timestep(f["sdxl.timestep"].to(device))
    unet.set_clip_text_embedding(f["sdxl.ctx"].to(device, dtype))
    unet.set_pooled_text_embedding(f["sdxl.pooled"].to(device, dtype))
    unet.set_time_ids(f["sdxl.time_ids"].to(device))
    with no_grad():
        return unet(f["sdxl.x"].to(device, dtype))


def test_sd1_unet_host():
    """BASELINE config 1: SD1UNet, fp32, CPU - the Chain/Context plumbing end to end."""
    f = load_file(str(GOLDEN / "unets.safetensors"))
    unet = load_unet(SD1UNet, seed=1)
    y1 = run_sd1(unet, f, "cpu", torch.float32)
    check(y1, f["sd1.y"], "host")
    assert torch.equal(y1, run_sd1(unet, f, "cpu", torch.float32))  # context is flushed between calls


def test_sdxl_unet_host():
    f = load_file(str(GOLDEN / "unets.safetensors"))
    unet = load_unet(SDXLUNet, seed=2)
    check(run_sdxl(unet, f, "cpu", torch.float32), f["sdxl.y"], "host")


def load_control_lora_unet(device="cpu", dtype=torch.float32):
    """SDXLUNet + ControlLoraAdapter('canny') + the synthetic LoRA checkpoint, keyed weights (seed 3) -
    the construction recorded by oracle/pin_against_reference.py."""
    from oracle.pin_against_reference import control_lora_test_weights
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import ControlLoraAdapter

    def build():
        unet = SDXLUNet(4, device="meta")
        adapter = ControlLoraAdapter("canny", unet, scale=0.8).inject()
        ControlLoraAdapter.load_lora_layers("canny", control_lora_test_weights(), adapter.control_lora)
        return unet, adapter

    unet, adapter = build()
    shapes = {k: tuple(v.shape) for k, v in unet.state_dict().items()}
    sd = keyed_state_dict(shapes, seed=3)
    # shared leaves appear under two keys (UNet path and ControlLora path); like load_state_dict on real
    # tensors, the LAST key in state-dict order provides the value of the shared storage
    final: dict[int, torch.Tensor] = {}
    params = dict(unet.state_dict(keep_vars=True))
    for k in shapes:
        final[id(params[k])] = sd[k]
    unet.load_state_dict({k: final[id(params[k])].to(device, dtype) for k in shapes}, assign=True)
    return unet, adapter


def run_control_lora(unet, adapter, f, device, dtype):
    unet.set_timestep(f["cl.timestep"].to(device))
    unet.set_clip_text_embedding(f["cl.ctx"].to(device, dtype))
    unet.set_pooled_text_embedding(f["cl.pooled"].to(device, dtype))
    unet.set_time_ids(f["cl.time_ids"].to(device))
    adapter.set_condition(f["cl.cond"].to(device, dtype))
    with no_grad():
        return unet(f["cl.x"].to(device, dtype))


def test_sdxl_control_lora_host():
    """BASELINE config 4 graph (SDXL + ControlLora): shared leaves, LoRAs inside the control copy,
    condition encoder, zero-convs accumulating into the UNet's residual slots."""
    f = load_file(str(GOLDEN / "unets.safetensors"))
    unet, adapter = load_control_lora_unet()
    check(run_control_lora(unet, adapter, f, "cpu", torch.float32), f["cl.y"], "host")


def test_sd1_denoising_step_host():
    """LatentDiffusionModel.forward (model.py:128-159 in the reference): contexts, CFG doubling, sigma
    scaling, UNet, CFG combine, Euler update - one whole step of StableDiffusion_1 against the
    reference's recorded outputs at the first, a middle and the last step."""
    from refiners_b200.foundationals.latent_diffusion import StableDiffusion_1

    f = load_file(str(GOLDEN / "step.safetensors"))
    sd = StableDiffusion_1(unet=load_unet(SD1UNet, seed=1), solver=Euler(num_inference_steps=30))
    for step, scale in ((0, 7.5), (7, 5.0), (29, 7.5)):
        with no_grad():
            y = sd(f["step.x"], step=step, clip_text_embedding=f["step.ctx"], condition_scale=scale)
        check(y, f[f"step.y_{step}"], "host")


def test_vae_host():
    """LatentDiffusionAutoencoder: encode / decode on keyed weights against the reference,
    and the PIL round trip of the helpers."""
    from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder

    f = load_file(str(GOLDEN / "vae.safetensors"))
    lda = LatentDiffusionAutoencoder(device="meta")
    sd = keyed_state_dict({k: tuple(v.shape) for k, v in lda.state_dict().items()}, seed=5)
    lda.load_state_dict(sd, assign=True)
    with no_grad():
        check(lda.decode(f["vae.z"]), f["vae.decoded"], "host")
        check(lda.encode(f["vae.image"]), f["vae.encoded"], "host")
        images = lda.latents_to_images(f["vae.z"])
    assert len(images) == 2 and images[0].size == (128, 96) and images[0].mode == "RGB"
    want = ((f["vae.decoded"][:1] + 1) / 2).clamp(0, 1)
    from refiners_b200.fluxion.utils import image_to_tensor

    assert (image_to_tensor(images[0]) - want).abs().max().item() <= 1 / 255 + 1e-6


def tiled_vae(device, dtype):
    """The reference's recorded tiled round trip (224x160 image, 128x96 tiles, 32-pixel blend): (latents, decoded,
    tree restored)."""
    from PIL import Image

    from refiners_b200.foundationals.latent_diffusion.auto_encoder import FixedGroupNorm, LatentDiffusionAutoencoder

    f = load_file(str(GOLDEN / "vae_tiled.safetensors"))
    lda = LatentDiffusionAutoencoder(device="meta")
    sd = keyed_state_dict({k: tuple(v.shape) for k, v in lda.state_dict().items()}, seed=5)
    lda.load_state_dict({k: v.to(device, dtype) for k, v in sd.items()}, assign=True)
    image = Image.fromarray(f["tiled.pixels"].numpy())
    before = repr(lda)
    with pytest.raises(ValueError):
        lda.tiled_image_to_latents(image)
    with no_grad(), lda.tiled_inference(image, tile_size=(128, 96), blending=32):
        assert len([*lda.layers(FixedGroupNorm)]) == 52 and all(n.mean is not None for n in lda.layers(FixedGroupNorm))
        latents = lda.tiled_image_to_latents(image)
        decoded = lda._tiled_decode(f["tiled.latents"].to(device, dtype), lda._tile_size, 32)
        picture = lda.tiled_latents_to_image(f["tiled.latents"].to(device, dtype))
    assert picture.size == (224, 160) and repr(lda) == before
    return f, latents, decoded


def test_vae_tiled_host():
    """Tiled inference with frozen GroupNorm statistics against the reference's recorded latents and pixels, and the
    oracle's restatement against the same."""
    from oracle import vae as ovae

    f, latents, decoded = tiled_vae("cpu", torch.float32)
    check(latents, f["tiled.latents"], "host")
    check(decoded, f["tiled.decoded"], "host")
    from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder

    sd = keyed_state_dict({k: tuple(v.shape) for k, v in LatentDiffusionAutoencoder(device="meta").state_dict().items()}, seed=5)
    full = f["tiled.pixels"].permute(2, 0, 1)[None].float() / 255
    small = f["tiled.small"].float() / 255
    ovae.frozen_stats = {}
    try:
        with no_grad():
            ovae.capture_statistics(sd, full, small)
            check(ovae.tiled(sd, ovae.decode, f["tiled.latents"], (20, 28), (96, 128), 32, 1, 8, 3), f["tiled.decoded"], "host")
    finally:
        ovae.frozen_stats = None


def test_dinov2_host():
    """DINOv2 ViT on keyed weights against the reference: the published small model and a tiny
    registers + SwiGLU configuration whose input grid (6 x 5) differs from the positional grid (4 x 4)."""
    from refiners_b200.foundationals.dinov2 import DINOv2_small, ViT

    f = load_file(str(GOLDEN / "dinov2.safetensors"))

    def keyed(model, seed):
        sd = keyed_state_dict({k: tuple(v.shape) for k, v in model.state_dict().items()}, seed=seed)
        model.load_state_dict(sd, assign=True)
        return model

    with no_grad():
        check(keyed(DINOv2_small(device="meta"), 6)(f["small.x"]), f["small.y"], "host")
        tiny = ViT(embedding_dim=64, patch_size=4, image_size=16, num_layers=2, num_heads=2, num_registers=3, feedforward_dim=96,
                   interpolate_antialias=True, activation=fl.GLU(fl.SiLU()), device="meta")
        check(keyed(tiny, 7)(f["tiny.x"]), f["tiny.y"], "host")


def load_controlnet_unet(device="cpu", dtype=torch.float32):
    """SD1UNet + SD1ControlnetAdapter('canny', scale 0.9, decay 0.825) with keyed weights (seed 4) - the
    construction recorded by oracle/pin_against_reference.py::pin_controlnet."""
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1 import SD1ControlnetAdapter

    def build():
        unet = SD1UNet(4, device="meta")
        return unet, SD1ControlnetAdapter(unet, name="canny", scale=0.9, scale_decay=0.825).inject()

    unet, adapter = build()
    shapes = {k: tuple(v.shape) for k, v in unet.state_dict().items()}
    sd = keyed_state_dict(shapes, seed=4)
    unet.load_state_dict({k: v.to(device, dtype) for k, v in sd.items()}, assign=True)
    return unet, adapter


def run_controlnet(unet, adapter, f, device, dtype):
    unet.set_timestep(f["cn.timestep"].to(device))
    unet.set_clip_text_embedding(f["cn.ctx"].to(device, dtype))
    if adapter is not None:
        adapter.set_controlnet_condition(f["cn.cond"].to(device, dtype))
    with no_grad():
        return unet(f["cn.x"].to(device, dtype))


def test_sd1_controlnet_host():
    """SD 1.5 ControlNet (A16): 13 scaled zero-conv taps into the UNet's residual slots, then eject."""
    f = load_file(str(GOLDEN / "controlnet.safetensors"))
    unet, adapter = load_controlnet_unet()
    before = len(unet.state_dict())
    check(run_controlnet(unet, adapter, f, "cpu", torch.float32), f["cn.y"], "host")
    adapter.scale = 0.0  # taps muted: the plain UNet's output
    check(run_controlnet(unet, adapter, f, "cpu", torch.float32), f["cn.y_plain"], "host")
    adapter.eject()
    assert len(unet.state_dict()) == before - 340 and unet.parent is None
    check(run_controlnet(unet, None, f, "cpu", torch.float32), f["cn.y_plain"], "host")


# ------------------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
def test_vae_gpu(cuda_device, dtype):
    """LatentDiffusionAutoencoder.encode / decode on the kernels (the bottleneck attention has ONE head of dim 512)."""
    from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder

    f = load_file(str(GOLDEN / "vae.safetensors"))
    lda = LatentDiffusionAutoencoder(device="meta")
    sd = keyed_state_dict({k: tuple(v.shape) for k, v in lda.state_dict().items()}, seed=5)
    lda.load_state_dict({k: v.to(cuda_device, dtype) for k, v in sd.items()}, assign=True)
    with no_grad():
        check(lda.decode(f["vae.z"].to(cuda_device, dtype)), f["vae.decoded"], dtype)
        check(lda.encode(f["vae.image"].to(cuda_device, dtype)), f["vae.encoded"], dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
def test_vae_tiled_gpu(cuda_device, dtype):
    """Same round trip with every tile on the kernels: GroupNorms run rb200_group_norm_fixed (statistics captured on the
    resized image, frozen for the tiles)."""
    f, latents, decoded = tiled_vae(cuda_device, dtype)
    check(latents, f["tiled.latents"], dtype)
    check(decoded, f["tiled.decoded"], dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=str)
def test_group_norm_fixed_kernel(cuda_device, dtype):
    """rb200_group_norm_fixed: the capturing pass equals GroupNorm and returns (mean, rstd); a frozen pass applies those
    statistics to a different tensor (checked against the formula in fp32)."""
    from refiners_b200 import backend as B

    gen = torch.Generator().manual_seed(5)
    x = (torch.randn(2, 64, 9, 7, generator=gen) * 2 + 0.5).to(cuda_device, dtype)
    other = torch.randn(2, 64, 5, 11, generator=gen).to(cuda_device, dtype)
    gamma, beta = torch.randn(64, generator=gen).to(cuda_device, dtype), torch.randn(64, generator=gen).to(cuda_device, dtype)
    y, stats = B.group_norm_fixed(x, 8, gamma, beta, 1e-5)
    assert torch.equal(y, B.group_norm(x, 8, gamma, beta, 1e-5))
    g = x.float().reshape(2, 8, -1)
    assert torch.allclose(stats[..., 0], g.mean(2), atol=1e-5) and torch.allclose(stats[..., 1], (g.var(2, correction=0) + 1e-5).rsqrt(), rtol=1e-4)
    kept = stats.clone()
    z, again = B.group_norm_fixed(other, 8, gamma, beta, 1e-5, stats)
    assert again is stats and torch.equal(stats, kept)
    o = other.float().reshape(2, 8, -1)
    want = ((o - stats[..., :1]) * stats[..., 1:]).reshape(other.shape) * gamma.float().view(1, -1, 1, 1) + beta.float().view(1, -1, 1, 1)
    tol = {torch.float32: 1e-5, torch.bfloat16: 2**-7, torch.float16: 2**-10}[dtype]
    assert (z.float() - want).abs().max().item() <= tol * want.abs().max().item()
    with pytest.raises(B.BackendError):
        B.group_norm_fixed(other[:1], 8, gamma, beta, 1e-5, stats)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
def test_sd1_controlnet_gpu(cuda_device, dtype):
    f = load_file(str(GOLDEN / "controlnet.safetensors"))
    unet, adapter = load_controlnet_unet(device=cuda_device, dtype=dtype)
    check(run_controlnet(unet, adapter, f, cuda_device, dtype), f["cn.y"], dtype)
    adapter.eject()
    check(run_controlnet(unet, None, f, cuda_device, dtype), f["cn.y_plain"], dtype)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=str)
@pytest.mark.parametrize("fusion", [True, False], ids=["fused", "unfused"])
def test_blocks_gpu(cuda_device, blocks, dtype, fusion):
    from refiners_b200 import backend as B

    prev = B.set_fusion(fusion)
    try:
        before = B.launch_count()
        for tag in ("res_same", "res_proj"):
            check(run_residual(blocks, tag, cuda_device, dtype), blocks[f"{tag}.y"], dtype)
        for tag in ("xattn_linear", "xattn_conv"):
            check(run_xattn(blocks, tag, cuda_device, dtype), blocks[f"{tag}.y"], dtype)
        assert B.launch_count() > before
    finally:
        B.set_fusion(prev)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
def test_sd1_unet_gpu(cuda_device, dtype):
    f = load_file(str(GOLDEN / "unets.safetensors"))
    unet = load_unet(SD1UNet, seed=1, device=cuda_device, dtype=dtype)
    y1 = run_sd1(unet, f, cuda_device, dtype)
    check(y1, f["sd1.y"], dtype)
    assert torch.equal(y1, run_sd1(unet, f, cuda_device, dtype)), "two identical forwards must be bit-identical"


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
def test_sdxl_unet_gpu(cuda_device, dtype):
    f = load_file(str(GOLDEN / "unets.safetensors"))
    unet = load_unet(SDXLUNet, seed=2, device=cuda_device, dtype=dtype)
    y1 = run_sdxl(unet, f, cuda_device, dtype)
    check(y1, f["sdxl.y"], dtype)
    assert torch.equal(y1, run_sdxl(unet, f, cuda_device, dtype))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
def test_sdxl_control_lora_gpu(cuda_device, dtype):
    f = load_file(str(GOLDEN / "unets.safetensors"))
    unet, adapter = load_control_lora_unet(device=cuda_device, dtype=dtype)
    check(run_control_lora(unet, adapter, f, cuda_device, dtype), f["cl.y"], dtype)
