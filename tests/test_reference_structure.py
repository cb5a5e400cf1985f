"""Structure parity against the REAL reference, when it is mounted (this container; skipped on the GPU box).

Every model of the hot path is built twice on the meta device - once from /root/reference, once from
refiners_b200 - and the two must agree on the full `repr()` tree (class names, tags, argument echo: what
the reference's own structure tests assert on, e.g. tests/adapters/test_ip_adapter.py:29-41) and on the
state-dict contract (same keys in the same order, same shapes), before and after adapter injection."""

import re
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference/src/refiners")
pytestmark = pytest.mark.skipif(not REF.exists(), reason="/root/reference is not mounted here")


@pytest.fixture(scope="module")
def ref():
    from oracle.pin_against_reference import _import_reference

    _import_reference()
    import refiners  # noqa: F401

    return refiners


def contract(module):
    return [(k, tuple(v.shape)) for k, v in module.state_dict().items()]


def tree(module) -> str:
    """repr() with the signature echo of Lambda layers reduced to the function name: the reference annotates its
    helper functions with jaxtyping shapes, which is typing style, not structure."""
    return re.sub(r"Lambda\((\w+)\(.*$", r"Lambda(\1)", repr(module), flags=re.MULTILINE)


def same(mine, theirs):
    assert contract(mine) == contract(theirs)
    assert tree(mine) == tree(theirs)


def test_sd1_unet_and_controlnet(ref):
    from refiners.foundationals.latent_diffusion.stable_diffusion_1.controlnet import SD1ControlnetAdapter as RAdapter
    from refiners.foundationals.latent_diffusion.stable_diffusion_1.unet import SD1UNet as RUNet

    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1 import SD1ControlnetAdapter, SD1UNet

    mine, theirs = SD1UNet(4, device="meta"), RUNet(4, device="meta")
    same(mine, theirs)
    a, b = SD1ControlnetAdapter(mine, name="canny", scale=0.9).inject(), RAdapter(theirs, name="canny", scale=0.9).inject()
    same(mine, theirs)
    assert tree(a) == tree(b)
    a.eject(), b.eject()
    same(mine, theirs)


def test_sdxl_unet_control_lora_and_ip_adapter(ref):
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import ControlLoraAdapter as RControl
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.image_prompt import SDXLIPAdapter as RIP
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet as RUNet

    from refiners_b200.foundationals.latent_diffusion import SDXLUNet
    from refiners_b200.foundationals.latent_diffusion.image_prompt import SDXLIPAdapter
    from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.control_lora import ControlLoraAdapter

    mine, theirs = SDXLUNet(4, device="meta"), RUNet(4, device="meta")
    same(mine, theirs)
    a, b = ControlLoraAdapter("canny", mine, scale=0.8).inject(), RControl("canny", theirs, scale=0.8).inject()
    same(mine, theirs)
    a.eject(), b.eject()
    same(mine, theirs)
    ia, ib = SDXLIPAdapter(mine, scale=0.5), RIP(theirs, scale=0.5)
    ia.inject(), ib.inject()
    same(mine, theirs)
    assert contract(ia.image_proj) == contract(ib.image_proj)
    ia.eject(), ib.eject()
    same(mine, theirs)


def test_ip_adapter_plus_perceiver_resampler(ref):
    """fine_grained=True: the 16-token PerceiverResampler image projection, structure and numbers (same weights)."""
    from refiners.foundationals.latent_diffusion.image_prompt import PerceiverResampler as RResampler
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.image_prompt import SDXLIPAdapter as RIP
    from refiners.foundationals.latent_diffusion.stable_diffusion_xl.unet import SDXLUNet as RUNet

    from refiners_b200.foundationals.latent_diffusion import SDXLUNet
    from refiners_b200.foundationals.latent_diffusion.image_prompt import SDXLIPAdapter
    from refiners_b200.foundationals.latent_diffusion.perceiver import PerceiverResampler

    ia, ib = SDXLIPAdapter(SDXLUNet(4, device="meta"), fine_grained=True), RIP(RUNet(4, device="meta"), fine_grained=True)
    assert isinstance(ia.image_proj, PerceiverResampler)
    same(ia.image_proj, ib.image_proj)
    cfg = dict(latents_dim=64, num_attention_layers=2, num_attention_heads=4, head_dim=16, num_tokens=5, input_dim=48, output_dim=40)
    torch.manual_seed(0)
    mine, theirs = PerceiverResampler(**cfg), RResampler(**cfg)
    mine.load_state_dict(theirs.state_dict())
    x = torch.randn(3, 11, 48)
    with torch.no_grad():
        assert torch.equal(mine(x), theirs(x))


def test_clip_image_encoder_and_image_embedding(ref):
    """CLIP vision tower (structure of H; numbers on a tiny tower with the reference's weights) and the IP-Adapter's
    once-per-prompt path image -> context tensor, including weights and token concatenation."""
    import refiners.fluxion.layers as rfl
    from refiners.foundationals.clip.image_encoder import CLIPImageEncoder as REnc, CLIPImageEncoderH as REncH
    from refiners.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock2d as RBlock
    from refiners.foundationals.latent_diffusion.image_prompt import ImageProjection as RProj, IPAdapter as RIP

    import refiners_b200.fluxion.layers as fl
    from refiners_b200.foundationals.clip import CLIPImageEncoder, CLIPImageEncoderH
    from refiners_b200.foundationals.latent_diffusion import CrossAttentionBlock2d
    from refiners_b200.foundationals.latent_diffusion.image_prompt import ImageProjection, IPAdapter

    mine_h, theirs_h = CLIPImageEncoderH(device="meta"), REncH(device="meta")
    same(mine_h, theirs_h)
    same(IPAdapter.convert_to_grid_features(mine_h), RIP.convert_to_grid_features(theirs_h))

    cfg = dict(image_size=32, embedding_dim=48, output_dim=24, patch_size=8, num_layers=2, num_attention_heads=3, feedforward_dim=96)
    block = dict(channels=64, context_embedding_dim=40, context_key="ctx", num_attention_heads=2, use_linear_projection=True)
    torch.manual_seed(0)
    enc_r, proj_r, tgt_r = REnc(**cfg), RProj(clip_image_embedding_dim=24, clip_text_embedding_dim=40), rfl.Chain(RBlock(**block))
    enc_m, proj_m, tgt_m = CLIPImageEncoder(**cfg), ImageProjection(clip_image_embedding_dim=24, clip_text_embedding_dim=40), fl.Chain(CrossAttentionBlock2d(**block))
    enc_m.load_state_dict(enc_r.state_dict()), proj_m.load_state_dict(proj_r.state_dict())
    ip_r, ip_m = RIP(tgt_r, enc_r, proj_r), IPAdapter(tgt_m, enc_m, proj_m)
    images = torch.randn(3, 3, 32, 32)
    with torch.no_grad():
        assert torch.equal(enc_m(images), enc_r(images))
        for kwargs in ({}, {"weights": [1.0, 0.5, 2.0]}, {"concat_batches": False}):
            assert torch.equal(ip_m.compute_clip_image_embedding(images, **kwargs), ip_r.compute_clip_image_embedding(images, **kwargs))


def test_lora_adapters_on_cross_attention(ref):
    import refiners.fluxion.layers as rfl
    from refiners.fluxion.adapters.lora import LinearLora as RLora, LoraAdapter as RAdapter
    from refiners.foundationals.latent_diffusion.cross_attention import CrossAttentionBlock2d as RBlock

    import refiners_b200.fluxion.layers as fl
    from refiners_b200.fluxion.adapters import LinearLora, LoraAdapter
    from refiners_b200.foundationals.latent_diffusion import CrossAttentionBlock2d

    kw = dict(channels=64, context_embedding_dim=48, context_key="ctx", num_attention_heads=2, num_attention_layers=2,
              use_linear_projection=True, device="meta")
    mine, theirs = fl.Chain(CrossAttentionBlock2d(**kw)), rfl.Chain(RBlock(**kw))
    same(mine, theirs)
    for chain, linear, lora_cls, adapter_cls in ((mine, fl.Linear, LinearLora, LoraAdapter), (theirs, rfl.Linear, RLora, RAdapter)):
        for lin, parent in list(chain.walk(linear, recurse=True)):
            loras = [lora_cls(f"l{j}", in_features=lin.in_features, out_features=lin.out_features, rank=4, scale=s, device="meta")
                     for j, s in enumerate((1.0, 1.4))]
            adapter_cls(lin, *loras).inject(parent)
    same(mine, theirs)


def test_sam_vit_h(ref):
    from refiners.foundationals.segment_anything.image_encoder import SAMViTH as RViT

    from refiners_b200.foundationals.segment_anything import SAMViTH

    same(SAMViTH(device="meta"), RViT(device="meta"))


def test_vae(ref):
    from refiners.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder as RVAE

    from refiners_b200.foundationals.latent_diffusion.auto_encoder import LatentDiffusionAutoencoder

    same(LatentDiffusionAutoencoder(device="meta"), RVAE(device="meta"))


@pytest.mark.parametrize("name", ["DINOv2_small", "DINOv2_base_reg", "DINOv2_large", "DINOv2_giant_reg"])
def test_dinov2(ref, name):
    import refiners.foundationals.dinov2 as theirs

    import refiners_b200.foundationals.dinov2 as mine

    same(getattr(mine, name)(device="meta"), getattr(theirs, name)(device="meta"))


def test_solver_tables(ref):
    from refiners.foundationals.latent_diffusion.solvers import DDIM as RDDIM, Euler as REuler

    from refiners_b200.foundationals.latent_diffusion import DDIM, Euler

    for mine, theirs in ((Euler(num_inference_steps=30), REuler(num_inference_steps=30)), (DDIM(num_inference_steps=20), RDDIM(num_inference_steps=20))):
        assert torch.equal(mine.timesteps, theirs.timesteps)
        assert torch.equal(mine.cumulative_scale_factors, theirs.cumulative_scale_factors)
        assert torch.equal(mine.noise_std, theirs.noise_std)
