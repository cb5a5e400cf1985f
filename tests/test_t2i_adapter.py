"""T2I-Adapter: condition encoders and the adapters on SD 1.5 / SDXL UNets against the
reference-recorded fixture tests/golden/t2i.safetensors (oracle/pin_against_reference.py --only-t2i)."""

from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file

from oracle import t2i as ot2i
from oracle.cases import keyed_input
from oracle.weights import keyed_state_dict
from refiners_b200.fluxion.utils import no_grad
from refiners_b200.foundationals.latent_diffusion import SD1UNet, SDXLUNet
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_1.t2i_adapter import SD1T2IAdapter
from refiners_b200.foundationals.latent_diffusion.stable_diffusion_xl.t2i_adapter import SDXLT2IAdapter

GOLDEN = Path(__file__).parent / "golden"
CASES = {"sd1": (SD1UNet, SD1T2IAdapter, 1), "sdxl": (SDXLUNet, SDXLT2IAdapter, 2)}


def keyed(module, seed, device, dtype):
    sd = keyed_state_dict({k: tuple(v.shape) for k, v in module.state_dict().items()}, seed=seed)
    module.load_state_dict({k: v.to(device, dtype) for k, v in sd.items()}, assign=True)
    return sd


def run(tag, device, dtype):
    """(features, UNet output with the adapter at scale 0.8, UNet output after scale -> 0 and after eject, oracle features)"""
    unet_cls, adapter_cls, seed = CASES[tag]
    unet = unet_cls(4, device="meta")
    keyed(unet, seed, device, dtype)
    adapter = adapter_cls(unet, name="depth", scale=0.8)
    esd = keyed(adapter.condition_encoder, 31, device, dtype)
    adapter.inject()
    condition = keyed_input(f"t2i.{tag}.condition", (1, 3, 256, 256))
    x = keyed_input(f"t2i.{tag}.x", (1, 4, 32, 32)).to(device, dtype)

    def forward():
        unet.set_timestep(torch.tensor([601.0], device=device))
        if tag == "sd1":
            unet.set_clip_text_embedding(keyed_input("t2i.sd1.ctx", (1, 77, 768)).to(device, dtype))
        else:
            unet.set_clip_text_embedding(keyed_input("t2i.sdxl.ctx", (1, 77, 2048)).to(device, dtype))
            unet.set_pooled_text_embedding(keyed_input("t2i.sdxl.pooled", (1, 1280)).to(device, dtype))
            unet.set_time_ids(torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]], device=device))
        return unet(x)

    with no_grad():
        features = adapter.compute_condition_features(condition.to(device, dtype))
        adapter.set_condition_features(features)
        y = forward()
        adapter.scale = 0.0
        adapter.set_condition_features(features)
        y_off = forward()
        adapter.eject()
        y_plain = forward()
        want = ot2i.condition_encoder(esd, condition, xl=(tag == "sdxl"))
    return features, y, y_off, y_plain, want


def close(got, want, tol):
    got, want = got.float().cpu(), want.float()
    err, scale = (got - want).abs().max().item(), max(want.abs().max().item(), 1e-3)
    assert err <= tol * scale, f"max abs {err:.3e} > {tol:g} x {scale:.3f}"


@pytest.mark.parametrize("tag", ["sd1", "sdxl"])
def test_t2i_adapter_host(tag):
    f = load_file(str(GOLDEN / "t2i.safetensors"))
    features, y, y_off, y_plain, want = run(tag, "cpu", torch.float32)
    close(features[3], f[f"{tag}.feature_3"], 1e-5)
    for got, ref in zip(features, want):
        close(got, ref, 2e-5)
    close(y, f[f"{tag}.y"], 1e-5)
    assert torch.equal(y_off, y_plain) and not torch.equal(y, y_plain)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16], ids=str)
@pytest.mark.parametrize("tag", ["sd1", "sdxl"])
def test_t2i_adapter_gpu(cuda_device, tag, dtype):
    """Condition encoder (convs, ReLU, rb200_avg_pool2d) and the adapted UNet (one add launch per feature map) on the kernels."""
    f = load_file(str(GOLDEN / "t2i.safetensors"))
    features, y, y_off, y_plain, want = run(tag, cuda_device, dtype)
    tol = 2e-4 if dtype == torch.float32 else 3e-2
    for got, ref in zip(features, want):
        close(got, ref, tol)
    close(y, f[f"{tag}.y"], 2e-4 if dtype == torch.float32 else 5e-2)
    assert torch.equal(y_off, y_plain)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16], ids=str)
def test_avg_pool_kernel(cuda_device, dtype):
    from refiners_b200 import backend as B

    x = torch.randn(3, 64, 10, 14, generator=torch.Generator().manual_seed(1)).to(cuda_device, dtype)
    for k in (2, 3):
        want = torch.nn.functional.avg_pool2d(x.float(), k)
        got = B.avg_pool2d(x, k)
        assert got.shape == want.shape
        eps = {torch.float32: 1e-6, torch.bfloat16: 2**-8, torch.float16: 2**-11}[dtype]
        assert (got.float() - want).abs().max().item() <= eps * want.abs().max().item() + 1e-7


@pytest.mark.gpu
def test_t2i_adapter_under_cuda_graph(cuda_device):
    """The feature maps reach the UNet as a TUPLE of tensors in a context: the captured graph reads them at their own
    addresses, so a replay must follow new features (identity-based signature -> re-capture), a new scale (value epoch),
    and otherwise replay without re-capturing."""
    from refiners_b200.engine.graph import GraphedChain

    dtype = torch.bfloat16
    unet = SD1UNet(4, device="meta")
    keyed(unet, 1, cuda_device, dtype)
    adapter = SD1T2IAdapter(unet, name="depth", scale=0.8)
    keyed(adapter.condition_encoder, 31, cuda_device, dtype)
    adapter.inject()
    x = keyed_input("t2i.sd1.x", (1, 4, 32, 32)).to(cuda_device, dtype)
    ctx = keyed_input("t2i.sd1.ctx", (1, 77, 768)).to(cuda_device, dtype)
    timestep = torch.tensor([601.0], device=cuda_device)

    def contexts(features):
        unet.set_timestep(timestep)
        unet.set_clip_text_embedding(ctx)
        adapter.set_condition_features(features)

    with no_grad():
        f1 = adapter.compute_condition_features(keyed_input("t2i.sd1.condition", (1, 3, 256, 256)).to(cuda_device, dtype))
        f2 = tuple(f * 0.5 for f in f1)
        runner = GraphedChain(unet)
        for features, scale in ((f1, 0.8), (f1, 0.8), (f2, 0.8), (f2, 0.3), (f1, 0.3)):
            adapter.scale = scale
            contexts(features)
            want = unet(x)
            contexts(features)
            got = runner(x).clone()
            assert torch.equal(got, want), (scale, features is f1)
        assert runner.captures == 4 and runner.replays == 5  # the second call replays, every change re-captures
