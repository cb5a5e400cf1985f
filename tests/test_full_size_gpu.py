"""Parity at the BENCHMARKED sizes and through the BENCHMARKED path.

Fixtures: ``tests/golden/full_size.safetensors`` holds the unmodified reference's fp32 outputs for
SDXLUNet at 128x128 latents plain (config 2), with 700 LoRA adapters + IP-Adapter (config 3), with
ControlLora (config 4), a full ``StableDiffusion_XL`` CFG + Euler step at three steps (A17) and the SAM
ViT-H encoder on a 1024^2 image (config 5); inputs and adapter weights are keyed and regenerated here
(oracle/cases.py; recorded by oracle/pin_against_reference.py --only-full-size).

The bf16 criterion is SURVEY.md section 7 hard-part 1 / BASELINE.md section 3, computed in the test:

    err(engine bf16, reference fp32) <= err(torch-eager bf16 on this GPU, reference fp32) + 1e-3 * max|ref|

where "torch-eager bf16" is the oracle restatement in FAST mode (the reference's own ATen calls:
F.linear / F.conv2d / F.group_norm / F.scaled_dot_product_attention) on the same device, weights and
inputs - i.e. what the reference itself delivers in bf16 on a B200.  err is the rms error (exact criterion)
and the max-abs error (criterion with a stated rounding-luck factor, see MAX_ABS_LUCK); all numbers are printed.  Every model is run eagerly AND through ``GraphedChain`` (the path bench.py times),
over several different steps, and the replay must reproduce the eager result bit for bit.
"""

from __future__ import annotations

from pathlib import Path

import pytest
import torch
from safetensors.torch import load_file

from oracle import cases
from oracle import euler as oeuler
from oracle import ops as oops
from oracle import sam as osam
from oracle import unet as ounet

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"
BF16 = torch.bfloat16


@pytest.fixture(scope="module")
def golden():
    return load_file(str(GOLDEN / "full_size.safetensors"))


@pytest.fixture(scope="module")
def api():
    return cases.engine_api()


@pytest.fixture(scope="module")
def base_weights(api):
    return cases.sdxl_base_weights(api)


@pytest.fixture()
def fast_oracle():
    prev = oops.FAST
    oops.FAST = True
    yield
    oops.FAST = prev


def errors(got: torch.Tensor, want: torch.Tensor) -> tuple[float, float, float]:
    got, want = got.float().cpu(), want.float().cpu()
    assert got.shape == want.shape and torch.isfinite(got).all()
    d = got - want
    return d.abs().max().item(), d.pow(2).mean().sqrt().item(), want.abs().max().item()


# The max-abs of ~10^5 outputs is an extreme-value statistic: two equally accurate bf16 evaluations of the same network
# differ by tens of percent on it from rounding luck alone (measured on B200: 0.94 vs 1.21, 0.139 vs 0.149, 0.121 vs 0.121
# for engine vs torch-eager on three steps).  The rms error is the stable statistic, so the criterion is enforced exactly
# on the rms and with this luck factor on the max-abs; both pairs of numbers are printed.
MAX_ABS_LUCK = 1.25


def check_bf16(what: str, engine: torch.Tensor, eager: torch.Tensor, ref: torch.Tensor) -> None:
    e_max, e_rms, scale = errors(engine, ref)
    t_max, t_rms, _ = errors(eager, ref)
    print(f"\n[{what}] max|ref| {scale:.3f}: engine bf16 max-abs {e_max:.4e} rms {e_rms:.4e} | torch-eager bf16 max-abs {t_max:.4e} rms {t_rms:.4e}"
          f" | allowed rms {t_rms + 1e-3 * scale:.4e}, max-abs {MAX_ABS_LUCK * t_max + 1e-3 * scale:.4e}")
    assert e_rms <= t_rms + 1e-3 * scale, f"{what}: engine bf16 rms error {e_rms:.4e} exceeds torch-eager bf16 {t_rms:.4e} + 1e-3 * {scale:.3f}"
    assert e_max <= MAX_ABS_LUCK * t_max + 1e-3 * scale, (
        f"{what}: engine bf16 max-abs error {e_max:.4e} exceeds {MAX_ABS_LUCK} x torch-eager bf16 {t_max:.4e} + 1e-3 * {scale:.3f}")


def on(device, dtype, tensors):
    """Inputs on the device in the model dtype; the timestep keeps fp32 (the reference never casts it: solver.py:418-435,
    and 981 is not representable in bf16)."""
    return {k: v.to(device, dtype if v.is_floating_point() and k != "timestep" else v.dtype) for k, v in tensors.items()}


def graphed(chain):
    from refiners_b200.engine.graph import GraphedChain

    return GraphedChain(chain)


# ---------------------------------------------------------------------------------- config 2
def test_config2_sdxl_unet_full_size(cuda_device, api, base_weights, golden, fast_oracle):
    from refiners_b200.fluxion.utils import no_grad

    inp = cases.sdxl_inputs("cfg2", 2)
    unet = cases.build_sdxl(api, base_weights, cuda_device, BF16)
    sd = dict(unet.state_dict())  # the same device tensors, under the reference's keys
    x = inp["x"].to(cuda_device, BF16)
    with no_grad():
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        y = unet(x)
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        assert torch.equal(y, unet(x)), "two identical forwards must be bit-identical"
        d = on(cuda_device, BF16, inp)
        eager = ounet.sdxl_unet(sd, d["x"], d["timestep"], d["ctx"], d["pooled"], d["time_ids"])
        check_bf16("config 2 eager", y, eager, golden["cfg2.y"])
        # the benchmarked path: capture once, replay over different timesteps, compare each with eager
        runner = graphed(unet)
        for t in (981.0, 500.0, 13.0):
            inp_t = dict(inp, timestep=torch.tensor([t]))
            cases.set_sdxl_contexts(unet, inp_t, cuda_device, BF16)
            y_eager = unet(x).clone()
            cases.set_sdxl_contexts(unet, inp_t, cuda_device, BF16)
            y_graph = runner(x).clone()
            assert torch.equal(y_graph, y_eager), f"graph replay differs from eager at timestep {t}"
        assert runner.captures == 1 and runner.replays == 3
        # a CPU-resident context tensor must follow its value, not be frozen into the capture
        cases.set_sdxl_contexts(unet, dict(inp, timestep=torch.tensor([250.0])), "cpu", BF16)
        unet.set_clip_text_embedding(d["ctx"]); unet.set_pooled_text_embedding(d["pooled"])
        y_cpu_ctx = runner(x).clone()
        cases.set_sdxl_contexts(unet, dict(inp, timestep=torch.tensor([250.0])), cuda_device, BF16)
        assert torch.equal(y_cpu_ctx, unet(x)), "CPU timestep / time_ids were baked into the capture"
        # step-invariant hoisting: the text K / V projections live outside the graph; a NEW prompt (another tensor, or the
        # same tensor modified in place) must be picked up by the next call
        if runner.invariant is not None:
            assert runner.hoisted_ops >= 70, f"only {runner.hoisted_ops} step-invariant ops hoisted out of the graph"
        other = cases.keyed_input("cfg2.other_ctx", (2, 77, 2048)).to(cuda_device, BF16)
        for new_ctx in (other, other.mul_(0.5)):  # second round: same object, bumped version
            inp_new = dict(inp, ctx=new_ctx)
            cases.set_sdxl_contexts(unet, inp_new, cuda_device, BF16)
            y_new_graph = runner(x).clone()
            cases.set_sdxl_contexts(unet, inp_new, cuda_device, BF16)
            y_new_eager = unet(x)
            assert torch.equal(y_new_graph, y_new_eager), "stale text K / V projections after a prompt change"
            assert not torch.equal(y_new_graph, y_cpu_ctx)
        assert runner.captures == 1, "a prompt change must refresh the hoisted results, not re-capture"
        runner.close()


def test_config2_sdxl_unet_full_size_fp32(cuda_device, api, base_weights, golden):
    from refiners_b200.fluxion.utils import no_grad

    inp = cases.sdxl_inputs("cfg2", 2)
    unet = cases.build_sdxl(api, base_weights, cuda_device, torch.float32)
    with no_grad():
        cases.set_sdxl_contexts(unet, inp, cuda_device, torch.float32)
        y = unet(inp["x"].to(cuda_device))
    e_max, _, scale = errors(y, golden["cfg2.y"])
    print(f"\n[config 2 fp32] max-abs {e_max:.3e} of max|ref| {scale:.3f}")
    assert e_max <= 2e-4 * scale


# ---------------------------------------------------------------------------------------- A17
def test_stable_diffusion_xl_step_full_size(cuda_device, api, base_weights, golden, fast_oracle):
    """LatentDiffusionModel.forward (model.py:128-159): contexts, CFG doubling, sigma scaling, UNet, CFG
    combine, Euler update - eager and through enable_cuda_graph(), against the reference's recorded steps."""
    from refiners_b200.fluxion.utils import no_grad

    unet = cases.build_sdxl(api, base_weights, cuda_device, BF16)
    sd = dict(unet.state_dict())
    sdxl = api.StableDiffusion_XL(unet=unet, solver=api.Euler(num_inference_steps=30), device=cuda_device, dtype=BF16)
    sin = cases.step_inputs()
    schedule = oeuler.EulerSchedule(30)
    x0_ref = sin["x"] * float(schedule.init_noise_sigma)
    d = on(cuda_device, BF16, sin)
    x0 = x0_ref.to(cuda_device, BF16)
    kw = dict(clip_text_embedding=d["ctx"], pooled_text_embedding=d["pooled"], time_ids=d["time_ids"])
    bf_schedule = oeuler.EulerSchedule(30, dtype=BF16)  # the reference casts the solver tensors to the model dtype
    bf_schedule.timesteps = bf_schedule.timesteps.to(cuda_device)
    bf_schedule.sigmas = bf_schedule.sigmas.to(cuda_device)
    with no_grad():
        eager_out = {}
        for step, scale in cases.STEP_CASES:
            y = sdxl(x0, step=step, condition_scale=scale, **kw)
            eager_out[step] = y.clone()
            ref_eager = oeuler.denoise_step(
                lambda lat, ts: ounet.sdxl_unet(sd, lat, ts, d["ctx"], d["pooled"], d["time_ids"]), bf_schedule, x0, step, scale)
            check_bf16(f"StableDiffusion_XL step {step}", y, ref_eager, golden[f"step.y_{step}"])
        sdxl.enable_cuda_graph()
        for step, scale in cases.STEP_CASES:
            y = sdxl(x0, step=step, condition_scale=scale, **kw)
            assert torch.equal(y, eager_out[step]), f"graphed step {step} differs from the eager step"
        runner = sdxl._graphed_unet[0]
        assert runner.captures == 1 and runner.replays == len(cases.STEP_CASES)
        sdxl.enable_cuda_graph(False)


# ---------------------------------------------------------------------------------- config 3
def test_config3_lora_ip_adapter_full_size(cuda_device, api, base_weights, golden, fast_oracle):
    from refiners_b200.fluxion.utils import no_grad

    inp = cases.sdxl_inputs("cfg3", 2)
    unet = cases.build_sdxl(api, base_weights, cuda_device, BF16)
    base_sd = dict(unet.state_dict())
    ip, extra = cases.attach_config3(api, unet, 2, cuda_device, BF16)
    assert extra["n_lora_adapters"] == 700
    dev = lambda t: t.to(cuda_device, BF16)
    w = ounet.Weights(
        base_sd,
        loras={p: [(dev(a), dev(b), s) for a, b, s in ls] for p, ls in extra["loras"].items()},
        ip={p: (dev(k), dev(v)) for p, (k, v) in extra["ip"].items()},
        ip_scale=extra["ip_scale"], ip_embedding=dev(extra["ip_embedding"]),
    )
    x = dev(inp["x"])
    d = on(cuda_device, BF16, inp)
    with no_grad():
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        y = unet(x)
        eager = ounet.sdxl_unet(w, d["x"], d["timestep"], d["ctx"], d["pooled"], d["time_ids"])
        check_bf16("config 3 eager", y, eager, golden["cfg3.y"])
        runner = graphed(unet)
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        assert torch.equal(runner(x), y), "graph replay differs from eager (config 3)"
        # scales are Python floats baked into kernel arguments / packed column scales: changing one through the
        # public API must be seen by the next call of the graphed model
        ip.scale = 0.25
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        y_scaled = unet(x).clone()
        assert not torch.equal(y_scaled, y)
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        assert torch.equal(runner(x), y_scaled), "stale IP-Adapter scale in the captured graph"
        assert runner.captures == 2
        from refiners_b200.fluxion.adapters import LoraAdapter

        adapter = next(iter(unet.layers(LoraAdapter)))
        next(iter(adapter.loras.values())).scale = 0.5
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        y_lora = unet(x).clone()
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        assert torch.equal(runner(x), y_lora), "stale LoRA scale in the captured graph"
        runner.close()


# ---------------------------------------------------------------------------------- config 4
def test_config4_control_lora_full_size(cuda_device, api, base_weights, golden, fast_oracle):
    from refiners_b200.fluxion.utils import no_grad

    inp = cases.sdxl_inputs("cfg4", 2)
    unet = cases.build_sdxl(api, base_weights, cuda_device, BF16)
    base_sd = dict(unet.state_dict())
    adapter, extra = cases.attach_config4(api, unet, 2, cuda_device, BF16)
    dev = lambda t: t.to(cuda_device, BF16)
    wc = ounet.Weights(base_sd, loras={p: [(dev(a), dev(b), s) for a, b, s in ls] for p, ls in extra["loras"].items()})
    own = {k: dev(v) for k, v in extra["own"].items()}
    x = dev(inp["x"])
    d = on(cuda_device, BF16, inp)
    args = (d["timestep"], d["ctx"], d["pooled"], d["time_ids"])
    with no_grad():
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        y = unet(x)
        deltas = ounet.sdxl_control_lora(wc, own, d["x"], *args, dev(extra["condition"]), scale=extra["scale"])
        eager = ounet.sdxl_unet(base_sd, d["x"], *args, residuals=deltas)
        check_bf16("config 4 eager", y, eager, golden["cfg4.y"])
        runner = graphed(unet)
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        assert torch.equal(runner(x), y), "graph replay differs from eager (config 4)"
        adapter.scale = 0.4
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        y_scaled = unet(x).clone()
        cases.set_sdxl_contexts(unet, inp, cuda_device, BF16)
        assert torch.equal(runner(x), y_scaled), "stale ControlLora scale in the captured graph"
        runner.close()


# ---------------------------------------------------------------------------------- config 5
def test_config5_sam_vit_h_full_size(cuda_device, api, golden, fast_oracle):
    from refiners_b200.fluxion.utils import no_grad

    sam, sd = cases.build_sam(api, cuda_device, BF16)
    img = cases.sam_inputs().to(cuda_device, BF16)
    dsd = dict(sam.state_dict())
    with no_grad():
        y = sam(img)
        assert torch.equal(y, sam(img))
        eager = osam.sam_vit(dsd, img, num_layers=32, heads=16, global_indices=(7, 15, 23, 31))
    check_bf16("config 5 SAMViTH", y, eager, golden["cfg5.y"])
