#!/usr/bin/env python
"""bench.py - SDXL 1024^2 bf16 denoising steps/s on N x B200 (BASELINE.json metric).

    python bench.py --gpus 1 --steps 30 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 ... bench.py --gpus N ...
    python bench.py --impl reference ...      # the reference algorithm's CPU path (oracle port), host cores

A "step" is one ``StableDiffusion_XL(x, step=..)`` call at latent batch 8 with classifier-free
guidance (UNet batch 16): set contexts, sigma-scale, SDXLUNet forward, CFG combine, Euler update
(BASELINE configs[1]; random-init weights, synthetic embeddings).  Each rank runs its own batch
of 8 latents (weak scaling, no per-step collective; NCCL only broadcasts the weights at init).

One JSON line is printed by rank 0: value (device-resident inputs), e2e (host buffers through the
public API, H2D/D2H inside the timed region), roofline of the dominant kernel measured live,
cpu_baseline (oracle port on the host cores, bounded sample), clocks, gpu_launches.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

SDXL_TFLOP_PER_SAMPLE = 6.761  # algorithmic FLOPs of one SDXLUNet forward per UNet-batch row
LATENT_BATCH = 8
CFG_ROWS = 2 * LATENT_BATCH
METRIC = "SDXL 1024^2 bf16 denoising steps/s (latent batch 8, CFG, Euler)"
UNIT = "steps/s"


T0 = time.time()


def log(msg: str) -> None:
    print(f"[bench {time.time() - T0:7.1f}s] {msg}", file=sys.stderr, flush=True)


def measured_peaks() -> tuple[dict, str]:
    path = ROOT / "MEASURED_PEAKS.json"
    if path.exists():
        try:
            return json.loads(path.read_text()), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def usable_cores() -> int:
    """Host threads this process can really use: affinity mask and cgroup CPU quota, capped at 32
    (ATen's CPU kernels stop scaling - and oversubscribed boxes collapse - well before 128 threads)."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        quota, period = Path("/sys/fs/cgroup/cpu.max").read_text().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, min(n, 32))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    FIELDS = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int) -> None:
        self.index = index
        self.rows: list[list[str]] = []
        self._stop = threading.Event()
        self._thread: threading.Thread | None = None

    def _run(self) -> None:
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.splitlines()[0].split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self) -> "ClockSampler":
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc) -> None:
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)

    def summary(self) -> dict:
        sm = sorted(int(float(r[0])) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [int(float(r[1])) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for name, val in zip(names, r[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {
            "sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------- CPU baseline
def cpu_reference_step_seconds(warmup: int, steps: int, threads: int) -> tuple[float, str]:
    """Time the oracle port of the reference's SDXL UNet on the host cores (fp32).
    Bounded sample: ONE UNet-batch row (1/16 of a batch-8 CFG step) at the full 1024^2
    resolution per timed iteration."""
    from oracle import ops as oops
    from oracle import unet as ounet
    from oracle.weights import keyed_state_dict

    oops.FAST = True  # call the same fused ATen CPU kernels the reference calls (see oracle/ops.py)
    from refiners_b200.foundationals.latent_diffusion import SDXLUNet

    torch.set_num_threads(threads)
    shapes = {k: tuple(v.shape) for k, v in SDXLUNet(4, device="meta").state_dict().items()}
    sd = keyed_state_dict(shapes, seed=2)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 128, 128, generator=g)
    ctx, pooled = torch.randn(1, 77, 2048, generator=g), torch.randn(1, 1280, generator=g)
    ids = torch.tensor([[1024, 1024, 0, 0, 1024, 1024]])
    ts = torch.tensor([981.0])
    with torch.no_grad():
        for _ in range(warmup):
            ounet.sdxl_unet(sd, x, ts, ctx, pooled, ids)
        t0 = time.perf_counter()
        for _ in range(steps):
            ounet.sdxl_unet(sd, x, ts, ctx, pooled, ids)
        dt = (time.perf_counter() - t0) / steps
    return dt, "1 UNet-batch row (1/16 of a latent-batch-8 CFG step) at 128x128 latents, fp32, oracle port calling the reference's ATen CPU ops"


def run_reference_arm(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = usable_cores()
    dt_row, sample = cpu_reference_step_seconds(max(1, min(args.warmup, 1)), max(1, args.steps), threads)
    step_s = dt_row * CFG_ROWS
    value = 1.0 / step_s
    line = {
        "impl": "reference",
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SDXLUNet 1024^2, 30-step Euler, latent batch 8 + CFG (UNet batch 16)",
                   "timed": "each timed iteration = " + sample + "; ms_per_step = 16 x that"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
def dominant_kernel_roofline(device: torch.device, peaks: dict, peaks_kind: str) -> dict:
    """tc_gemm on the most frequent SDXL GEMM (240 of 743 Linear calls per forward):
    [B*1024, 1280] x [1280, 1280]^T at UNet batch 16, timed alone with CUDA events, L2 flushed
    between launches."""
    from refiners_b200 import backend as B

    M, K, N = CFG_ROWS * 1024, 1280, 1280
    x = torch.randn(M, K, device=device, dtype=torch.bfloat16)
    w = torch.randn(N, K, device=device, dtype=torch.bfloat16) * 0.03
    flush = torch.empty(256 * 1024 * 1024, device=device, dtype=torch.uint8)
    with torch.no_grad():
        for _ in range(5):
            B.linear(x, w)
        torch.cuda.synchronize()
        times = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            B.linear(x, w)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
    ms = sum(times) / len(times)
    flops = 2.0 * M * N * K
    achieved = flops / (ms * 1e-3) / 1e12
    peak = float(peaks["bf16_tflops"])
    return {
        "bound": "tensor", "kernel": "tc_gemm_kernel<bf16,256,cta_group::2 pair> [16384x1280]x[1280x1280]^T", "achieved": achieved,
        "peak": peak, "peak_source": f"{peaks_kind} bf16_tflops (burst: kernel timed alone)", "unit": "TFLOP/s",
        "frac": achieved / peak,
        # dram__bytes_read.sum + dram__bytes_write.sum of this kernel on this shape, one `ncu --set full` launch
        # (profiles/r01_ncu_full_gemm_full.txt, captured in the multicast-pair mode: DRAM traffic is the compulsory
        # operand read either way); algorithmic bytes are 87.2 MB, the output mostly stays in L2
        "traffic": 47.49e6, "traffic_source": "profiles/r01_ncu_full_gemm_full.txt", "ms_per_launch": ms,
        "algorithmic_flops_per_launch": flops,
    }


def run_gpu_arm(args) -> None:
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the GPU arm has no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    from refiners_b200 import backend as B
    from refiners_b200.fluxion.utils import manual_seed, no_grad
    from refiners_b200.foundationals.latent_diffusion import Euler, SDXLUNet, StableDiffusion_XL

    B.load_library()
    dtype = torch.bfloat16
    manual_seed(0)
    log('building SDXLUNet (random init on device)')
    unet = SDXLUNet(in_channels=4, device=device, dtype=dtype)
    if world > 1:  # identical replicas: weights come from rank 0 over NCCL/NVLink, once
        from refiners_b200.engine.sharding import broadcast_parameters

        broadcast_parameters(unet, src=0)
    sdxl = StableDiffusion_XL(unet=unet, solver=Euler(num_inference_steps=30), device=device, dtype=dtype)
    if not args.no_graph and not args.profile_step:
        sdxl.enable_cuda_graph()

    log('model ready')
    g = torch.Generator().manual_seed(1000 + rank)
    lb = args.latent_batch
    host = {
        "x": (torch.randn(lb, 4, 128, 128, generator=g) * float(sdxl.solver.init_noise_sigma)).to(dtype).pin_memory(),
        "clip": torch.randn(2 * lb, 77, 2048, generator=g).to(dtype).pin_memory(),
        "pooled": torch.randn(2 * lb, 1280, generator=g).to(dtype).pin_memory(),
        "ids": torch.tensor([[1024, 1024, 0, 0, 1024, 1024]]).repeat(2 * lb, 1).pin_memory(),
    }
    dev = {k: v.to(device) for k, v in host.items()}
    out_host = torch.empty_like(host["x"]).pin_memory()

    def step_resident(x: torch.Tensor, s: int) -> torch.Tensor:
        return sdxl(x, step=s % 30, clip_text_embedding=dev["clip"], pooled_text_embedding=dev["pooled"], time_ids=dev["ids"])

    def step_e2e(s: int) -> None:
        x = host["x"].to(device, non_blocking=True)
        clip = host["clip"].to(device, non_blocking=True)
        pooled = host["pooled"].to(device, non_blocking=True)
        ids = host["ids"].to(device, non_blocking=True)
        y = sdxl(x, step=s % 30, clip_text_embedding=clip, pooled_text_embedding=pooled, time_ids=ids)
        out_host.copy_(y, non_blocking=True)

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    if args.profile_step:
        with no_grad():
            for s in range(2):
                step_resident(dev["x"], s)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
            step_resident(dev["x"], 2)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        print(json.dumps({"profiled_step": True, "launches_in_step": B.launch_count()}), flush=True)
        return

    with no_grad():
        x = dev["x"]
        for s in range(max(args.warmup, 3)):  # includes the capture
            step_resident(x, s)
        log('warm-up / capture done')
        launches0 = B.launch_count()
        runner = sdxl._graphed_unet[0] if sdxl._graphed_unet else None
        replays0 = runner.replays if runner else 0
        barrier()
        sampler = ClockSampler(local_rank)
        with sampler:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for s in range(args.steps):
                x = step_resident(dev["x"], s)
            e1.record()
            barrier()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        eager_launches = B.launch_count() - launches0
        graph_launches = (runner.replays - replays0) * runner.launches_per_replay if runner else 0
        gpu_launches = eager_launches + graph_launches

        log(f'resident loop done: {ms_total / args.steps:.2f} ms/step')
        if args.resident_only:
            return
        # end to end: host buffers in, host result out, every step
        for s in range(3):
            step_e2e(s)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in range(args.steps):
            step_e2e(s)
        e1.record()
        barrier()
        ms_e2e = max_over_ranks(e0.elapsed_time(e1))

    ms_per_step = ms_total / args.steps
    value = world * 1000.0 / ms_per_step
    e2e_value = world * 1000.0 / (ms_e2e / args.steps)
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    d2h = out_host.numel() * out_host.element_size()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    log(f'e2e loop done: {ms_e2e / args.steps:.2f} ms/step')
    peaks, peaks_kind = measured_peaks()
    roofline = dominant_kernel_roofline(device, peaks, peaks_kind)
    log('kernel roofline done')
    step_tflops = SDXL_TFLOP_PER_SAMPLE * 2 * lb / (ms_per_step * 1e-3)
    roofline_step = {
        "bound": "tensor", "achieved": step_tflops, "peak": float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])),
        "peak_source": f"{peaks_kind} bf16_tflops_sustained (kernel timed inside a long step)", "unit": "TFLOP/s",
        "frac": step_tflops / float(peaks.get("bf16_tflops_sustained", peaks["bf16_tflops"])),
        "algorithmic_tflop_per_step": SDXL_TFLOP_PER_SAMPLE * 2 * lb,
    }
    cpu = None
    if not args.skip_cpu_baseline:
        threads = usable_cores()
        log(f'cpu baseline on {threads} threads (os.cpu_count() = {os.cpu_count()})')
        dt_row, sample = cpu_reference_step_seconds(1, 1, threads)
        log('cpu baseline done')
        cpu = {"value": 1.0 / (dt_row * CFG_ROWS), "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
               "seconds_per_sample": dt_row}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {
            "workload": "SDXLUNet 1024^2 bf16, 30-step Euler, latent batch %d + CFG (UNet batch %d) per GPU" % (lb, 2 * lb),
            "weights": "random init (seed 0), broadcast from rank 0", "cuda_graph": not args.no_graph,
            "l2": "per-step working set (5.1 GB weights + activations) exceeds the 126 MB L2; no flush needed",
            "parallelism": f"replicas x{world} (batch-sharded, no per-step collective)",
        },
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(gpu_launches),
        "roofline": roofline,
        "roofline_step": roofline_step,
        "cpu_baseline": cpu,
        "clocks": sampler.summary(),
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--latent-batch", type=int, default=LATENT_BATCH)
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--resident-only", action="store_true", help="stress mode: skip the e2e loop and the extras")
    ap.add_argument("--profile-step", action="store_true",
                    help="run ONE eager step between cudaProfilerStart/Stop and exit (for ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
