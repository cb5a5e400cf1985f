"""SAM ViT-H image encoder throughput (BASELINE config 5): images/s for a batch sweep on one GPU.

    python tools/bench_sam.py [--batches 1 2 4 8] [--steps 5]
Random-init weights, synthetic 1024x1024 images, bf16, eager (no CUDA graph), CUDA-event timing.
Multi-GPU is replicas by image (refiners_b200.engine.sharding); run one process per GPU."""

import argparse
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from refiners_b200 import backend as B  # noqa: E402
from refiners_b200.fluxion.utils import manual_seed, no_grad  # noqa: E402
from refiners_b200.foundationals.segment_anything import SAMViTH  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batches", type=int, nargs="+", default=[1, 2, 4, 8])
ap.add_argument("--steps", type=int, default=5)
args = ap.parse_args()
dev = torch.device("cuda")
manual_seed(0)
model = SAMViTH(device=dev, dtype=torch.bfloat16)
for m in model.modules():  # rel-pos tables are zero-initialised; make them non-trivial
    if hasattr(m, "horizontal_embedding"):
        m.horizontal_embedding.data.normal_(0, 0.02)
        m.vertical_embedding.data.normal_(0, 0.02)
SAM_TFLOP_PER_IMAGE = 5.9
for b in args.batches:
    x = torch.randn(b, 3, 1024, 1024, device=dev, dtype=torch.bfloat16)
    with no_grad():
        for _ in range(2):
            y = model(x)
        torch.cuda.synchronize()
        n0 = B.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            y = model(x)
        e1.record()
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"metric": "SAM ViT-H image encoder images/s", "batch": b, "value": b * 1000.0 / ms, "unit": "images/s",
                      "ms_per_batch": ms, "tflops": SAM_TFLOP_PER_IMAGE * b / (ms * 1e-3), "out_shape": list(y.shape),
                      "gpu_launches_per_batch": (B.launch_count() - n0) // args.steps}))
