"""How the oracle's CPU SDXL forward (fast ATen mode) scales with threads on this host."""
import sys, time
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from oracle import ops, unet as O
from oracle.weights import keyed_state_dict
from refiners_b200.foundationals.latent_diffusion import SDXLUNet
ops.FAST = True
shapes = {k: tuple(v.shape) for k, v in SDXLUNet(4, device="meta").state_dict().items()}
sd = keyed_state_dict(shapes, seed=2)
x = torch.randn(1, 4, 128, 128); ctx = torch.randn(1, 77, 2048); pooled = torch.randn(1, 1280)
ids = torch.tensor([[1024, 1024, 0, 0, 1024, 1024]]); ts = torch.tensor([981.0])
for n in [int(a) for a in sys.argv[1:]] or [8, 16, 32]:
    torch.set_num_threads(n)
    with torch.no_grad():
        t = time.perf_counter(); O.sdxl_unet(sd, x, ts, ctx, pooled, ids); dt = time.perf_counter() - t
    print(f"threads={n}: {dt:.1f} s per UNet-batch row", flush=True)
