"""N > 1 plumbing on CPU: two gloo ranks run replica shards of a CFG batch and reproduce the
single-process result; weights come from rank 0 by broadcast.  No per-step collective exists."""

import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import refiners_b200.fluxion.layers as fl
from refiners_b200.engine.sharding import broadcast_parameters, gather_batch, shard_cfg_batch, shard_range
from refiners_b200.fluxion.utils import no_grad
from refiners_b200.foundationals.latent_diffusion import CrossAttentionBlock2d, ResidualBlock


def build(seed: int) -> fl.Chain:
    torch.manual_seed(seed)
    return fl.Chain(
        ResidualBlock(32, 64),
        CrossAttentionBlock2d(64, context_embedding_dim=24, context_key="ctx", num_attention_heads=2, num_attention_layers=1,
                              use_bias=False, use_linear_projection=True),
    )


def inputs():
    g = torch.Generator().manual_seed(7)
    return torch.randn(4, 32, 8, 8, generator=g), torch.randn(4, 5, 24, generator=g)  # 2 latents x (uncond, cond)


def worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    model = build(seed=100 + rank)          # ranks start with DIFFERENT weights
    sent = broadcast_parameters(model, src=0)
    assert sent > 0
    x, ctx = inputs()
    xs, cs = shard_cfg_batch(x, rank, world), shard_cfg_batch(ctx, rank, world)
    assert xs.shape[0] == 2                 # one latent: its uncond row and its cond row
    model.set_context("cross_attention_block", {"ctx": cs})
    with no_grad():
        y = model(xs)
    full = gather_batch(y, world)
    if rank == 0:
        torch.save(full, os.path.join(out_dir, "gathered.pt"))
    dist.destroy_process_group()


def test_two_rank_replicas_match_single_process(tmp_path):
    assert [shard_range(5, r, 2) for r in range(2)] == [(0, 3), (3, 5)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    gathered = torch.load(tmp_path / "gathered.pt")
    model = build(seed=100)                 # rank 0's weights
    x, ctx = inputs()
    model.set_context("cross_attention_block", {"ctx": ctx})
    with no_grad():
        ref = model(x)
    # gathered rows: rank0 = (uncond0, cond0), rank1 = (uncond1, cond1); reference rows: u0, u1, c0, c1
    order = torch.tensor([0, 2, 1, 3])
    assert torch.allclose(gathered, ref[order], atol=1e-6)
