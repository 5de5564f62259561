"""Data-parallel plumbing shared by bench.py and the gloo tests: one process per GPU, per-rank batch, DDP
gradient all-reduce (the reference's only parallelism: lib/training/training.py:148-153), max-over-ranks timing."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .synthetic import make_batch

RAW_KEYS = ("num_nodes", "node_mask", "node_features", "distance_matrix", "feature_matrix", "dft_coords", "target")


def env_rank_world():
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))


def rank_batch(B: int, N: int, rank: int):
    """Each rank owns its own `B` graphs (weak scaling, the reference's per-GPU batch_size semantics)."""
    host = make_batch(B, N, seed=1 + rank, with_3d=False)
    return {k: host[k] for k in RAW_KEYS}


def wrap_ddp(model: torch.nn.Module, world: int, device_ids=None, bucket_cap_mb: int = 25,
             bf16_grads: bool = False) -> torch.nn.Module:
    """torch DDP exactly as the reference wraps its model (training.py:148-153) plus the in-library knobs SURVEY 5.9
    lists: buckets as views of the gradients, bucket size, and the stock bf16 gradient-compression hook (the all-reduce
    moves half the bytes; gradients are decompressed back to fp32 before the optimizer sees them)."""
    if world <= 1:
        return model
    ddp = torch.nn.parallel.DistributedDataParallel(model, device_ids=device_ids, gradient_as_bucket_view=True,
                                                    bucket_cap_mb=bucket_cap_mb)
    if bf16_grads:
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        ddp.register_comm_hook(state=None, hook=default_hooks.bf16_compress_hook)
    return ddp


def max_over_ranks(value: float, world: int, device) -> float:
    t = torch.tensor([value], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
