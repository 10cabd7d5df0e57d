"""One process per GPU: frame sharding and the single collective of the path.

Frames and evaluation images are independent and the 3x3 stencil never crosses an image border, so the
data path needs no exchange.  The only collective is an all-reduce(sum) of a handful of float64
accumulators: the eight pooled metric sums, or {n_images, sum over images of the 7 per-image metrics}
for the reference's mean-over-images semantics (trainer.py:1426-1428), plus output checksums for the
sequence benchmark.  Backend: NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, local_rank, world)."""
    rank, local_rank, world = env_world()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(total, rank, world):
    """Contiguous block [lo, hi) of `total` items owned by `rank`: sizes differ by at most one."""
    base, extra = divmod(int(total), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def all_reduce_sums(t):
    """In-place sum of a small float64 tensor over all ranks (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def barrier():
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def max_over_ranks(value, device):
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def mean_over_images(per_image_metrics):
    """Reference eval semantics: mean over ALL images (all ranks) of the per-image metric rows [B_local, 7].

    Each rank finalises its own images, then one all-reduce of {n_images, 7 sums} (8 float64).  NaN rows
    (empty masks) poison the mean exactly as np.array(errors).mean(0) does in the reference.
    """
    rows = per_image_metrics.to(torch.float64)
    acc = torch.cat((torch.tensor([float(rows.shape[0])], dtype=torch.float64, device=rows.device), rows.sum(dim=0)))
    all_reduce_sums(acc)
    return acc[1:] / acc[0]
