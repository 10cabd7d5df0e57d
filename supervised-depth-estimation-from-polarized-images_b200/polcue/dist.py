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


class PeerExchange:
    """The sum over ranks of a small float64 vector through NVLink peer memory (`polcue_peer_*`, csrc/peer.cuh): one CTA per
    rank stores its values into every rank's exchange block and adds what arrived in rank order, so all ranks hold the
    identical, bitwise reproducible sum.  `polcue.ops.eval_pass(..., peer=this)` fuses the exchange into the pass's last
    kernel; `all_reduce` is the stand-alone launch (ranks without images take part through it).

    One process per GPU on ONE node (CUDA IPC).  Collective over `group`: every rank constructs it, makes the same sequence
    of exchanges on stream-ordered calls, and closes it.  The IPC handles travel through `all_gather_object`, so any
    torch.distributed backend serves (NCCL in production, gloo in tests)."""

    MAX_VALUES = 128

    def __init__(self, device=None, group=None):
        import ctypes as C

        from . import _lib
        self._lib = _lib
        self.handle = None
        on = dist.is_available() and dist.is_initialized()
        self.group = group
        self.rank = dist.get_rank(group) if on else 0
        self.world = dist.get_world_size(group) if on else 1
        if self.world > 1 and int(os.environ.get("LOCAL_WORLD_SIZE", self.world)) != self.world:
            raise RuntimeError("PeerExchange maps peer memory with CUDA IPC: all ranks must run on one node")
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        mine = C.create_string_buffer(64)
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            # Every step's outcome is agreed on by all ranks before the next one, so a failure (no IPC in this container, no
            # peer access between two devices) raises on EVERY rank instead of leaving the others inside a collective.
            rc = _lib.lib().polcue_peer_create(self.world, self.rank, C.byref(ptr), mine)
            self.handle = ptr if rc == 0 else None
            if self.world > 1:
                got = [None] * self.world
                dist.all_gather_object(got, (rc, mine.raw), group=group)
                if all(r == 0 for r, _ in got):
                    rc = _lib.lib().polcue_peer_connect(self.handle, b"".join(h for _, h in got))
                codes = [None] * self.world
                dist.all_gather_object(codes, rc, group=group)   # also the barrier: every rank has mapped every block
                rc = next((c for c in codes if c != 0), 0)
            if rc != 0:
                if self.handle is not None:
                    _lib.lib().polcue_peer_destroy(self.handle)
                    self.handle = None
                raise RuntimeError(f"PeerExchange: setting up the peer-memory blocks failed on some rank: {_lib.lib().polcue_error_string(rc).decode()}")

    def all_reduce(self, values, out=None):
        """Sum of `values` (<= 128 float64 on this rank's device) over all ranks -> `out` (a new tensor by default)."""
        if values.dtype != torch.float64 or values.device != self.device or not values.is_contiguous() or not 1 <= values.numel() <= self.MAX_VALUES:
            raise ValueError("PeerExchange.all_reduce: 1..128 contiguous float64 values on this rank's device")
        out = torch.empty_like(values) if out is None else out
        if out.dtype != torch.float64 or out.device != self.device or not out.is_contiguous() or out.numel() != values.numel():
            raise ValueError("PeerExchange.all_reduce: `out` must match `values`")
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib().polcue_peer_allreduce_f64(self.handle, values.data_ptr(), values.numel(), out.data_ptr(),
                                                                      torch.cuda.current_stream(self.device).cuda_stream),
                            "polcue_peer_allreduce_f64")
        return out

    def status(self):
        """(exchanges made, first exchange whose wait for a peer timed out or 0); synchronises the device."""
        import ctypes as C
        calls, failed = C.c_ulonglong(), C.c_ulonglong()
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib().polcue_peer_status(self.handle, C.byref(calls), C.byref(failed)), "polcue_peer_status")
        return calls.value, failed.value

    def close(self):
        if self.handle is None:
            return
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)             # no peer is still exchanging through this rank's block
        with torch.cuda.device(self.device):
            self._lib.check(self._lib.lib().polcue_peer_destroy(self.handle), "polcue_peer_destroy")
        self.handle = None


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
