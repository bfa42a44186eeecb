"""Multi-GPU plumbing: the pipeline is data-partitioned (each rank owns a contiguous slice of the
batch / dataset; K1-K4 exchange nothing).  The only exchange is the sum of the packed metric buffers
(integer PCK counts + float64 EPE/loss sums, ~1 KB), enqueued on the compute stream right after K5.

One process per GPU, torch.distributed for the rendezvous; backend "nccl" on GPUs (NVLink 5 /
NVSwitch), "gloo" in the CPU tests of the host logic.
"""
from __future__ import annotations

from typing import Iterable, Tuple

import torch
import torch.distributed as dist


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def world() -> Tuple[int, int]:
    if is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n_total: int, rank: int = None, world_size: int = None) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of ``n_total`` samples owned by ``rank``; remainders go to the first
    ranks so slice sizes differ by at most one (99,792 = 8 x 12,474 exactly)."""
    if rank is None or world_size is None:
        rank, world_size = world()
    base, rem = divmod(n_total, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_reduce_metrics(tensors: Iterable[torch.Tensor], group=None) -> None:
    """In-place SUM of additive metric buffers across ranks.  Integer counts are exact and
    order-independent (N-GPU == 1-GPU bit for bit); float64 sums differ by reduction order only.
    Buffers of the same dtype are packed so each dtype costs one collective."""
    if not is_initialized() or dist.get_world_size(group) == 1:
        return
    by_dtype = {}
    for t in tensors:
        by_dtype.setdefault(t.dtype, []).append(t)
    for dtype, ts in by_dtype.items():
        if len(ts) == 1 and ts[0].is_contiguous():
            dist.all_reduce(ts[0], op=dist.ReduceOp.SUM, group=group)
            continue
        flat = torch.cat([t.reshape(-1) for t in ts])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for t in ts:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
