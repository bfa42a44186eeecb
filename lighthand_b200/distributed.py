"""Multi-GPU plumbing: the pipeline is data-partitioned (each rank owns a contiguous slice of the
batch / dataset; K1-K4 exchange nothing).  The only exchange is the sum of the packed metric buffers
(integer PCK counts + float64 EPE/loss sums, ~1 KB), enqueued on the compute stream right after K5.

One process per GPU, torch.distributed for the rendezvous.  On GPUs the collective itself is the C ABI's
``lh_metrics_allreduce_multi`` (csrc/comm.cu): ONE grouped NCCL operation over every buffer, in place,
on the caller's stream, CUDA-graph capturable, through a communicator this module owns (its 128-byte
id travels over the torch.distributed store).  CPU tensors (the gloo tests of the host logic) go through
``dist.all_reduce``.
"""
from __future__ import annotations

import ctypes as C
from typing import Iterable, Optional, Tuple

import torch
import torch.distributed as dist

from . import _lib


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


class MetricComm:
    """NCCL communicator for the metric all-reduce, created through the C ABI (lh_comm_*).  Collective
    constructor: every rank of the torch.distributed default group must call it."""

    def __init__(self, device: Optional[torch.device] = None):
        self.lib = _lib.load()
        self.rank, self.world = world()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        ident = (C.c_char * _lib.LH_COMM_ID_BYTES)()
        if self.rank == 0:
            _lib.check(self.lib.lh_comm_unique_id(ident), "lh_comm_unique_id")
        if self.world > 1:
            box = [bytes(ident.raw)]
            dist.broadcast_object_list(box, src=0)       # the id rides on the torch.distributed store / backend
            ident = (C.c_char * _lib.LH_COMM_ID_BYTES).from_buffer_copy(box[0])
        self._h = C.c_void_p()
        _lib.check(self.lib.lh_comm_init(ident, self.world, self.rank, self.device.index or 0, C.byref(self._h)), "lh_comm_init")
        # first collective of a communicator builds its channels (milliseconds): pay for it here, not in a timed loop
        warm_c = torch.zeros(4, dtype=torch.int64, device=self.device)
        warm_s = torch.zeros(4, dtype=torch.float64, device=self.device)
        self.all_reduce([warm_c, warm_s])
        torch.cuda.synchronize(self.device)

    def all_reduce(self, tensors: Iterable[torch.Tensor], stream: Optional[int] = None) -> None:
        """In-place SUM over ranks of int64/uint64 and float64 CUDA tensors: one grouped NCCL launch on ``stream``
        (default: torch's current stream)."""
        ts = list(tensors)
        n = len(ts)
        if n == 0:
            return
        ptrs, elems, f64 = (C.c_void_p * n)(), (C.c_size_t * n)(), (C.c_int * n)()
        for i, t in enumerate(ts):
            if not t.is_cuda or not t.is_contiguous():
                raise _lib.LighthandError("MetricComm.all_reduce: contiguous CUDA tensors required")
            if t.dtype in (torch.int64, torch.uint64):
                f64[i] = 0
            elif t.dtype == torch.float64:
                f64[i] = 1
            else:
                raise _lib.LighthandError(f"MetricComm.all_reduce: dtype {t.dtype} is not a metric accumulator type")
            ptrs[i], elems[i] = t.data_ptr(), t.numel()
        st = torch.cuda.current_stream(self.device).cuda_stream if stream is None else stream
        _lib.check(self.lib.lh_metrics_allreduce_multi(self._h, n, ptrs, elems, f64, st), "lh_metrics_allreduce_multi")

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self.lib.lh_comm_destroy(self._h)
            self._h = None


_COMM: Optional[MetricComm] = None


def metric_comm() -> MetricComm:
    """The process-wide communicator (created collectively on first use)."""
    global _COMM
    if _COMM is None:
        _COMM = MetricComm()
    return _COMM


def shutdown() -> None:
    global _COMM
    if _COMM is not None:
        _COMM.close()
        _COMM = None


def all_reduce_metrics(tensors: Iterable[torch.Tensor], group=None, stream: Optional[int] = None) -> None:
    """In-place SUM of additive metric buffers across ranks.  Integer counts are exact and
    order-independent (N-GPU == 1-GPU bit for bit); float64 sums differ by reduction order only.
    CUDA tensors: one grouped NCCL operation through the C ABI, no packing copies.  CPU tensors (gloo):
    one ``dist.all_reduce`` per tensor."""
    if not is_initialized() or dist.get_world_size(group) == 1:
        return
    ts = list(tensors)
    if ts and all(t.is_cuda for t in ts):
        metric_comm().all_reduce(ts, stream)
        return
    for t in ts:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
