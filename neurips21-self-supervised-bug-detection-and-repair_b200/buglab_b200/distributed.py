"""Data-parallel plumbing: one process per GPU, graphs sharded across ranks, ONE fp32 all-reduce of the flat
gradient buffer per step (NCCL over NVLink/NVSwitch on the B200 box; gloo in the CPU tests).

The reference trains on a single device only; this is new functionality.  There is no
exchange inside the forward/backward (graphs are independent, buglab/models/gnn.py:251 averages over
graphs), so the all-reduce is the only collective.  Global-norm clipping happens AFTER the all-reduce on
identical buffers, so all ranks apply bit-identical updates.
"""
import os
from typing import Iterable, Iterator, Optional, TypeVar

import torch
import torch.distributed as dist

T = TypeVar("T")


def is_distributed() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def world_size() -> int:
    return dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1


def rank() -> int:
    return dist.get_rank() if (dist.is_available() and dist.is_initialized()) else 0


def init_from_env(backend: Optional[str] = None) -> int:
    """Initialise torch.distributed from torchrun's env (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*). Returns local rank."""
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not (dist.is_available() and dist.is_initialized()):
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
        dist.init_process_group(backend=backend)
    return local_rank


def shard_for_rank(items: Iterable[T], rank_: Optional[int] = None, world: Optional[int] = None) -> Iterator[T]:
    """Round-robin sharding of an iterable of independent units (shard files or graphs) across ranks."""
    r = rank() if rank_ is None else rank_
    w = world_size() if world is None else world
    for i, item in enumerate(items):
        if i % w == r:
            yield item


def allreduce_flat_gradient(flat_grad: torch.Tensor) -> float:
    """Sum the flat gradient bucket over all ranks (async on the current stream for NCCL).
    Returns the scale (1/world) the optimiser must apply — averaging is folded into the Adam kernel."""
    if not is_distributed():
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM)
    return 1.0 / dist.get_world_size()


class OverlappedGradientReducer:
    """All-reduces a flat gradient buffer bucket by bucket WHILE backward is still running.

    The parameters' ``.grad`` are views into one flat buffer (``FlatAdam``); buckets are contiguous slices of it, formed
    over the parameters in reverse registration order (the order backward produces them: last layers first).  A
    post-accumulate hook on every parameter counts its bucket down; a bucket is reduced on a side stream as soon as it
    AND every earlier bucket are complete, so all ranks issue the same collectives in the same order no matter which
    parameters a rank's minibatch leaves unused (their buckets are flushed, in order, by :meth:`finish`).  The global-norm
    clip and the Adam step run after :meth:`finish`, on identical buffers on all ranks.

    ``weight`` (see trainer._RankSync) scales this rank's gradient before the reduction, on the side stream."""

    def __init__(self, flat_grad: torch.Tensor, params, offsets, bucket_bytes: int = 32 << 20):
        self.flat = flat_grad
        self._cuda = flat_grad.is_cuda
        order = sorted(range(len(params)), key=lambda i: -offsets[i])  # reverse registration order
        self._buckets = []  # (lo, hi) element ranges, in launch order
        self._bucket_of = {}
        lo = hi = None
        count = 0
        for i in order:
            start, end = offsets[i], offsets[i] + params[i].numel()
            lo = start if lo is None else min(lo, start)
            hi = end if hi is None else max(hi, end)
            self._bucket_of[i] = len(self._buckets)
            count += 1
            if (hi - lo) * 4 >= bucket_bytes:
                self._buckets.append((lo, hi, count))
                lo = hi = None
                count = 0
        if count:
            self._buckets.append((lo, hi, count))
        # the padding between parameter views belongs to no parameter: extend the buckets so they tile the buffer
        edges = sorted(b[0] for b in self._buckets)
        self._buckets = [(b[0], min([e for e in edges if e > b[0]], default=flat_grad.numel()), b[2]) for b in self._buckets]
        self._stream = torch.cuda.Stream(flat_grad.device) if self._cuda else None
        self._handles = [p.register_post_accumulate_grad_hook(lambda _p, i=i: self._on_grad(self._bucket_of[i]))
                         for i, p in enumerate(params)]
        self._pending, self._next, self._weight, self._active = [], 0, 1.0, False
        self.exposed_ms_events = None

    @property
    def num_buckets(self) -> int:
        return len(self._buckets)

    def begin(self, weight: float = 1.0) -> None:
        """Call before ``loss.backward()``."""
        self._pending = [b[2] for b in self._buckets]
        self._next, self._weight, self._active = 0, float(weight), is_distributed()

    def _on_grad(self, bucket: int) -> None:
        if not self._active:
            return
        self._pending[bucket] -= 1
        while self._next < len(self._buckets) and self._pending[self._next] <= 0:
            self._launch(self._next)
            self._next += 1

    def _launch(self, bucket: int) -> None:
        lo, hi, _ = self._buckets[bucket]
        chunk = self.flat[lo:hi]
        if self._cuda:
            ready = torch.cuda.Event()
            ready.record(torch.cuda.current_stream(self.flat.device))
            with torch.cuda.stream(self._stream):
                self._stream.wait_event(ready)
                if self._weight != 1.0:
                    chunk.mul_(self._weight)
                dist.all_reduce(chunk, op=dist.ReduceOp.SUM)
        else:
            if self._weight != 1.0:
                chunk.mul_(self._weight)
            dist.all_reduce(chunk, op=dist.ReduceOp.SUM)

    def finish(self) -> float:
        """Call after ``loss.backward()``: flushes the remaining buckets, makes the compute stream wait for the reductions
        and returns the scale (1/world) the optimiser applies.  A no-op (scale 1) outside data-parallel runs."""
        if not self._active:
            if self._weight != 1.0:
                self.flat.mul_(self._weight)
            return 1.0
        compute = torch.cuda.current_stream(self.flat.device) if self._cuda else None
        if self._cuda:
            backward_done = torch.cuda.Event(enable_timing=True)
            backward_done.record(compute)
        while self._next < len(self._buckets):
            self._launch(self._next)
            self._next += 1
        if self._cuda:
            compute.wait_stream(self._stream)
            reduced = torch.cuda.Event(enable_timing=True)
            reduced.record(compute)
            self.exposed_ms_events = (backward_done, reduced)
        self._active = False
        return 1.0 / dist.get_world_size()

    def exposed_ms(self) -> Optional[float]:
        """Time the compute stream spent waiting for the reductions after the last backward kernel of the latest step
        (synchronises on two events; for reporting, not for the training loop)."""
        if self.exposed_ms_events is None:
            return None
        a, b = self.exposed_ms_events
        b.synchronize()
        return a.elapsed_time(b)

    def close(self) -> None:
        for h in self._handles:
            h.remove()
        self._handles = []


def broadcast_module(module: torch.nn.Module, src: int = 0) -> None:
    """Make every rank start from rank ``src``'s parameters and buffers."""
    if not is_distributed():
        return
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src=src)


def all_ranks_max(value: float, device) -> float:
    """Max over ranks of a host scalar (used for max-over-ranks device timings)."""
    if not is_distributed():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_ranks_sum(value: float, device) -> float:
    if not is_distributed():
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
