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
