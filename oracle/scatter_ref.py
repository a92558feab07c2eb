"""torch_scatter (CPU) semantics restated in plain PyTorch — oracle, test infrastructure only.

torch_scatter 2.0.x CPU kernels (the release line the reference's Dockerfile:9 wheel index serves):
``scatter_max/min`` walk the source in order and update on a STRICT comparison, so the FIRST element
attaining the extreme wins; segments that receive nothing hold 0 and their arg is ``src.size(dim)``.
Backward of max/min routes the gradient to the arg element only.  Call sites in the reference:
buglab/models/utils.py:20,32,38,43,48.
"""
from typing import Optional, Tuple

import torch


def _to_2d(src: torch.Tensor, dim: int):
    if src.dim() == 1:
        return src.reshape(-1, 1), True
    if src.dim() == 2 and dim in (0, -2):
        return src, False
    raise NotImplementedError("oracle scatter ops: 1-D src, or 2-D src reduced along dim 0")


def _size(index: torch.Tensor, dim_size: Optional[int]) -> int:
    if dim_size is not None:
        return int(dim_size)
    return int(index.max()) + 1 if index.numel() else 0


def _extreme(src: torch.Tensor, index: torch.Tensor, dim: int, dim_size: Optional[int], is_min: bool):
    src2, was_1d = _to_2d(src, dim)
    L, F = src2.shape
    S = _size(index, dim_size)
    index = index.long()
    idx2 = index.view(-1, 1).expand(L, F)
    with torch.no_grad():
        fill = float("inf") if is_min else -float("inf")
        ext = torch.full((S, F), fill, dtype=src2.dtype)
        ext = ext.scatter_reduce(0, idx2, src2.detach(), "amin" if is_min else "amax", include_self=True)
        pos = torch.arange(L).view(-1, 1).expand(L, F)
        cand = torch.where(src2.detach() == ext[index], pos, torch.full_like(pos, L))
        arg = torch.full((S, F), L, dtype=torch.int64).scatter_reduce(0, idx2, cand, "amin", include_self=True)
    empty = arg >= L
    picked = torch.gather(src2, 0, arg.clamp(max=max(L - 1, 0))) if L > 0 else torch.zeros((S, F), dtype=src2.dtype)
    out = torch.where(empty, torch.zeros_like(picked), picked)
    if was_1d:
        return out.view(-1), arg.view(-1)
    return out, arg


def scatter_max(src, index, dim: int = -1, dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    return _extreme(src, index, dim, dim_size, is_min=False)


def scatter_min(src, index, dim: int = -1, dim_size: Optional[int] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    return _extreme(src, index, dim, dim_size, is_min=True)


def scatter_sum(src, index, dim: int = -1, dim_size: Optional[int] = None) -> torch.Tensor:
    src2, was_1d = _to_2d(src, dim)
    S = _size(index, dim_size)
    out = torch.zeros((S, src2.shape[1]), dtype=src2.dtype).index_add(0, index.long(), src2)
    return out.view(-1) if was_1d else out


def scatter_mean(src, index, dim: int = -1, dim_size: Optional[int] = None) -> torch.Tensor:
    src2, was_1d = _to_2d(src, dim)
    S = _size(index, dim_size)
    total = torch.zeros((S, src2.shape[1]), dtype=src2.dtype).index_add(0, index.long(), src2)
    count = torch.zeros(S, dtype=src2.dtype).index_add(0, index.long(), torch.ones(src2.shape[0], dtype=src2.dtype))
    out = total / count.clamp(min=1).view(-1, 1)
    return out.view(-1) if was_1d else out


def scatter_log_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1, eps: float = 1e-12) -> torch.Tensor:
    """Follows buglab/models/utils.py:15-28 line by line (max -> gather -> sub -> exp -> scatter_add -> log)."""
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_log_softmax` can only be computed over tensors with floating point data types.")
    max_value_per_index = scatter_max(src, index, dim=dim)[0]
    max_per_src_element = max_value_per_index.gather(dim, index.long())
    recentered = src - max_per_src_element
    sum_per_index = torch.zeros_like(max_value_per_index).scatter_add(-1, index.long(), recentered.exp())
    normalizing = (sum_per_index + eps).log().gather(dim, index.long())
    return recentered - normalizing
