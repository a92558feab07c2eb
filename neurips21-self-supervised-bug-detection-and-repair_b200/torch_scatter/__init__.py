"""``torch_scatter`` operator surface on the buglab_b200 segment kernels.

The reference calls ``torch_scatter.scatter_{max,min,sum,mean}(src, index, dim, dim_size=)`` at
buglab/models/utils.py:20,32,38,43,48.  torch-scatter is an unpinned C++/CUDA extension that is not installable
offline; these functions keep its signatures and CPU semantics (first extreme wins, empty segments give 0 and
``arg == src.size(dim)``) and run on the sm_100a kernels behind ``include/buglab_b200.h``.  CUDA tensors only.
"""
from typing import Optional, Tuple

import torch

__version__ = "2.0.9+buglab_b200"


def _ops():
    from buglab_b200 import ops

    return ops


def scatter_max(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None, dim_size: Optional[int] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    if out is not None:
        raise NotImplementedError("`out=` is not supported")
    return _ops().segment_minmax(src, index, dim=_norm(src, dim), dim_size=dim_size, is_min=False)


def scatter_min(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None, dim_size: Optional[int] = None
                ) -> Tuple[torch.Tensor, torch.Tensor]:
    if out is not None:
        raise NotImplementedError("`out=` is not supported")
    return _ops().segment_minmax(src, index, dim=_norm(src, dim), dim_size=dim_size, is_min=True)


def scatter_sum(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None, dim_size: Optional[int] = None
                ) -> torch.Tensor:
    if out is not None:
        raise NotImplementedError("`out=` is not supported")
    return _ops().segment_sum(src, index, dim=_norm(src, dim), dim_size=dim_size)


scatter_add = scatter_sum


def scatter_mean(src: torch.Tensor, index: torch.Tensor, dim: int = -1, out=None, dim_size: Optional[int] = None
                 ) -> torch.Tensor:
    total = scatter_sum(src, index, dim, dim_size=dim_size)
    size = total.shape[0]
    count = scatter_sum(torch.ones(index.shape[0], device=src.device, dtype=torch.float32), index, 0, dim_size=size)
    count = count.clamp(min=1)
    return total / (count if total.dim() == 1 else count.view(-1, 1))


def _norm(src: torch.Tensor, dim: int) -> int:
    """1-D sources reduce along their only dim; 2-D sources only along dim 0 (the head use cases)."""
    if src.dim() == 1:
        return 0
    d = dim if dim >= 0 else src.dim() + dim
    if d != 0:
        raise NotImplementedError("buglab_b200 torch_scatter: 2-D sources are reduced along dim 0 only")
    return 0
