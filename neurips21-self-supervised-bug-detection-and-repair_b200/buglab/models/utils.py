"""Scatter wrappers, optimiser and LR schedule of the BugLab models (reference: buglab/models/utils.py:15-66).

``scatter_log_softmax`` is one fused kernel chain (``bl_segment_log_softmax_fwd/_bwd``) instead of six launches;
the optimiser is the fused flat-buffer Adam (clip folded in)."""
from typing import Optional

import torch
import torch_scatter
from ptgnn.baseneuralmodel import AbstractScheduler
from torch.optim.lr_scheduler import LambdaLR

from buglab_b200 import ops
from buglab_b200.flat import FlatAdam


def scatter_log_softmax(src: torch.Tensor, index: torch.Tensor, dim: int = -1, eps: float = 1e-12) -> torch.Tensor:
    if not torch.is_floating_point(src):
        raise ValueError("`scatter_log_softmax` can only be computed over tensors with floating point data types.")
    return ops.segment_log_softmax(src, index, eps=eps)


def scatter_sum(src, index, dim: int = -1, dim_size: Optional[int] = None):
    return torch_scatter.scatter_sum(src, index, dim, dim_size=dim_size)


def scatter_mean(src, index, dim: int = -1, dim_size: Optional[int] = None):
    return torch_scatter.scatter_mean(src.float(), index, dim, dim_size=dim_size)


def scatter_max(src, index, dim: int = -1, dim_size: Optional[int] = None):
    return torch_scatter.scatter_max(src.float(), index, dim, dim_size=dim_size)


def scatter_min(src, index, dim: int = -1, dim_size: Optional[int] = None):
    return torch_scatter.scatter_min(src.float(), index, dim, dim_size=dim_size)


def optimizer(p, lr: float = 0.0001) -> torch.optim.Optimizer:
    """Adam(lr=1e-4, torch defaults) — reference utils.py:51-52 — as the fused flat-buffer implementation."""
    return FlatAdam(p, lr=lr)


class LinearWarmupScheduler(AbstractScheduler):
    """LR x min(1, step / num_warmup_steps), stepped once per minibatch (reference utils.py:55-66)."""

    def __init__(self, optimizer, num_warmup_steps: int = 800, last_epoch=-1):
        self.__num_warmup_steps = num_warmup_steps
        self.__scheduler = LambdaLR(optimizer, self.lr_lambda, last_epoch=last_epoch)

    def lr_lambda(self, current_step: int):
        if current_step < self.__num_warmup_steps:
            return float(current_step) / float(max(1.0, self.__num_warmup_steps))
        return 1.0

    def step(self, epoch_idx: int, epoch_step: int) -> None:
        self.__scheduler.step()
