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


def compute_generator_loss(arg_swap_logprobs, arrange, candidate_rewrite_idxs, candidate_symbol_to_location_group,
                           localization_logprobs, loss_type, pair_rewrite_idxs, rewrite_logprobs, rewrite_to_graph_id,
                           rewrite_to_location_group, swapped_pair_to_call_location_group, text_repair_logprobs,
                           text_rewrite_idxs, varmisuse_logprobs):
    """Selector ("bug generator") loss — reference buglab/models/utils.py:101-179, same argument order.

    Every candidate rewrite r of every graph gets the model's joint log-probability of GENERATING it,
    log P(location of r) + log P(r | location), the per-graph NO_BUG slot gets log P(NO_BUG); ``rewrite_logprobs`` holds
    the detector's log-probabilities for the same slots (-inf = not observed).  Over the observed slots, grouped by
    graph, the loss pushes the generator towards rewrites the detector finds hard (five variants).  All segment
    reductions run on the buglab_b200 segment kernels."""
    num_graphs = arrange.shape[0]
    num_rewrites = rewrite_logprobs.shape[0] - num_graphs
    slots = torch.cat((text_rewrite_idxs, candidate_rewrite_idxs, pair_rewrite_idxs, arrange + num_rewrites))
    values = torch.cat((
        localization_logprobs[rewrite_to_location_group] + text_repair_logprobs,
        localization_logprobs[candidate_symbol_to_location_group] + varmisuse_logprobs,
        localization_logprobs[swapped_pair_to_call_location_group] + arg_swap_logprobs,
        localization_logprobs[-num_graphs:],
    ))
    generation_logprobs = torch.zeros_like(rewrite_logprobs).index_put((slots,), values)  # slots are unique

    observed = torch.isinf(rewrite_logprobs).logical_not()
    index = torch.cat((rewrite_to_graph_id, arrange))[observed]
    detection = rewrite_logprobs[observed]
    generation = generation_logprobs[observed]
    if loss_type in ("norm-kl", "norm-rmse", "classify-max-loss"):
        generation = scatter_log_softmax(generation, index=index)  # renormalise over the observed slots of each graph
        if loss_type == "norm-rmse":
            log_total = torch.logaddexp(scatter_log_softmax(detection, index=index), generation)
            return (log_total ** 2).mean()
        if loss_type == "norm-kl":
            failed = torch.log(torch.clamp(1.0 - detection.exp(), min=1e-30))  # log P(detector misses the rewrite)
            kl_terms = failed.exp() * (scatter_log_softmax(failed, index=index) - generation)
            return scatter_sum(kl_terms, index=index).mean()
        _, hardest = scatter_min(detection, index=index)  # the rewrite the detector is least sure about, per graph
        return -generation[hardest].mean()
    if loss_type == "expectation":
        return scatter_sum(generation.exp() * detection, index=index).mean()
    raise ValueError(f"Unknown loss type `{loss_type}`")
