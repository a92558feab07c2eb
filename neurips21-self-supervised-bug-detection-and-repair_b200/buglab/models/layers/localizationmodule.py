"""Bug localisation head: per-graph softmax over candidate nodes plus a virtual NO_BUG slot
(reference: buglab/models/layers/localizationmodule.py:11-124; attribute names kept for state_dict parity).

The two segment reductions (per-graph max summary, grouped log-softmax) run in the buglab_b200 segment kernels;
metrics are accumulated ON DEVICE and read back only when ``_module_metrics`` is called (the reference syncs
four times per step, localizationmodule.py:103-113)."""
import math
from typing import Any, Callable, Dict

import torch
import torch.nn as nn
from ptgnn.baseneuralmodel import ModuleWithMetrics

from buglab.models.utils import scatter_log_softmax, scatter_max


class LocalizationModule(ModuleWithMetrics):
    def __init__(self, representation_size: int, buggy_samples_weight_schedule: Callable[[int], float],
                 abstain_weight: float = 0.0):
        super().__init__()
        self._summary_repr = nn.Linear(representation_size, representation_size, bias=True)
        self._l1 = nn.Linear(2 * representation_size, representation_size, bias=True)
        self._repr_to_localization_score = nn.Linear(representation_size, 1, bias=False)
        self._buggy_samples_weight_schedule = buggy_samples_weight_schedule
        self._abstain_weight = abstain_weight

    def _reset_module_metrics(self) -> None:
        if not hasattr(self, "_epoch_idx"):
            self._epoch_idx = 0
        elif self.training and self.__num_steps > 0:
            self._epoch_idx += 1  # metrics are reset once per epoch
        # [num_correct, localization_loss, num_no_bug, num_no_bug_correct, total_samples]
        self.__stats = None
        self.__num_steps = 0

    def _module_metrics(self) -> Dict[str, Any]:
        if self.__stats is None:
            return {}
        num_correct, loss, num_no_bug, no_bug_correct, total = self.__stats.tolist()  # the one D2H
        if total == 0:
            return {}
        return {
            "Localization Accuracy": num_correct / total,
            "No Bug Recall": no_bug_correct / num_no_bug if num_no_bug > 0 else float("nan"),
            "Localization Loss": loss / total,
            "Weight of Buggy Samples": self._buggy_samples_weight_schedule(self._epoch_idx),
        }

    def compute_localization_logprobs(self, candidate_reprs, candidate_to_sample_idx, num_samples):
        """candidate_reprs [C, H], candidate_to_sample_idx [C] -> (group ids [C+B], log-probs [C+B], arange(B))."""
        summary_per_sample = scatter_max(self._summary_repr(candidate_reprs), index=candidate_to_sample_idx, dim=0)[0]
        hidden = torch.sigmoid(self._l1(torch.cat((candidate_reprs, summary_per_sample[candidate_to_sample_idx]), dim=-1)))
        candidate_scores = self._repr_to_localization_score(hidden).squeeze(-1)
        arange = torch.arange(num_samples, dtype=torch.int64, device=candidate_to_sample_idx.device)
        scores = torch.cat((candidate_scores, torch.ones(num_samples, dtype=torch.float32, device=arange.device)))
        groups = torch.cat((candidate_to_sample_idx, arange))
        return groups, scatter_log_softmax(scores, groups), arange

    def forward(self, candidate_reprs, candidate_to_sample_idx, has_bug, correct_candidate_idxs):
        num_candidates = candidate_reprs.shape[0]
        groups, log_probs, arange = self.compute_localization_logprobs(candidate_reprs, candidate_to_sample_idx,
                                                                       has_bug.shape[0])
        no_bug_slot = arange + num_candidates
        correct = torch.where(has_bug, correct_candidate_idxs, no_bug_slot)
        per_sample = log_probs[correct].clamp(max=math.log(0.995))
        if getattr(self, "_abstain_weight", 0.0) > 0:
            per_sample = per_sample + torch.where(has_bug, self._abstain_weight * log_probs[no_bug_slot],
                                                  torch.zeros_like(per_sample))
        with torch.no_grad():
            predicted = scatter_max(log_probs, groups, dim_size=has_bug.shape[0])[1]
            is_correct = predicted == correct
            no_bug = has_bug.logical_not()
            n = torch.full((), float(per_sample.shape[0]), device=per_sample.device, dtype=torch.float64)  # no H2D copy
            stats = torch.stack((is_correct.sum().double(), -per_sample.sum().double(), no_bug.sum().double(),
                                 (no_bug & is_correct).sum().double(), n))
            self.__stats = stats if self.__stats is None else self.__stats + stats
            self.__num_steps += 1
        weight = self._buggy_samples_weight_schedule(self._epoch_idx)
        if weight == 1.0:
            return -per_sample.mean()
        weights = torch.where(has_bug, torch.full_like(per_sample, weight), torch.ones_like(per_sample))
        return -(per_sample * weights).sum() / weights.sum()
