"""Repair heads: text rewrite, variable-misuse candidate, argument-swap pair scorers
(reference: buglab/models/layers/fixermodules.py:9-147; attribute names kept for state_dict parity).

Deviation (documented): the reference's ``CandidatePairSelectorModule`` reads ``self._input_dim``
which it never assigns (fixermodules.py:120) and so raises on any ArgSwap candidate; here the attribute is set.
Accuracy counters live on the device until ``_module_metrics`` is read."""
from typing import Any, Dict

import torch
import torch.nn as nn
from ptgnn.baseneuralmodel import ModuleWithMetrics

from buglab.models.layers.mlp import MLP


class _FixerMetrics(ModuleWithMetrics):
    _metric_name = "Fixer"
    _stats_name = "Fixer Stats"

    def _reset_module_metrics(self) -> None:
        self._counts = None  # device tensor [num_correct, num_samples]

    def _count(self, selected_fixes, targets) -> None:
        if selected_fixes is None:
            return
        with torch.no_grad():
            # torch.full, not torch.tensor: a device tensor made from a host scalar is a synchronising H2D copy
            n = torch.full((), targets.shape[0], device=targets.device, dtype=torch.float64)
            c = torch.stack((selected_fixes[targets].sum().double(), n))
            self._counts = c if self._counts is None else self._counts + c

    def _module_metrics(self) -> Dict[str, Any]:
        if self._counts is None:
            return {}
        correct, samples = (int(v) for v in self._counts.tolist())
        if samples == 0:
            return {}
        return {self._metric_name: correct / samples,
                self._stats_name: f"{correct / samples:.2%} ({correct}/{samples})"}


class TextRepairModule(_FixerMetrics):
    _metric_name = "Text Repair Fixer Accuracy"
    _stats_name = "Text Repair Fixer Stats"

    def __init__(self, input_representation_size: int, rewrite_vocab_size: int):
        super().__init__()
        self.__text_rewrite_embeddings = nn.Embedding(rewrite_vocab_size, embedding_dim=input_representation_size)
        self.__text_rewrite_scorer = MLP(2 * input_representation_size, 1, [input_representation_size])

    def compute_rewrite_logits(self, target_rewrite_node_representations, candidate_rewrites):
        """[N, D] node states and [N] rewrite-op ids -> [N] logits."""
        embedded = self.__text_rewrite_embeddings(candidate_rewrites)
        return self.__text_rewrite_scorer(torch.cat((embedded, target_rewrite_node_representations), dim=-1)).squeeze(-1)

    def forward(self, rewrite_logprobs, targets, selected_fixes=None):
        self._count(selected_fixes, targets)
        return -rewrite_logprobs[targets]


class SingleCandidateNodeSelectorModule(_FixerMetrics):
    _metric_name = "VarMisuse Repair Fixer Accuracy"
    _stats_name = "VarMisuse Repair Fixer Stats"

    def __init__(self, input_representation_size: int):
        super().__init__()
        self.__candidate_scorer = MLP(2 * input_representation_size, 1, [input_representation_size])

    def compute_per_slot_log_probability(self, slot_representations_per_target, target_nodes_representations):
        return self.__candidate_scorer(
            torch.cat((slot_representations_per_target, target_nodes_representations), dim=-1)).squeeze(-1)

    def forward(self, per_slot_logprobs, correct_symbol_node_idxs, selected_fixes=None):
        self._count(selected_fixes, correct_symbol_node_idxs)
        return -per_slot_logprobs[correct_symbol_node_idxs]


class CandidatePairSelectorModule(_FixerMetrics):
    _metric_name = "ArgSwap Repair Fixer Accuracy"
    _stats_name = "ArgSwap Repair Fixes Stats"

    def __init__(self, input_node_representation: int):
        super().__init__()
        self._input_dim = input_node_representation
        self.__pair_scorer = MLP(3 * input_node_representation, 1, [input_node_representation])

    def compute_per_pair_logits(self, slot_representations_per_pair, pair_representations):
        """[N, D] call-node states and [N, 2, D] argument-pair states -> [N] logits."""
        pairs = pair_representations.reshape(pair_representations.shape[0], 2 * self._input_dim)
        return self.__pair_scorer(torch.cat((slot_representations_per_pair, pairs), dim=-1)).squeeze(-1)

    def forward(self, per_slot_logprobs, correct_pair_idx, selected_fixes=None):
        self._count(selected_fixes, correct_pair_idx)
        return -per_slot_logprobs[correct_pair_idx]
