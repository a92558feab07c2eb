"""TEST INFRASTRUCTURE — CPU oracle of the sequence model's module (``SeqBugLabModule``; seq-great / seq-rat), the arithmetic
of SURVEY.md §8(f) row 2 end to end: subtoken embedding -> + positional table -> LayerNorm -> relational transformer
layers -> localisation / repair heads -> loss.  Restates reference buglab/models/seqmodel.py:65-396 on top of
oracle/seq_ref.py (encoder layers), oracle/mp_ref.py (subtoken max-pool embedder) and oracle/model_ref.py (the heads, which
the sequence and graph models share).  Pinned: tests/golden/seq_model.npz holds loss, representations, log-probabilities and
gradients of the real reference module (tests/golden/make_seq_model_golden.py); tests/test_seq_golden.py compares.

The class carries the reference's name so that its name-mangled parameters get the same state_dict keys."""
from typing import Optional

import torch
from torch import nn

from .model_ref import (CandidatePairSelectorModule, LocalizationModule, SingleCandidateNodeSelectorModule, TextRepairModule,
                        compute_generator_loss_ref)
from .mp_ref import SubtokenUnitEmbedder
from .scatter_ref import scatter_log_softmax
from .seq_ref import RelationalTransformerEncoderLayer


class SeqBugLabModule(nn.Module):
    def __init__(self, vocabulary_size: int, embedding_dim: int, num_edge_types: int, num_layers: int, num_heads: int,
                 intermediate_dimension: int, rewrite_vocabulary_size: int, layer_type: str = "great",
                 dropout_rate: float = 0.0, buggy_samples_weight: float = 1.0, generator_loss_type: str = "norm-kl",
                 rezero_mode: str = "off", normalisation_mode: str = "postnorm", positional_rows: int = 5000):
        super().__init__()
        if layer_type not in ("great", "rat"):
            raise NotImplementedError("the oracle covers the relational layer types; 'transformer' and 'gru' are torch's own")
        self.__token_embedder = SubtokenUnitEmbedder(vocabulary_size, embedding_dim, dropout_rate)
        self.__positional_encoding = nn.Parameter(torch.randn(1, positional_rows, embedding_dim))
        self.__input_layer_norm = nn.LayerNorm(embedding_dim)
        self.__seq_layers = nn.ModuleList([
            RelationalTransformerEncoderLayer(
                d_model=embedding_dim, key_query_dimension=embedding_dim // num_heads,
                value_dimension=embedding_dim // num_heads, nhead=num_heads, num_edge_types=num_edge_types,
                dim_feedforward=intermediate_dimension, dropout=dropout_rate,
                use_edge_value_biases=(layer_type == "rat"),      # the scalar-bias switch is never passed (seqmodel.py:93-107)
                rezero_mode=rezero_mode, normalisation_mode=normalisation_mode)
            for _ in range(num_layers)])
        self.__localization_module = LocalizationModule(embedding_dim)
        self._text_repair_module = TextRepairModule(embedding_dim, rewrite_vocabulary_size)
        self._varmisuse_module = SingleCandidateNodeSelectorModule(embedding_dim)
        self._argswap_module = CandidatePairSelectorModule(embedding_dim)
        self._dropout_rate = dropout_rate
        self.buggy_samples_weight, self.generator_loss_type = buggy_samples_weight, generator_loss_type

    # seqmodel.py:351-396
    def compute_output_representation(self, input_sequence_ids, input_seq_num_subtokens, token_sequence_lengths, edges,
                                      edge_types):
        B, L, T = input_sequence_ids.shape
        x = self.__token_embedder(input_sequence_ids.reshape(B * L, T), input_seq_num_subtokens.reshape(B * L)).view(B, L, -1)
        is_token = torch.arange(L)[None, :] < token_sequence_lengths[:, None]          # [B, L]
        x = x + self.__positional_encoding[:, :L]
        x = nn.functional.dropout(self.__input_layer_norm(x), self._dropout_rate, self.training)
        x = x * is_token[..., None]                                                     # padding rows are zeroed once, here
        padding = ~is_token
        for layer in self.__seq_layers:
            x = layer(x, padding, edges, edge_types)
        return x

    def compute_localization_logprobs(self, rep, candidate_location_idxs, num_samples):
        candidates = rep[candidate_location_idxs[:, 0], candidate_location_idxs[:, 1]]
        groups, logprobs, arange = self.__localization_module.compute_localization_logprobs(
            candidates, candidate_location_idxs[:, 0], num_samples)
        return groups, logprobs, arange

    # seqmodel.py:164-225
    def compute_repair_logprobs(self, rep, target_rewrite_node_ids, target_rewrites, rewrite_to_location_group,
                                varmisused_node_ids, candidate_symbol_node_ids, candidate_symbol_to_location_group,
                                call_node_ids, candidate_swapped_node_ids, swapped_pair_to_call_location_group):
        at = lambda idx: rep[idx[:, 0], idx[:, 1]]  # noqa: E731
        text = (self._text_repair_module.compute_rewrite_logits(at(target_rewrite_node_ids), target_rewrites)
                if target_rewrites.shape[0] > 0 else torch.zeros(0, dtype=rep.dtype))
        misuse = (self._varmisuse_module.compute_per_slot_log_probability(at(varmisused_node_ids), at(candidate_symbol_node_ids))
                  if varmisused_node_ids.shape[0] > 0 else torch.zeros(0, dtype=rep.dtype))
        swap = (self._argswap_module.compute_per_pair_logits(
            at(call_node_ids), rep[candidate_swapped_node_ids[:, 0].unsqueeze(-1), candidate_swapped_node_ids[:, 1:]])
            if call_node_ids.shape[0] > 0 else torch.zeros(0, dtype=rep.dtype))
        sizes = [text.shape[0], misuse.shape[0], swap.shape[0]]
        logits = torch.cat((text, misuse, swap))
        groups = torch.cat((rewrite_to_location_group, candidate_symbol_to_location_group, swapped_pair_to_call_location_group))
        text_lp, misuse_lp, swap_lp = torch.split(scatter_log_softmax(logits, groups), sizes)
        return swap_lp, text_lp, misuse_lp

    # seqmodel.py:232-349
    def forward(self, *, input_sequence_ids, input_seq_num_subtokens, token_sequence_lengths, edges, edge_types, has_bug,
                candidate_location_idxs, target_location_idxs, target_rewrite_node_ids, target_rewrites,
                rewrite_to_location_group, correct_rewrite_idxs, text_rewrite_idxs, varmisused_node_ids,
                candidate_symbol_node_ids, candidate_symbol_to_location_group, correct_candidate_symbols,
                candidate_rewrite_idxs, call_node_ids, candidate_swapped_node_ids, swapped_pair_to_call_location_group,
                correct_swapped_pair, pair_rewrite_idxs, rewrite_to_graph_id: Optional[torch.Tensor] = None,
                rewrite_logprobs: Optional[torch.Tensor] = None, return_details: bool = False, **_):
        rep = self.compute_output_representation(input_sequence_ids, input_seq_num_subtokens, token_sequence_lengths, edges,
                                                 edge_types)
        swap_lp, text_lp, misuse_lp = self.compute_repair_logprobs(
            rep, target_rewrite_node_ids, target_rewrites, rewrite_to_location_group, varmisused_node_ids,
            candidate_symbol_node_ids, candidate_symbol_to_location_group, call_node_ids, candidate_swapped_node_ids,
            swapped_pair_to_call_location_group)
        candidates = rep[candidate_location_idxs[:, 0], candidate_location_idxs[:, 1]]
        if rewrite_logprobs is not None:  # selector training
            _, loc_lp, arange = self.__localization_module.compute_localization_logprobs(
                candidates, candidate_location_idxs[:, 0], has_bug.shape[0])
            return compute_generator_loss_ref(
                swap_lp, arange, candidate_rewrite_idxs, candidate_symbol_to_location_group, loc_lp, self.generator_loss_type,
                pair_rewrite_idxs, rewrite_logprobs.to(loc_lp.dtype), rewrite_to_graph_id, rewrite_to_location_group,
                swapped_pair_to_call_location_group, text_lp, text_rewrite_idxs, misuse_lp)
        loc_loss, loc_lp, loc_groups = self.__localization_module(candidates, candidate_location_idxs[:, 0], has_bug,
                                                                  target_location_idxs, self.buggy_samples_weight)
        repair = -(text_lp[correct_rewrite_idxs].sum() + misuse_lp[correct_candidate_symbols].sum()
                   + swap_lp[correct_swapped_pair].sum()) * self.buggy_samples_weight
        loss = loc_loss + repair / has_bug.shape[0]
        if return_details:
            return loss, dict(output_representation=rep, localization_logprobs=loc_lp, localization_groups=loc_groups,
                              text_logprobs=text_lp, varmisuse_logprobs=misuse_lp, argswap_logprobs=swap_lp)
        return loss
