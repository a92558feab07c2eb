"""CPU oracle of the whole gnn-mlp detector step — plain PyTorch, per-edge formulation, autograd backward.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates, with the reference's class / attribute names so that ``state_dict`` keys coincide with the product's:
  * ptgnn ``GraphNeuralNetwork.forward`` (parity unpinned) over the gnn-mlp layer list of
    buglab/models/gnnlayerdefs.py:26-39;
  * ``LocalizationModule`` — buglab/models/layers/localizationmodule.py:54-124;
  * the three repair heads — buglab/models/layers/fixermodules.py:31-39,65-73,110-124 and layers/mlp.py:6-20
    (with ``_input_dim`` defined: the reference's F9 bug would raise on every ArgSwap candidate);
  * ``GnnBugLabModule.forward`` / ``_compute_repair_logprobs`` — buglab/models/gnn.py:144-322 (discriminator branch);
  * optimiser step — Adam(1e-4) + clip_grad_norm_(0.5) + linear warm-up (utils.py:51-66, train.py:104).
The in-repo pieces are pinned against the real reference code by tests/golden (make_golden.py).
"""
import math
from typing import Any, Dict, List, Optional

import torch
from torch import nn

from .mp_ref import MlpMessagePassingLayer, SubtokenUnitEmbedder
from .scatter_ref import scatter_log_softmax, scatter_max, scatter_min, scatter_sum


def compute_generator_loss_ref(arg_swap_logprobs, arrange, candidate_rewrite_idxs, candidate_symbol_to_location_group,
                               localization_logprobs, loss_type, pair_rewrite_idxs, rewrite_logprobs, rewrite_to_graph_id,
                               rewrite_to_location_group, swapped_pair_to_call_location_group, text_repair_logprobs,
                               text_rewrite_idxs, varmisuse_logprobs):
    """Follows buglab/models/utils.py:101-179 step by step (in-place accumulation into a zero vector, masked_select)."""
    gen = torch.zeros_like(rewrite_logprobs)
    num_graphs = arrange.shape[0]
    gen[-num_graphs:] = localization_logprobs[-num_graphs:]                                   # :118-120
    gen[text_rewrite_idxs] += localization_logprobs[rewrite_to_location_group] + text_repair_logprobs          # :122-124
    gen[candidate_rewrite_idxs] += localization_logprobs[candidate_symbol_to_location_group] + varmisuse_logprobs  # :126-128
    gen[pair_rewrite_idxs] += localization_logprobs[swapped_pair_to_call_location_group] + arg_swap_logprobs    # :130-132
    observed = torch.isinf(rewrite_logprobs).logical_not()                                  # :135
    index = torch.cat((rewrite_to_graph_id, arrange)).masked_select(observed)
    det = rewrite_logprobs.masked_select(observed)
    g = gen.masked_select(observed)
    if loss_type in ("norm-kl", "norm-rmse", "classify-max-loss"):                           # :139-168
        g = scatter_log_softmax(g, index)
        if loss_type == "norm-rmse":
            return (torch.logaddexp(scatter_log_softmax(det, index), g) ** 2).mean()
        if loss_type == "norm-kl":
            failed = torch.log(torch.max(1.0 - det.exp(), torch.full_like(det, 1e-30)))
            kl = failed.exp() * (scatter_log_softmax(failed, index) - g)
            return scatter_sum(kl, index).mean()
        _, min_idx = scatter_min(det, index)
        return -g[min_idx].mean()
    if loss_type == "expectation":                                                           # :172-175
        return scatter_sum(g.exp() * det, index).mean()
    raise ValueError(loss_type)


class _NoParams(nn.Module):
    """Placeholder for the parameter-free dummy / concat-residual entries (keeps ModuleList indices aligned)."""

    def __init__(self, kind: str):
        super().__init__()
        self.kind = kind


class GraphNeuralNetwork(nn.Module):
    def __init__(self, hidden: int, num_edge_types: int, vocabulary_size: int, dropout_rate: float = 0.0,
                 embedding_dropout_rate: float = 0.0, use_message_bias: bool = True):
        super().__init__()

        def mp(f):
            return MlpMessagePassingLayer(f * hidden, f * hidden, hidden, num_edge_types, "max", dropout_rate,
                                          use_message_bias=use_message_bias)

        block = lambda: [_NoParams("remember"), mp(1), mp(1), mp(1), _NoParams("concat"), mp(2)]  # noqa: E731
        self.__message_passing_layers = nn.ModuleList(block() + block())  # gnnlayerdefs.py:26-39
        self.__node_embedder = SubtokenUnitEmbedder(vocabulary_size, hidden, embedding_dropout_rate)

    @property
    def layers(self):
        return self.__message_passing_layers

    def force_winners(self, per_layer_winners) -> None:
        """Route every max-aggregation like the traced implementation did (None resets to the oracle's own argmax)."""
        mp_layers = [l for l in self.__message_passing_layers if not isinstance(l, _NoParams)]
        if per_layer_winners is None:
            per_layer_winners = [None] * len(mp_layers)
        assert len(per_layer_winners) == len(mp_layers)
        for layer, winners in zip(mp_layers, per_layer_winners):
            layer.forced_winners = winners

    def routing_audits(self):
        """Per message-passing layer: the audit of the last forced routing (see mp_ref.MlpMessagePassingLayer)."""
        return [l.routing_audit for l in self.__message_passing_layers if not isinstance(l, _NoParams)]

    def forward(self, node_data, adjacency_lists, return_all_states: bool = False):
        state = self.__node_embedder(**node_data)
        states, remembered = [state], None
        for layer in self.__message_passing_layers:
            if isinstance(layer, _NoParams):
                if layer.kind == "remember":
                    remembered = state
                else:
                    state = torch.cat((remembered, state), dim=-1)
            else:
                state = layer(state, adjacency_lists)
            states.append(state)
        return torch.cat(states, dim=-1) if return_all_states else state


class AuditedReLU(nn.Module):
    """ReLU whose kink decisions (pre-activation > 0) can be forced to another implementation's — the third kind of
    non-differentiable point on the path besides the two arg-routed maxima.  A pre-activation within rounding distance of
    zero may legitimately fall on either side; the forward value is unaffected (|x| tiny) but the unit's whole gradient
    switches on or off: ONE such unit among the argswap scorer's 256 moved 0.5 % of every gradient tensor in smoke().
    ``forced_masks``: list of bool tensors consumed in call order; the audit records how far from zero the forced
    decisions that disagree with this evaluation's own sign are (a wrong mask misses by O(1))."""

    def __init__(self):
        super().__init__()
        self.forced_masks = None
        self.routing_audit = None
        self._calls = 0

    def forward(self, x):
        if self.forced_masks is None:
            return torch.relu(x)
        mask = self.forced_masks[self._calls]
        self._calls += 1
        with torch.no_grad():
            differing = mask != (x > 0)
            worst = float(x[differing].abs().max()) if bool(differing.any()) else 0.0
            prev = self.routing_audit or dict(decisions=0, differing=0, wrong_segment=0, empty_mismatch=0,
                                              max_relative_deficit=0.0, later_edge_on_exact_tie=0)
            self.routing_audit = dict(prev, decisions=prev["decisions"] + int(mask.numel()),
                                      differing=prev["differing"] + int(differing.sum()),
                                      max_relative_deficit=max(prev["max_relative_deficit"], worst))
        return torch.where(mask, x, torch.zeros_like(x))


class MLP(nn.Module):
    def __init__(self, input_dim: int, out_dim: int, hidden_layer_dims: List[int]):
        super().__init__()
        layers, d = [], input_dim
        for h in hidden_layer_dims:
            layers += [nn.Linear(d, h), AuditedReLU()]
            d = h
        layers.append(nn.Linear(d, out_dim))
        self._layers = nn.Sequential(*layers)

    def forward(self, x):
        return self._layers(x)


class LocalizationModule(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self._summary_repr = nn.Linear(dim, dim)
        self._l1 = nn.Linear(2 * dim, dim)
        self._repr_to_localization_score = nn.Linear(dim, 1, bias=False)
        # Like the message-passing layers (mp_ref.MlpMessagePassingLayer.forced_winners): route the per-sample max over the
        # candidates like another implementation did ([S, dim] candidate positions; C = empty) and audit that routing
        # against this evaluation's own maximum.  Two candidates within rounding distance of each other may legitimately
        # swap; one swap moves ~1/(S*dim) of the gradient mass, far above the elementwise gradient tolerance.
        self.forced_summary_args = None
        self.routing_audit = None

    def _summary(self, candidate_reprs, candidate_to_sample_idx):
        projected = self._summary_repr(candidate_reprs)
        own_max, own_arg = scatter_max(projected, candidate_to_sample_idx, dim=0)
        if self.forced_summary_args is None:
            return own_max
        arg, C = self.forced_summary_args, projected.shape[0]
        picked = projected.gather(0, arg.clamp(max=C - 1))
        forced = torch.where(arg >= C, torch.zeros_like(picked), picked)
        with torch.no_grad():
            valid = arg < C
            sample_of_winner = candidate_to_sample_idx[arg.clamp(max=C - 1)]
            wrong_segment = valid & (sample_of_winner != torch.arange(arg.shape[0]).view(-1, 1))
            deficit = (own_max - forced) / (1.0 + own_max.abs())
            self.routing_audit = dict(
                decisions=int(arg.numel()), differing=int((arg != own_arg).sum()), wrong_segment=int(wrong_segment.sum()),
                empty_mismatch=int(((arg >= C) != (own_arg >= C)).sum()),
                max_relative_deficit=float(deficit.max()) if deficit.numel() else 0.0,
                later_edge_on_exact_tie=int((valid & (own_arg < C) & (arg > own_arg) & (forced == own_max)).sum()))
        return forced

    def compute_localization_logprobs(self, candidate_reprs, candidate_to_sample_idx, num_samples):  # :54-79
        summary = self._summary(candidate_reprs, candidate_to_sample_idx)[candidate_to_sample_idx]
        l1 = torch.sigmoid(self._l1(torch.cat([candidate_reprs, summary], dim=-1)))
        scores = self._repr_to_localization_score(l1).squeeze(-1)
        arange = torch.arange(num_samples, dtype=torch.int64)
        scores = torch.cat((scores, torch.ones(num_samples, dtype=scores.dtype)))
        groups = torch.cat((candidate_to_sample_idx, arange))
        return groups, scatter_log_softmax(scores, groups), arange

    def forward(self, candidate_reprs, candidate_to_sample_idx, has_bug, correct_candidate_idxs, weight: float = 1.0):  # :81-124
        groups, log_probs, arange = self.compute_localization_logprobs(candidate_reprs, candidate_to_sample_idx, has_bug.shape[0])
        correct = torch.where(has_bug, correct_candidate_idxs, arange + candidate_reprs.shape[0])
        per_sample = log_probs[correct].clamp(min=-math.inf, max=math.log(0.995))
        if weight == 1.0:
            return -per_sample.mean(), log_probs, groups
        w = torch.where(has_bug, torch.full_like(per_sample, weight), torch.ones_like(per_sample))
        return -(per_sample * w).sum() / w.sum(), log_probs, groups


class TextRepairModule(nn.Module):
    def __init__(self, dim: int, rewrite_vocab_size: int):
        super().__init__()
        self.__text_rewrite_embeddings = nn.Embedding(rewrite_vocab_size, dim)
        self.__text_rewrite_scorer = MLP(2 * dim, 1, [dim])

    def compute_rewrite_logits(self, node_reprs, candidate_rewrites):  # fixermodules.py:31-39
        emb = self.__text_rewrite_embeddings(candidate_rewrites)
        return self.__text_rewrite_scorer(torch.cat((emb, node_reprs), dim=-1)).squeeze(-1)


class SingleCandidateNodeSelectorModule(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self.__candidate_scorer = MLP(2 * dim, 1, [dim])

    def compute_per_slot_log_probability(self, slot_reprs, target_reprs):  # fixermodules.py:65-73
        return self.__candidate_scorer(torch.cat((slot_reprs, target_reprs), dim=-1)).squeeze(-1)


class CandidatePairSelectorModule(nn.Module):
    def __init__(self, dim: int):
        super().__init__()
        self._input_dim = dim
        self.__pair_scorer = MLP(3 * dim, 1, [dim])

    def compute_per_pair_logits(self, slot_reprs, pair_reprs):  # fixermodules.py:110-124
        return self.__pair_scorer(torch.cat((slot_reprs, pair_reprs.reshape(pair_reprs.shape[0], 2 * self._input_dim)), dim=-1)).squeeze(-1)


class GnnBugLabModule(nn.Module):
    """Discriminator training step of buglab/models/gnn.py:144-251 (selector branch :189-219 not restated)."""

    def __init__(self, hidden: int, num_edge_types: int, vocabulary_size: int, rewrite_vocabulary_size: int,
                 dropout_rate: float = 0.0, embedding_dropout_rate: float = 0.0, buggy_samples_weight: float = 1.0,
                 use_message_bias: bool = True, generator_loss_type: str = "classify-max-loss",
                 use_all_gnn_layer_outputs: bool = False):
        super().__init__()
        self._gnn = GraphNeuralNetwork(hidden, num_edge_types, vocabulary_size, dropout_rate, embedding_dropout_rate,
                                       use_message_bias)
        self.use_all_gnn_layer_outputs = use_all_gnn_layer_outputs
        if use_all_gnn_layer_outputs:  # gnn.py:65-69: Linear over [embedding ; every layer's output] -> output width
            widths, remembered, cur = [hidden], None, hidden
            for layer in self._gnn.layers:
                if isinstance(layer, _NoParams):
                    if layer.kind == "remember":
                        remembered = cur
                    else:
                        cur = remembered + cur
                else:
                    cur = layer.output_state_dimension
                widths.append(cur)
            self.__summarization_layer = nn.Linear(sum(widths), cur)
        self.__localization_module = LocalizationModule(hidden)
        self._text_repair_module = TextRepairModule(hidden, rewrite_vocabulary_size)
        self._varmisuse_module = SingleCandidateNodeSelectorModule(hidden)
        self._argswap_module = CandidatePairSelectorModule(hidden)
        self.buggy_samples_weight = buggy_samples_weight
        self.generator_loss_type = generator_loss_type

    def node_representations(self, graph_data):
        if self.use_all_gnn_layer_outputs:  # gnn.py:109-114
            return self.__summarization_layer(self._gnn(graph_data["node_data"], graph_data["adjacency_lists"], return_all_states=True))
        return self._gnn(graph_data["node_data"], graph_data["adjacency_lists"])

    def force_routing(self, mp_winners, head_args, relu_masks=None) -> None:
        """Evaluate with another implementation's discrete decisions: ``mp_winners`` = its per-layer winning edges
        (``ops.WINNER_TRACE``), ``head_args`` = the args of its differentiable segment maxima in call order
        (``ops.MINMAX_TRACE``; the discriminator step has exactly one: the localisation module's candidate summary),
        ``relu_masks`` = {name of the Linear feeding a ReLU: [pre-activation > 0 per call]} (``parity.relu_trace``).
        ``None`` restores this module's own decisions."""
        self._gnn.force_winners(mp_winners)
        if head_args is not None:
            assert len(head_args) == 1, f"expected one traced head maximum, got {len(head_args)}"
        self.__localization_module.forced_summary_args = head_args[0] if head_args is not None else None
        modules = dict(self.named_modules())
        for name, module in modules.items():
            if isinstance(module, AuditedReLU):
                container, _, index = name.rpartition(".")
                linear = f"{container}.{int(index) - 1}"       # MLP._layers: Linear at i - 1 feeds the ReLU at i
                module.forced_masks = list(relu_masks[linear]) if relu_masks is not None and linear in relu_masks else None
                module.routing_audit, module._calls = None, 0

    def routing_audits(self):
        """Audits of the last forced evaluation: the message-passing layers', the localisation summary's, the ReLU kinks'
        (max_relative_deficit = the largest |pre-activation| whose forced side differs from this evaluation's sign)."""
        audits = list(self._gnn.routing_audits())
        if self.__localization_module.forced_summary_args is not None:
            audits.append(self.__localization_module.routing_audit)
        for module in self.modules():
            if isinstance(module, AuditedReLU) and module.forced_masks is not None and module.routing_audit is not None:
                audits.append(module.routing_audit)
        return audits

    def compute_localization_logprobs(self, graph_data):  # gnn.py:125-142
        states = self.node_representations(graph_data)
        cand = states[graph_data["reference_node_ids"]["candidate_nodes"]]
        groups, lp, arange = self.__localization_module.compute_localization_logprobs(
            cand, graph_data["reference_node_graph_idx"]["candidate_nodes"], graph_data["num_graphs"])
        return groups, lp, states, arange

    def _compute_repair_logprobs(self, states, refs, target_rewrites, rewrite_to_location_group,
                                 candidate_symbol_to_location_group, swapped_pair_to_call_location_group):  # gnn.py:253-322
        text = (self._text_repair_module.compute_rewrite_logits(states[refs["target_rewrite_nodes"]], target_rewrites)
                if target_rewrites.shape[0] > 0 else torch.zeros(0, dtype=states.dtype))
        misuse = (self._varmisuse_module.compute_per_slot_log_probability(
            states[refs["varmisused_node_ids"]], states[refs["candidate_symbol_node_ids"]])
            if refs["varmisused_node_ids"].shape[0] > 0 else torch.zeros(0, dtype=states.dtype))
        swap = (self._argswap_module.compute_per_pair_logits(
            states[refs["call_node_ids"]], states[refs["candidate_swapped_node_ids"]])
            if refs["call_node_ids"].shape[0] > 0 else torch.zeros(0, dtype=states.dtype))
        sizes = [text.shape[0], misuse.shape[0], swap.shape[0]]
        all_logits = torch.cat((text, misuse, swap))
        groups = torch.cat((rewrite_to_location_group, candidate_symbol_to_location_group, swapped_pair_to_call_location_group))
        if all_logits.shape[0] == 0:
            e = torch.zeros(0, dtype=torch.bool)
            return swap, text, misuse, (e, e, e)
        text_lp, misuse_lp, swap_lp = torch.split(scatter_log_softmax(all_logits, groups), sizes)
        with torch.no_grad():
            per_rewrite_max = scatter_max(all_logits, groups)[0].gather(-1, groups)
            text_sel, misuse_sel, swap_sel = torch.split(per_rewrite_max == all_logits, sizes)
        return swap_lp, text_lp, misuse_lp, (swap_sel, text_sel, misuse_sel)

    def forward(self, *, graph_data, correct_candidate_node_idxs, has_bug, target_rewrites, rewrite_to_location_group,
                correct_rewrite_idxs, text_rewrite_idxs=None, candidate_symbol_to_location_group=None,
                correct_candidate_symbols=None, candidate_rewrite_idxs=None, swapped_pair_to_call_location_group=None,
                correct_swapped_pair=None, pair_rewrite_idxs=None, rewrite_to_graph_id=None, rewrite_logprobs=None,
                return_details: bool = False, **_):
        states = self.node_representations(graph_data)
        refs = graph_data["reference_node_ids"]
        cand = states[refs["candidate_nodes"]]
        swap_lp, text_lp, misuse_lp, selected = self._compute_repair_logprobs(
            states, refs, target_rewrites, rewrite_to_location_group, candidate_symbol_to_location_group,
            swapped_pair_to_call_location_group)
        if rewrite_logprobs is not None:  # selector branch, gnn.py:189-219
            _, loc_lp, arange = self.__localization_module.compute_localization_logprobs(
                cand, graph_data["reference_node_graph_idx"]["candidate_nodes"], has_bug.shape[0])
            return compute_generator_loss_ref(
                swap_lp, arange, candidate_rewrite_idxs, candidate_symbol_to_location_group, loc_lp, self.generator_loss_type,
                pair_rewrite_idxs, rewrite_logprobs.to(loc_lp.dtype), rewrite_to_graph_id, rewrite_to_location_group,
                swapped_pair_to_call_location_group, text_lp, text_rewrite_idxs, misuse_lp)
        loc_loss, loc_lp, loc_groups = self.__localization_module(
            cand, graph_data["reference_node_graph_idx"]["candidate_nodes"], has_bug, correct_candidate_node_idxs,
            self.buggy_samples_weight)
        repair = (-text_lp[correct_rewrite_idxs]).sum() + (-misuse_lp[correct_candidate_symbols]).sum() \
            + (-swap_lp[correct_swapped_pair]).sum()
        loss = loc_loss + repair * self.buggy_samples_weight / has_bug.shape[0]  # gnn.py:240-251
        if return_details:
            return loss, dict(node_states=states, localization_logprobs=loc_lp, localization_groups=loc_groups,
                              text_logprobs=text_lp, varmisuse_logprobs=misuse_lp, argswap_logprobs=swap_lp)
        return loss


def minibatch_to_cpu(mb: Dict[str, Any], dtype: Optional[torch.dtype] = None) -> Dict[str, Any]:
    """Product minibatch (possibly on a GPU) -> plain CPU tensors / lists the oracle consumes."""
    def conv(v):
        if isinstance(v, torch.Tensor):
            return v.detach().cpu()
        if isinstance(v, dict):
            return {k: conv(x) for k, x in v.items()}
        if isinstance(v, (list, tuple)) and v and isinstance(v[0], (tuple, list)) and len(v[0]) == 2 \
                and isinstance(v[0][0], torch.Tensor):
            return [(a.detach().cpu().long(), b.detach().cpu().long()) for a, b in v]
        return v

    out = {k: conv(v) for k, v in mb.items()}
    out["graph_data"] = {k: v for k, v in out["graph_data"].items() if k != "h2d_bytes"}
    return out


def train_step_ref(module: nn.Module, optimizer: torch.optim.Optimizer, minibatch: Dict[str, Any],
                   clip_gradient_norm: float = 0.5) -> float:
    """zero_grad -> loss -> backward -> clip_grad_norm_ -> Adam step (SURVEY.md §8a P6 order; train.py:98-107)."""
    optimizer.zero_grad()
    loss = module(**minibatch)
    loss.backward()
    torch.nn.utils.clip_grad_norm_(module.parameters(), clip_gradient_norm)
    optimizer.step()
    return float(loss.detach())
