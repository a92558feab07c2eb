"""Rewrite-candidate bookkeeping shared by the BugLab models (reference: buglab/models/basemodel.py:13-346).

Pure integer / list work on the host; results must equal the reference's bit for bit (tests/golden).  A sample's
candidate rewrites are split by scout into three families — text rewrites, variable-misuse candidates and
argument swaps — and flattened into (node ids, location-group ids, index of the correct candidate, original
rewrite indices) tables; groups are numbered by the position of the location among the sorted unique
candidate nodes.
"""
from collections import defaultdict
from contextlib import contextmanager
from typing import Any, Dict, Iterator, List, NamedTuple, Optional, Tuple

import numpy as np
from dpu_utils.mlutils import Vocabulary

from buglab.representations.data import BugLabData

# the rewrite vocabulary of the text-repair head (reference basemodel.py:13-64); language specific
_ARITHMETIC = ["+", "-", "*", "/", "**", "//", "%", "@", "<<", ">>", "|", "&", "^"]
_REWRITE_OPERATORS = (
    _ARITHMETIC + [op + "=" for op in _ARITHMETIC] + ["=", "<", "<=", ">", ">=", "==", "!="]
    + [" in ", " not in ", " is ", " is not "] + ["0", "1", "2", "-1", "-2"] + ["and", "or", "not ", "", "True", "False"]
)


class _Family:
    """Candidates of one rewrite family, grouped by location node in first-seen order."""

    def __init__(self):
        self.payload: Dict[int, List[Any]] = defaultdict(list)
        self.original_idx: Dict[int, List[int]] = defaultdict(list)
        self.correct: Optional[Tuple[int, int]] = None  # (location node, position within that location)

    def add(self, location: int, payload: Any, rewrite_idx: int, is_target: bool) -> None:
        if is_target:
            self.correct = (location, len(self.payload[location]))
        self.payload[location].append(payload)
        self.original_idx[location].append(rewrite_idx)

    def flatten(self, group_of_location: Dict[int, int]):
        locations, payloads, groups, originals = [], [], [], []
        correct_idx = None
        for location, items in self.payload.items():
            if self.correct is not None and self.correct[0] == location:
                correct_idx = len(payloads) + self.correct[1]
            locations.extend([location] * len(items))
            payloads.extend(items)
            groups.extend([group_of_location[location]] * len(items))
            originals.extend(self.original_idx[location])
        return locations, payloads, groups, correct_idx, originals


class AbstractBugLabModel:
    OPERATOR_REWRITES = frozenset(_REWRITE_OPERATORS)

    def _init(self):
        self._target_rewrite_ops = Vocabulary.create_vocabulary(
            self.OPERATOR_REWRITES, max_size=len(self.OPERATOR_REWRITES), count_threshold=0, add_unk=False)
        self._tensorize_only_at_target_location_rewrites = True

    @contextmanager
    def _tensorize_all_location_rewrites(self):
        try:
            self._tensorize_only_at_target_location_rewrites = False
            yield
        finally:
            self._tensorize_only_at_target_location_rewrites = True

    def _compute_rewrite_data(self, datapoint: BugLabData, candidate_node_idxs):
        graph = datapoint["graph"]
        # positional arguments of every Call node, in Child-edge order (needed by the arg-swap family)
        nodes = graph["nodes"]
        call_args: Dict[int, List[int]] = defaultdict(list)
        for edge in graph["edges"]["Child"]:
            if len(edge) == 3 and edge[2] == "args" and nodes[edge[0]] == "Call":
                call_args[edge[0]].append(edge[1])
        return self._compute_rewrite_data_from(
            graph["reference_nodes"], datapoint["candidate_rewrites"], datapoint["candidate_rewrite_metadata"],
            datapoint["target_fix_action_idx"], call_args, candidate_node_idxs)

    def _compute_rewrite_data_from(self, reference_nodes, candidate_rewrites, candidate_rewrite_metadata, target_action,
                                   call_args, candidate_node_idxs):
        """The rewrite families of one sample from its already-decoded fields (shared by the dict path above and the
        native shard path, buglab_b200/shards.py)."""
        target_node = None if target_action is None else reference_nodes[target_action]
        only_target = self._tensorize_only_at_target_location_rewrites

        text, varmisuse, argswap = _Family(), _Family(), _Family()
        for i, (location, rewrite, metadata) in enumerate(
                zip(reference_nodes, candidate_rewrites, candidate_rewrite_metadata)):
            if only_target and location != target_node:
                continue  # training only scores rewrites at the target location
            scout, scout_metadata = metadata
            rewrite_data = rewrite[1]
            is_target = target_action == i
            if scout == "VariableMisuseRewriteScout":
                varmisuse.add(location, scout_metadata, i, is_target)
            elif scout == "ArgSwapRewriteScout":
                args = call_args[location]
                argswap.add(location, (args[rewrite_data[0]], args[rewrite_data[1]]), i, is_target)
            else:
                text.add(location, self._target_rewrite_ops.get_id_or_unk(rewrite_data), i, is_target)

        group_of_location = {int(node): i for i, node in enumerate(candidate_node_idxs)}
        t_nodes, t_ops, t_groups, t_correct, t_orig = text.flatten(group_of_location)
        v_nodes, v_cands, v_groups, v_correct, v_orig = varmisuse.flatten(group_of_location)
        a_nodes, a_pairs, a_groups, a_correct, a_orig = argswap.flatten(group_of_location)
        return (
            t_nodes, t_ops, t_groups, t_correct, t_orig,
            v_nodes, v_groups, v_cands, v_correct, v_orig,
            a_nodes, a_pairs, a_correct, a_groups, a_orig,
            group_of_location,
        )

    def _iter_per_sample_results(self, mb_data, candidate_location_sample_idx, candidate_location_log_probs,
                                 arg_swap_logprobs, num_samples, original_datapoints, text_repair_logprobs,
                                 varmisuse_logprobs, node_mappings=None):
        """Splits minibatch-level predictions back into per-sample ``(datapoint, {node: logprob}, [rewrite logprob])``
        (reference basemodel.py:240-346).  Location groups are consumed in sample order, one per unique candidate node."""
        loc_sample = np.asarray(candidate_location_sample_idx)
        loc_logprobs = np.asarray(candidate_location_log_probs)
        order = np.argsort(loc_sample, kind="stable")
        bounds = np.searchsorted(loc_sample[order], np.arange(num_samples + 1))

        def by_group(logprobs, groups):
            table: Dict[int, List[float]] = defaultdict(list)
            for group, lp in zip(groups.cpu().numpy().tolist(), logprobs.cpu().numpy().tolist()):
                table[group].append(lp)
            return table

        swap_by_group = by_group(arg_swap_logprobs, mb_data["swapped_pair_to_call_location_group"])
        text_by_group = by_group(text_repair_logprobs, mb_data["rewrite_to_location_group"])
        misuse_by_group = by_group(varmisuse_logprobs, mb_data["candidate_symbol_to_location_group"])

        next_group = 0
        for sample_idx in range(num_samples):
            point = original_datapoints[sample_idx]
            candidate_nodes = np.unique(point["graph"]["reference_nodes"])
            sample_logprobs = loc_logprobs[order[bounds[sample_idx]: bounds[sample_idx + 1]]]
            assert len(sample_logprobs) == len(candidate_nodes) + 1
            keys = candidate_nodes if node_mappings is None else [node_mappings[sample_idx][k] for k in candidate_nodes]
            # sequence models: several graph nodes can share one token position; like the reference's dict
            # comprehension, a repeated position keeps the LAST of its log-probabilities (basemodel.py:291-294)
            location_logprobs = {int(n): lp for n, lp in zip(keys, sample_logprobs)}
            location_logprobs[-1] = sample_logprobs[-1]  # the NO_BUG slot comes last within a sample

            flat_swap, flat_text, flat_misuse = [], [], []
            for _ in range(len(candidate_nodes)):
                flat_swap.extend(swap_by_group[next_group])
                flat_text.extend(text_by_group[next_group])
                flat_misuse.extend(misuse_by_group[next_group])
                next_group += 1
            rewrite_probs: List[Optional[float]] = [None] * len(point["candidate_rewrites"])
            for idxs, lps in ((mb_data["text_rewrite_original_idxs"][sample_idx], flat_text),
                              (mb_data["candidate_rewrite_original_idxs"][sample_idx], flat_misuse),
                              (mb_data["pair_rewrite_original_idx"][sample_idx], flat_swap)):
                assert len(idxs) == len(lps)
                for i, lp in zip(idxs, lps):
                    assert rewrite_probs[i] is None
                    rewrite_probs[i] = lp
            assert None not in rewrite_probs
            if node_mappings is not None:
                # report per GRAPH node again: every reference node that was projected onto a position gets its log-prob
                nodes_at: Dict[int, List[int]] = defaultdict(list)
                reference_nodes = point["graph"]["reference_nodes"]
                for graph_node, position in node_mappings[sample_idx].items():
                    if graph_node in reference_nodes:
                        nodes_at[position].append(graph_node)
                per_node: Dict[int, float] = {}
                for position, lp in location_logprobs.items():
                    for node in (nodes_at[position] if position >= 0 else [position]):
                        per_node[node] = lp
                location_logprobs = per_node
            yield point, location_logprobs, rewrite_probs
