"""Host side of the sequence models (``seq-great`` / ``seq-rat`` / ``seq-transformer`` / ``seq-gru``): the projection of a
code graph onto its token sequence, per-sample tensors and minibatch packing — SURVEY.md §8(f) row 2.

Restates reference buglab/models/seqmodel.py:442-624 (graph -> tokens), :626-729 (``tensorize``), :731-975 (minibatch);
integer bookkeeping is bit-exact against the real reference (tests/golden/seq_model.npz, tests/test_seq_golden.py).

The module (``SeqBugLabModule``, reference seqmodel.py:65-396) runs its attention core on the hand-written
``bl_seq_attention_*`` kernels (first correct path; their source is pinned on the CPU through a host emulation, B200 parity
tests are gated until they have run once), LayerNorm / subtoken embedding / segment ops on the kernels the graph model
already uses, dense layers on library GEMMs.  The edge-free baselines ``transformer`` / ``gru`` are torch's own layers,
as they are in the reference.
The arithmetic is pinned in oracle/seq_ref.py + oracle/seq_model_ref.py.
"""
import logging
from collections import defaultdict
from typing import Any, Callable, Dict, Iterator, List, NamedTuple, Optional, Tuple, Union

import numpy as np
import torch
from ptgnn.baseneuralmodel import AbstractNeuralModel, ModuleWithMetrics
from ptgnn.neuralmodels.embeddings.strelementrepresentationmodel import StrElementRepresentationModel
from torch import nn

from buglab.models.basemodel import AbstractBugLabModel
from buglab.models.layers.fixermodules import (CandidatePairSelectorModule, SingleCandidateNodeSelectorModule,
                                                TextRepairModule)
from buglab.models.layers.localizationmodule import LocalizationModule
from buglab.models.layers.relational_transformer import RelationalTransformerEncoderLayer
from buglab.models.utils import compute_generator_loss, scatter_log_softmax, scatter_max
from buglab.representations.data import BugLabData, BugLabGraph

LOGGER = logging.getLogger(__name__)

# relations that describe the token/AST skeleton itself and are not turned into attention edges (seqmodel.py:452-463)
SKELETON_EDGE_KINDS = frozenset({"NextToken", "PossibleType", "CandidateCall", "CandidateCallDoc", "MayFormalName", "Child",
                                 "Sibling", "OccurrenceOf"})
_COMPARISON_TOKENS = frozenset({"<", "<=", "==", "!=", ">", ">=", "is", "in", "not"})
_TWO_TOKEN_COMPARISONS = frozenset({"IsNot", "NotIn"})
_BINARY_OPERATOR_TOKENS = frozenset({"+", "-", "*", "/", "//", "**", "%", "@", ">>", "<<", "|", "&", "^"})


class SeqModelTensorizedSample(NamedTuple):
    target_subtokens_ids: List[np.ndarray]
    intra_token_edges: Dict[str, List[Tuple[int, int]]]
    candidate_location_idxs: np.ndarray
    target_location_idx: Optional[int]
    node_mappings: Dict[int, int]
    # repair
    target_rewrite_node_ids: List[int]
    target_rewrites: List[int]
    target_rewrite_to_location_group: List[int]
    correct_rewrite_target: Optional[int]
    text_rewrite_original_idx: List[int]
    varmisused_node_ids: List[int]
    candidate_symbol_node_ids: List[int]
    candidate_symbol_to_varmisused_node: List[int]
    correct_candidate_symbol_node: Optional[int]
    candidate_rewrite_original_idx: List[int]
    call_node_ids: List[int]
    candidate_swapped_node_ids: List[Tuple[int, int]]
    swapped_pair_to_call: List[int]
    correct_swapped_pair: Optional[int]
    pair_rewrite_original_idx: List[int]
    num_rewrite_locations_considered: int
    # selector training
    rewrite_logprobs: Optional[List[float]]


class TokenProjectionError(Exception):
    """The graph does not have the shape the projection needs; the sample is dropped (seqmodel.py:585-590, 632-638)."""


class _TokenProjection:
    """Maps every node of one graph to a position in its token sequence.

    Tokens map to their own position.  An AST node maps to a representative token below it — the operator token for
    binary operations and comparison targets, the ``=`` token for assignments, otherwise the first token found walking
    down its children — or, failing that, to whatever its parent maps to.  Symbol nodes map to their first occurrence.
    The walk consults the mapping built SO FAR, so the order of the ``Child`` edges matters and is preserved."""

    def __init__(self, graph: BugLabGraph):
        self.labels = graph["nodes"]
        self.tokens = self._token_chain(graph)
        self.position: Dict[int, int] = {node: i for i, node in enumerate(self.tokens)}
        self._token_set = set(self.tokens)
        self.children: Dict[int, List[int]] = defaultdict(list)
        for edge in graph["edges"]["Child"]:
            self.children[edge[0]].append(edge[1])

    # ---- the token chain (seqmodel.py:592-624) -----------------------------------------------------
    @staticmethod
    def _token_chain(graph: BugLabGraph) -> List[int]:
        successor = {a: b for a, b in graph["edges"]["NextToken"]}
        heads = set(successor) - set(successor.values())
        if len(heads) != 1:
            LOGGER.error("Encountered graph where the tokens are not connected in %s", graph["path"])
            raise TokenProjectionError("token chain has %d heads" % len(heads))
        current = next(iter(heads))
        chain, seen = [current], {current}
        while current in successor:
            current = successor[current]
            if current in seen:
                LOGGER.error("Cyclic token sequence in %s", graph["path"])
                raise TokenProjectionError("cyclic token chain")
            seen.add(current)
            chain.append(current)
        if len(chain) != len(successor) + 1:
            LOGGER.error("Broken token sequence in %s", graph["path"])
            raise TokenProjectionError("broken token chain")
        return chain

    # ---- node -> token position (seqmodel.py:470-545) ----------------------------------------------
    def _first_token_below(self, node: int) -> int:
        stack = [node]
        while stack:
            current = stack.pop()
            if current in self._token_set:
                return current
            stack.extend(self.children[current])
        raise TokenProjectionError("no token below node %d" % node)

    def _via_parent(self, node: int) -> int:
        for parent, kids in self.children.items():
            if node in kids:
                return self.locate(parent)
        raise TokenProjectionError("node %d has no parent to fall back to" % node)

    def locate(self, node: int) -> int:
        known = self.position.get(node)
        if known is not None:
            return known
        position, labels, children = self.position, self.labels, self.children
        stack = [node]
        while stack:
            current = stack.pop()
            label = labels[current]
            if label == "ComparisonTarget":
                for kid in children[current]:
                    kid_label = labels[kid]
                    if kid_label in _COMPARISON_TOKENS:
                        return position[kid]
                    if kid_label in _TWO_TOKEN_COMPARISONS:
                        return position[self._first_token_below(kid)]
                raise TokenProjectionError("comparison target without a comparison operator")
            if label == "BinaryOperation":
                for kid in children[current]:
                    if labels[kid] in _BINARY_OPERATOR_TOKENS:
                        if kid not in position:
                            return self._via_parent(node)
                        return position[kid]
                raise TokenProjectionError("binary operation without an operator")
            if label in ("Assign", "AugAssign"):
                for kid in children[current]:
                    if "=" in labels[kid]:
                        return position[kid]
                raise TokenProjectionError("assignment without an equals token")
            for kid in children[current]:
                kid_position = position.get(kid)
                if kid_position is not None:
                    return kid_position
                stack.append(kid)
        return self._via_parent(node)  # rarely needed (e.g. f-strings)


def project_graph_to_tokens(graph: BugLabGraph):
    """``(token labels, node -> token position, {relation: [(from position, to position)]}, reference positions)`` or raises.
    Exceptions other than ``TokenProjectionError`` (KeyError on an unmapped node ...) are the reference's failure modes
    too and are handled by the callers in the same way: the sample is skipped."""
    proj = _TokenProjection(graph)
    position = proj.position
    for edge in graph["edges"]["Child"]:  # every AST node, in edge order (the mapping grows while it is consulted)
        parent, kid = edge[0], edge[1]
        position[parent] = proj.locate(parent)
        position[kid] = proj.locate(kid)

    occurrences: Dict[int, List[int]] = defaultdict(list)
    for token, symbol in graph["edges"]["OccurrenceOf"]:
        occurrences[symbol].append(token)
    for symbol, tokens in occurrences.items():
        position[symbol] = min(position[t] for t in tokens)

    relations: Dict[str, List[Tuple[int, int]]] = {}
    for kind, adjacency in graph["edges"].items():
        if kind not in SKELETON_EDGE_KINDS:
            relations[kind] = [(proj.locate(a), proj.locate(b)) for a, b in adjacency]

    reference_positions = []
    for node in graph["reference_nodes"]:
        where = proj.locate(node)
        position[node] = where
        reference_positions.append(where)
    return [proj.labels[t] for t in proj.tokens], position, relations, reference_positions



def _const_one(_epoch: int) -> float:
    return 1.0


class SeqBugLabModule(ModuleWithMetrics):
    """Subtoken embedding + positional table -> LayerNorm -> relational transformer layers -> localisation / repair heads."""

    def __init__(self, token_embedder, embedding_dim: int, num_edge_types: int, num_layers: int, num_heads: int,
                 intermediate_dimension: int, dropout_rate: float, rewrite_vocabulary_size: int, layer_type: str = "great",
                 buggy_samples_weight_schedule: Callable[[int], float] = _const_one,
                 generator_loss_type: Optional[str] = "norm-kl", rezero_mode: str = "off",
                 normalisation_mode: str = "postnorm"):
        super().__init__()
        if layer_type not in ("great", "rat", "transformer", "gru"):
            raise ValueError(f"Unrecognized layer type `{layer_type}`.")
        self.__generator_loss_type = generator_loss_type
        self.__token_embedder = token_embedder
        self.__positional_encoding = nn.Parameter(torch.randn(1, 5000, embedding_dim), requires_grad=True)
        self.__dropout_layer = nn.Dropout(dropout_rate)
        self.__input_layer_norm = nn.LayerNorm(embedding_dim)
        self.__layer_type = layer_type
        self.__num_edge_types = num_edge_types
        if layer_type in ("great", "rat"):
            self.__seq_layers = nn.ModuleList([
                RelationalTransformerEncoderLayer(
                    nhead=num_heads, num_edge_types=num_edge_types, d_model=embedding_dim,
                    key_query_dimension=embedding_dim // num_heads, value_dimension=embedding_dim // num_heads,
                    dim_feedforward=intermediate_dimension, dropout=dropout_rate,
                    use_edge_value_biases=layer_type == "rat", rezero_mode=rezero_mode,
                    normalisation_mode=normalisation_mode)
                for _ in range(num_layers)])
        elif layer_type == "transformer":
            # the two edge-free baselines are torch's own layers in the reference too (seqmodel.py:110-131): library code
            self.__seq_layers = nn.ModuleList([
                nn.TransformerEncoderLayer(d_model=embedding_dim, nhead=num_heads, dim_feedforward=intermediate_dimension,
                                           dropout=dropout_rate) for _ in range(num_layers)])
        else:
            self.__seq_layers = nn.GRU(input_size=embedding_dim, hidden_size=embedding_dim // 2, num_layers=num_layers,
                                       bidirectional=True, batch_first=True)
        self.__localization_module = LocalizationModule(embedding_dim, buggy_samples_weight_schedule=buggy_samples_weight_schedule)
        self._buggy_samples_weight_schedule = buggy_samples_weight_schedule
        self._text_repair_module = TextRepairModule(embedding_dim, rewrite_vocabulary_size)
        self._varmisuse_module = SingleCandidateNodeSelectorModule(embedding_dim)
        self._argswap_module = CandidatePairSelectorModule(embedding_dim)

    # ---- metrics (device-resident; one D2H when reported) ---------------------------------------
    def _reset_module_metrics(self) -> None:
        if not hasattr(self, "_epoch_idx"):
            self._epoch_idx = 0
        elif self.training and self.__num_batches > 0:
            self._epoch_idx += 1
        self.__sums = None
        self.__num_batches = 0

    def _module_metrics(self) -> Dict[str, Any]:
        if self.__sums is None:
            return {}
        loss, repair_loss, samples = self.__sums.tolist()
        metrics = {"Loss": loss / self.__num_batches}
        if samples > 0:
            metrics["Repair Loss"] = repair_loss / samples
        return metrics

    def __accumulate(self, loss, repair_loss, samples) -> None:
        with torch.no_grad():
            s = torch.stack((loss.detach().double(), repair_loss.detach().double(), samples.double()))
            self.__sums = s if self.__sums is None else self.__sums + s
            self.__num_batches += 1

    # ---- encoder (seqmodel.py:351-396) -------------------------------------------------------------
    def _compute_output_representation(self, input_sequence_ids, input_seq_num_subtokens, token_sequence_lengths, edges,
                                       edge_types):
        from buglab_b200 import ops

        B, L, T = input_sequence_ids.shape
        x = self.__token_embedder(token_idxs=input_sequence_ids.reshape(B * L, T),
                                  lengths=input_seq_num_subtokens.reshape(B * L)).view(B, L, -1)
        is_token = torch.arange(L, device=x.device)[None, :] < token_sequence_lengths[:, None]
        if self.__layer_type != "gru":  # positions and the BERT-style input LayerNorm are for the attention variants only
            x = x + self.__positional_encoding[:, :L]
            norm = self.__input_layer_norm
            x = self.__dropout_layer(ops.layer_norm(x.reshape(B * L, -1), norm.weight, norm.bias, norm.eps).view(B, L, -1))
        x = x * is_token.unsqueeze(-1)
        padding = ~is_token
        if self.__layer_type in ("great", "rat"):
            plan = ops.build_seq_attention_plan(edges, edge_types, token_sequence_lengths, L, self.__num_edge_types)
            for layer in self.__seq_layers:
                x = layer(src=x, src_mask=padding, edges=plan)
        elif self.__layer_type == "transformer":
            for layer in self.__seq_layers:
                x = layer(x.transpose(0, 1), src_key_padding_mask=padding).transpose(0, 1)
        else:
            packed = nn.utils.rnn.pack_padded_sequence(x, lengths=token_sequence_lengths.cpu(), batch_first=True,
                                                       enforce_sorted=False)
            x, _ = nn.utils.rnn.pad_packed_sequence(self.__seq_layers(packed)[0], batch_first=True)
        return x

    @staticmethod
    def _at(rep: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
        return rep[idx[:, 0], idx[:, 1]]

    def _compute_localization_logprobs(self, candidate_reprs, candidate_to_sample_idx, num_samples):
        groups, logprobs, _ = self.__localization_module.compute_localization_logprobs(
            candidate_reprs, candidate_to_sample_idx, num_samples)
        return groups, logprobs

    # ---- repair heads (seqmodel.py:164-225) --------------------------------------------------------
    def _compute_repair_logprobs(self, output_representations, target_rewrite_node_ids, target_rewrites,
                                 rewrite_to_location_group, varmisused_node_ids, candidate_symbol_node_ids,
                                 candidate_symbol_to_location_group, call_node_ids, candidate_swapped_node_ids,
                                 swapped_pair_to_call_location_group, return_selected: bool = False):
        rep, dev = output_representations, output_representations.device
        text = (self._text_repair_module.compute_rewrite_logits(self._at(rep, target_rewrite_node_ids), target_rewrites)
                if target_rewrites.shape[0] > 0 else torch.zeros(0, device=dev))
        misuse = (self._varmisuse_module.compute_per_slot_log_probability(
            self._at(rep, varmisused_node_ids), self._at(rep, candidate_symbol_node_ids))
            if varmisused_node_ids.shape[0] > 0 else torch.zeros(0, device=dev))
        swap = (self._argswap_module.compute_per_pair_logits(
            self._at(rep, call_node_ids), rep[candidate_swapped_node_ids[:, 0].unsqueeze(-1), candidate_swapped_node_ids[:, 1:]])
            if call_node_ids.shape[0] > 0 else torch.zeros(0, device=dev))
        sizes = [text.shape[0], misuse.shape[0], swap.shape[0]]
        logits = torch.cat((text, misuse, swap))
        groups = torch.cat((rewrite_to_location_group, candidate_symbol_to_location_group, swapped_pair_to_call_location_group))
        if logits.shape[0] == 0:
            empty = torch.zeros(0, dtype=torch.bool, device=dev)
            return (swap, text, misuse, (empty, empty, empty)) if return_selected else (swap, text, misuse)
        text_lp, misuse_lp, swap_lp = torch.split(scatter_log_softmax(logits, index=groups), sizes)
        if not return_selected:
            return swap_lp, text_lp, misuse_lp
        with torch.no_grad():
            group_max = scatter_max(logits, groups)[0]
            text_sel, misuse_sel, swap_sel = torch.split(group_max[groups] == logits, sizes)
        return swap_lp, text_lp, misuse_lp, (swap_sel, text_sel, misuse_sel)

    # ---- loss (seqmodel.py:232-349) ----------------------------------------------------------------
    def forward(self, *, input_sequence_ids, input_seq_num_subtokens, token_sequence_lengths, edges, edge_types, has_bug,
                candidate_location_idxs, target_location_idxs, target_rewrite_node_ids, target_rewrites,
                rewrite_to_location_group, correct_rewrite_idxs, text_rewrite_idxs, varmisused_node_ids,
                candidate_symbol_node_ids, candidate_symbol_to_location_group, correct_candidate_symbols,
                candidate_rewrite_idxs, call_node_ids, candidate_swapped_node_ids, swapped_pair_to_call_location_group,
                correct_swapped_pair, pair_rewrite_idxs, rewrite_to_graph_id: Optional[torch.Tensor] = None,
                rewrite_logprobs: Optional[torch.Tensor] = None, **_visualization_data):
        rep = self._compute_output_representation(input_sequence_ids, input_seq_num_subtokens, token_sequence_lengths, edges,
                                                  edge_types)
        candidates = self._at(rep, candidate_location_idxs)
        candidate_to_sample = candidate_location_idxs[:, 0]
        swap_lp, text_lp, misuse_lp, (swap_sel, text_sel, misuse_sel) = self._compute_repair_logprobs(
            rep, target_rewrite_node_ids, target_rewrites, rewrite_to_location_group, varmisused_node_ids,
            candidate_symbol_node_ids, candidate_symbol_to_location_group, call_node_ids, candidate_swapped_node_ids,
            swapped_pair_to_call_location_group, return_selected=True)

        if rewrite_logprobs is not None:  # selector training
            _, localization_logprobs, arange = self.__localization_module.compute_localization_logprobs(
                candidate_reprs=candidates, candidate_to_sample_idx=candidate_to_sample, num_samples=has_bug.shape[0])
            loss = compute_generator_loss(
                swap_lp, arange, candidate_rewrite_idxs, candidate_symbol_to_location_group, localization_logprobs,
                self.__generator_loss_type, pair_rewrite_idxs, rewrite_logprobs, rewrite_to_graph_id,
                rewrite_to_location_group, swapped_pair_to_call_location_group, text_lp, text_rewrite_idxs, misuse_lp)
            zero = torch.zeros((), device=loss.device)
            self.__accumulate(loss, zero, zero)
            return loss

        localization_loss = self.__localization_module(
            candidate_reprs=candidates, candidate_to_sample_idx=candidate_to_sample, has_bug=has_bug,
            correct_candidate_idxs=target_location_idxs)
        repair_loss = (self._text_repair_module(text_lp, correct_rewrite_idxs, selected_fixes=text_sel).sum()
                       + self._varmisuse_module(misuse_lp, correct_candidate_symbols, selected_fixes=misuse_sel).sum()
                       + self._argswap_module(swap_lp, correct_swapped_pair, selected_fixes=swap_sel).sum())
        repair_loss = repair_loss * self._buggy_samples_weight_schedule(self._epoch_idx)
        loss = localization_loss + repair_loss / has_bug.shape[0]
        self.__accumulate(loss, repair_loss, has_bug.sum())
        return loss


_PAIR_KEYS = ("candidate_location_idxs", "target_rewrite_node_ids", "varmisused_node_ids", "candidate_symbol_node_ids",
              "call_node_ids")
_FLAT_KEYS = ("target_location_idxs", "target_rewrites", "rewrite_to_location_group", "correct_rewrite_idxs",
              "candidate_symbol_to_location_group", "correct_candidate_symbols", "swapped_pair_to_call_location_group",
              "correct_swapped_pair", "text_rewrite_idxs", "candidate_rewrite_idxs", "pair_rewrite_idxs",
              "rewrite_to_graph_id", "edge_types")


class SeqBugLabModel(AbstractNeuralModel[BugLabData, SeqModelTensorizedSample, Any], AbstractBugLabModel):
    def __init__(self, representation_size: int, max_subtoken_vocab_size: int, dropout_rate: float,
                 layer_type: str = "great", max_seq_size: int = 500, num_heads: int = 8, num_layers: int = 6,
                 intermediate_dimension_size: int = 2048,
                 buggy_samples_weight_schedule: Callable[[int], float] = lambda _: 1.0,
                 generator_loss_type: Optional[str] = "classify-max-loss", rezero_mode: str = "off",
                 normalisation_mode: str = "postnorm"):
        super().__init__()
        self._init()
        if layer_type not in ("great", "rat", "transformer", "gru"):
            raise ValueError(f"Unrecognized layer type `{layer_type}`.")
        self.layer_type, self.max_seq_size = layer_type, max_seq_size
        self.representation_size, self.dropout_rate = representation_size, dropout_rate
        self.num_heads, self.num_layers, self.intermediate_dimension_size = num_heads, num_layers, intermediate_dimension_size
        self.generator_loss_type, self.rezero_mode, self.normalisation_mode = generator_loss_type, rezero_mode, normalisation_mode
        self.buggy_samples_weight_schedule = buggy_samples_weight_schedule
        self.__token_embedder = StrElementRepresentationModel(
            token_splitting="subtoken", embedding_size=representation_size, dropout_rate=dropout_rate,
            vocabulary_size=max_subtoken_vocab_size, subtoken_combination="max")
        self.__edge_types_seen = set()
        self.__edge_types: Optional[Tuple[str, ...]] = None
        self.__edge_type_to_idx: Dict[str, int] = {}

    @property
    def token_embedder(self) -> StrElementRepresentationModel:
        return self.__token_embedder

    @property
    def edge_types(self) -> Tuple[str, ...]:
        return self.__edge_types

    # ---- metadata (seqmodel.py:626-644) ----------------------------------------------------------
    @staticmethod
    def _project(graph: BugLabGraph):
        try:
            return project_graph_to_tokens(graph)
        except Exception as ex:  # noqa: BLE001 - any failure drops the sample, as in the reference
            LOGGER.exception("Error in generating token sequence for %s", graph["path"], exc_info=ex)
            return None

    def update_metadata_from(self, datapoint: BugLabData) -> None:
        projected = self._project(datapoint["graph"])
        if projected is None:
            return
        labels, _, relations, _ = projected
        for label in labels:
            self.__token_embedder.update_metadata_from(label)
        self.__edge_types_seen.update(relations.keys())

    def finalize_metadata(self) -> None:
        # The reference freezes ``list(set)`` (an order that changes with PYTHONHASHSEED); sorted here so that ranks and
        # runs agree.  Only the numbering of the relation kinds differs, not which edge gets which kind.
        self.__edge_types = tuple(sorted(self.__edge_types_seen))
        self.__edge_type_to_idx = {kind: i for i, kind in enumerate(self.__edge_types)}
        self.__edge_types_seen = None

    def build_neural_module(self) -> SeqBugLabModule:
        return SeqBugLabModule(
            token_embedder=self.__token_embedder.build_neural_module(), embedding_dim=self.__token_embedder.embedding_size,
            num_edge_types=len(self.__edge_types), num_layers=self.num_layers, num_heads=self.num_heads,
            intermediate_dimension=self.intermediate_dimension_size, dropout_rate=self.dropout_rate,
            rewrite_vocabulary_size=len(self._target_rewrite_ops), layer_type=self.layer_type,
            buggy_samples_weight_schedule=self.buggy_samples_weight_schedule,
            generator_loss_type=self.generator_loss_type, rezero_mode=self.rezero_mode,
            normalisation_mode=self.normalisation_mode)

    # ---- one sample (seqmodel.py:646-729) --------------------------------------------------------
    def tensorize(self, datapoint: BugLabData) -> Optional[SeqModelTensorizedSample]:
        if "candidate_rewrite_logprobs" in datapoint:
            assert not self._tensorize_only_at_target_location_rewrites
        projected = self._project(datapoint["graph"])
        if projected is None:
            return None
        labels, position, relations, _ = projected
        if len(labels) > self.max_seq_size:
            LOGGER.debug("Rejecting sample with %s tokens.", len(labels))
            return None

        graph = datapoint["graph"]
        # several graph nodes may share one token position, so the transformed candidates can contain duplicates
        candidate_nodes, inverse = np.unique(graph["reference_nodes"], return_inverse=True)
        candidate_positions = np.array([position[n] for n in candidate_nodes])
        target = datapoint["target_fix_action_idx"]
        if target is not None:
            target_location = inverse[target]
            assert position[graph["reference_nodes"][target]] == candidate_positions[target_location]
        else:
            target_location = None

        (text_nodes, text_ops, text_groups, text_correct, text_orig,
         misuse_nodes, misuse_groups, misuse_candidates, misuse_correct, misuse_orig,
         call_nodes, swapped_pairs, swap_correct, swap_groups, swap_orig,
         location_groups) = self._compute_rewrite_data(datapoint, candidate_nodes)

        embedder = self.__token_embedder
        return SeqModelTensorizedSample(
            target_subtokens_ids=[embedder.tensorize(label) for label in labels],
            intra_token_edges=relations,
            candidate_location_idxs=candidate_positions,
            target_location_idx=target_location,
            node_mappings=position,
            target_rewrite_node_ids=[position[n] for n in text_nodes],
            target_rewrites=text_ops, target_rewrite_to_location_group=text_groups,
            correct_rewrite_target=text_correct, text_rewrite_original_idx=text_orig,
            varmisused_node_ids=[position[n] for n in misuse_nodes],
            candidate_symbol_node_ids=[position[n] for n in misuse_candidates],
            candidate_symbol_to_varmisused_node=misuse_groups, correct_candidate_symbol_node=misuse_correct,
            candidate_rewrite_original_idx=misuse_orig,
            call_node_ids=[position[n] for n in call_nodes],
            candidate_swapped_node_ids=[(position[a], position[b]) for a, b in swapped_pairs],
            swapped_pair_to_call=swap_groups, correct_swapped_pair=swap_correct, pair_rewrite_original_idx=swap_orig,
            num_rewrite_locations_considered=len(location_groups),
            rewrite_logprobs=datapoint.get("candidate_rewrite_logprobs", None))

    # ---- minibatch (seqmodel.py:731-975) ---------------------------------------------------------
    def initialize_minibatch(self) -> Dict[str, Any]:
        mb: Dict[str, Any] = {key: [] for key in _PAIR_KEYS + _FLAT_KEYS}
        mb.update({
            "input_subtoken_ids": [], "edges": [], "candidate_swapped_node_ids": [], "has_bug": [], "node_mappings": [],
            "mb_num_repair_groups": 0, "mb_num_rewrite_candidates": 0, "num_candidate_locations": 0,
            "num_text": 0, "num_misuse": 0, "num_swap": 0,
            "text_rewrite_original_idxs": [], "candidate_rewrite_original_idxs": [], "pair_rewrite_original_idx": [],
            "rewrite_logprobs": [], "no_bug_rewrite_logprobs": [],
        })
        return mb

    def extend_minibatch_with(self, tensorized_datapoint: SeqModelTensorizedSample, partial_minibatch: Dict[str, Any]) -> bool:
        t, mb = tensorized_datapoint, partial_minibatch
        sample = len(mb["input_subtoken_ids"])
        mb["input_subtoken_ids"].append(t.target_subtokens_ids)
        for kind, adjacency in t.intra_token_edges.items():
            if adjacency:
                pairs = np.asarray(adjacency, dtype=np.int64).reshape(-1, 2)
                mb["edges"].append(np.concatenate((np.full((pairs.shape[0], 1), sample, dtype=np.int64), pairs), axis=1))
                mb["edge_types"].append(np.full(pairs.shape[0], self.__edge_type_to_idx[kind], dtype=np.int64))

        mb["has_bug"].append(t.target_location_idx is not None)
        mb["node_mappings"].append(t.node_mappings)
        mb["target_location_idxs"].append(np.int64((t.target_location_idx or 0) + mb["num_candidate_locations"]))
        mb["num_candidate_locations"] += len(t.candidate_location_idxs)

        def with_sample(positions) -> np.ndarray:
            positions = np.asarray(positions, dtype=np.int64).reshape(-1)
            return np.stack((np.full(positions.shape[0], sample, dtype=np.int64), positions), axis=1)

        mb["candidate_location_idxs"].append(with_sample(t.candidate_location_idxs))
        group_offset, rewrite_offset = mb["mb_num_repair_groups"], mb["mb_num_rewrite_candidates"]

        def family(correct, correct_key, count_key, groups, groups_key, original, original_key) -> int:
            if correct is not None:
                mb[correct_key].append(np.int64(correct + mb[count_key]))
            mb[groups_key].append(np.asarray(groups, dtype=np.int64) + group_offset)
            mb[original_key].append(np.asarray(original, dtype=np.int64) + rewrite_offset)
            mb[count_key] += len(groups)
            return len(original)

        n = family(t.correct_rewrite_target, "correct_rewrite_idxs", "num_text", t.target_rewrite_to_location_group,
                   "rewrite_to_location_group", t.text_rewrite_original_idx, "text_rewrite_idxs")
        mb["target_rewrite_node_ids"].append(with_sample(t.target_rewrite_node_ids))
        mb["target_rewrites"].append(np.asarray(t.target_rewrites, dtype=np.int64))
        n += family(t.correct_candidate_symbol_node, "correct_candidate_symbols", "num_misuse",
                    t.candidate_symbol_to_varmisused_node, "candidate_symbol_to_location_group",
                    t.candidate_rewrite_original_idx, "candidate_rewrite_idxs")
        mb["varmisused_node_ids"].append(with_sample(t.varmisused_node_ids))
        mb["candidate_symbol_node_ids"].append(with_sample(t.candidate_symbol_node_ids))
        n += family(t.correct_swapped_pair, "correct_swapped_pair", "num_swap", t.swapped_pair_to_call,
                    "swapped_pair_to_call_location_group", t.pair_rewrite_original_idx, "pair_rewrite_idxs")
        mb["call_node_ids"].append(with_sample(t.call_node_ids))
        pairs = np.asarray(t.candidate_swapped_node_ids, dtype=np.int64).reshape(-1, 2)
        mb["candidate_swapped_node_ids"].append(
            np.concatenate((np.full((pairs.shape[0], 1), sample, dtype=np.int64), pairs), axis=1))
        mb["mb_num_rewrite_candidates"] += n
        mb["mb_num_repair_groups"] += t.num_rewrite_locations_considered
        mb["rewrite_to_graph_id"].append(np.full(n, sample, dtype=np.int64))

        mb["text_rewrite_original_idxs"].append(t.text_rewrite_original_idx)
        mb["candidate_rewrite_original_idxs"].append(t.candidate_rewrite_original_idx)
        mb["pair_rewrite_original_idx"].append(t.pair_rewrite_original_idx)
        if t.rewrite_logprobs is not None:
            mb["rewrite_logprobs"].extend(t.rewrite_logprobs[:-1])
            mb["no_bug_rewrite_logprobs"].append(t.rewrite_logprobs[-1])
        return True

    def finalize_minibatch(self, accumulated_minibatch_data: Dict[str, Any], device: Union[str, torch.device]) -> Dict[str, Any]:
        mb = accumulated_minibatch_data
        sequences = mb["input_subtoken_ids"]
        num_samples, T = len(sequences), self.__token_embedder.max_num_subtokens
        lengths = np.array([len(seq) for seq in sequences], dtype=np.int64)
        max_len = int(lengths.max())
        ids = np.zeros((num_samples, max_len, T), dtype=np.int64)
        num_subtokens = np.ones((num_samples, max_len), dtype=np.int64)   # padding positions count one (pad) subtoken
        for i, seq in enumerate(sequences):
            for j, subtokens in enumerate(seq):
                n = min(len(subtokens), T)
                ids[i, j, :n] = subtokens[:n]
                num_subtokens[i, j] = n

        def cat(chunks: List[np.ndarray], width: Optional[int] = None) -> torch.Tensor:
            if chunks:
                arr = np.concatenate([np.atleast_1d(c) for c in chunks]) if width is None else np.concatenate(chunks, axis=0)
            else:
                arr = np.zeros((0,) if width is None else (0, width), dtype=np.int64)
            return torch.from_numpy(np.ascontiguousarray(arr, dtype=np.int64)).to(device)

        minibatch: Dict[str, Any] = {
            "input_sequence_ids": torch.from_numpy(ids).to(device),
            "input_seq_num_subtokens": torch.from_numpy(num_subtokens).to(device),
            "token_sequence_lengths": torch.from_numpy(lengths).to(device),
            "edges": cat(mb["edges"], 3),
            "has_bug": torch.tensor(mb["has_bug"], dtype=torch.bool, device=device),
            "candidate_swapped_node_ids": cat(mb["candidate_swapped_node_ids"], 3),
            "node_mappings": mb["node_mappings"],
            "text_rewrite_original_idxs": mb["text_rewrite_original_idxs"],
            "candidate_rewrite_original_idxs": mb["candidate_rewrite_original_idxs"],
            "pair_rewrite_original_idx": mb["pair_rewrite_original_idx"],
        }
        for key in _PAIR_KEYS:
            minibatch[key] = cat(mb[key], 2)
        for key in _FLAT_KEYS:
            minibatch[key] = cat(mb[key])
        if mb["rewrite_logprobs"]:
            minibatch["rewrite_logprobs"] = torch.tensor(mb["rewrite_logprobs"] + mb["no_bug_rewrite_logprobs"],
                                                         dtype=torch.float32, device=device)
        return minibatch

    def predict(self, data: Iterator[BugLabData], trained_nn: SeqBugLabModule, device, parallelize: bool
                ) -> Iterator[Tuple[BugLabData, Dict[int, float], List[float]]]:
        """Per sample: {graph node (or -1 = NO_BUG): log-prob} and one log-prob per candidate rewrite, scored at every
        location (reference seqmodel.py:977-1031; minibatches of <= 50 samples)."""
        trained_nn.eval()
        with torch.no_grad(), self._tensorize_all_location_rewrites():
            for mb, original_datapoints in self.minibatch_iterator(
                    self.tensorize_dataset(data, return_input_data=True, parallelize=parallelize), device,
                    max_minibatch_size=50, parallelize=parallelize):
                num_samples = mb["input_sequence_ids"].shape[0]
                rep = trained_nn._compute_output_representation(
                    mb["input_sequence_ids"], mb["input_seq_num_subtokens"], mb["token_sequence_lengths"], mb["edges"],
                    mb["edge_types"])
                where = mb["candidate_location_idxs"]
                groups, logprobs = trained_nn._compute_localization_logprobs(rep[where[:, 0], where[:, 1]], where[:, 0],
                                                                             num_samples)
                swap_lp, text_lp, misuse_lp = trained_nn._compute_repair_logprobs(
                    rep, mb["target_rewrite_node_ids"], mb["target_rewrites"], mb["rewrite_to_location_group"],
                    mb["varmisused_node_ids"], mb["candidate_symbol_node_ids"], mb["candidate_symbol_to_location_group"],
                    mb["call_node_ids"], mb["candidate_swapped_node_ids"], mb["swapped_pair_to_call_location_group"])
                yield from self._iter_per_sample_results(
                    mb, groups.cpu().numpy(), logprobs.cpu().numpy(), swap_lp, num_samples, original_datapoints, text_lp,
                    misuse_lp, node_mappings=mb["node_mappings"])
