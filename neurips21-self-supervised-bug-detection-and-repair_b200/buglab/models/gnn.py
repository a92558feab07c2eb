"""GNN bug detector / repair model (reference: buglab/models/gnn.py:29-645).

``GnnBugLabModel`` is the host half: it turns a ``BugLabData`` sample into numpy int32 tables, packs a minibatch
with vectorised offset arithmetic and ships it in one pinned staging buffer.  ``GnnBugLabModule`` is the device
half: GNN (buglab_b200 kernels) -> candidate-node gather -> localisation head + three repair heads -> loss.
Names, keyword arguments, minibatch keys and state_dict keys follow the reference so checkpoints and callers
(train.py, evaluate.py, the self-supervised controllers) are interchangeable.
"""
import logging
from typing import Any, Callable, Dict, Iterator, List, NamedTuple, Optional, Tuple, Union

import numpy as np
import torch
from ptgnn.baseneuralmodel import AbstractNeuralModel, ModuleWithMetrics
from ptgnn.neuralmodels.gnn import GnnOutput, GraphData, GraphNeuralNetwork, GraphNeuralNetworkModel, TensorizedGraphData
from torch import nn

from buglab.models.basemodel import AbstractBugLabModel
from buglab.models.layers.fixermodules import (
    CandidatePairSelectorModule,
    SingleCandidateNodeSelectorModule,
    TextRepairModule,
)
from buglab.models.layers.localizationmodule import LocalizationModule
from buglab.models.utils import compute_generator_loss, scatter_log_softmax, scatter_max
from buglab.representations.data import BugLabData

LOGGER = logging.getLogger(__name__)


class BaseTensorizedBugLabGnn(NamedTuple):
    graph_data: TensorizedGraphData
    # localisation
    target_location_node_idx: Optional[int]
    # repair: text rewrites
    target_rewrites: List[int]
    target_rewrite_to_location_group: List[int]
    correct_rewrite_target: Optional[int]
    text_rewrite_original_idx: List[int]
    # repair: variable misuse
    candidate_symbol_to_varmisused_node: List[int]
    correct_candidate_symbol_node: Optional[int]
    candidate_rewrite_original_idx: List[int]
    # repair: argument swaps
    swapped_pair_to_call: List[int]
    correct_swapped_pair: Optional[int]
    pair_rewrite_original_idx: List[int]
    num_rewrite_locations_considered: int
    # bug selection
    rewrite_logprobs: Optional[List[float]]


def _const_one(_epoch: int) -> float:
    return 1.0


class GnnBugLabModule(ModuleWithMetrics):
    def __init__(self, gnn: GraphNeuralNetwork, rewrite_vocabulary_size: int, use_all_gnn_layer_outputs: bool = False,
                 generator_loss_type: Optional[str] = "norm-kl",
                 buggy_samples_weight_schedule: Callable[[int], float] = _const_one):
        super().__init__()
        self.__generator_loss_type = generator_loss_type
        self._gnn = gnn
        self.__use_all_gnn_layer_outputs = use_all_gnn_layer_outputs
        if use_all_gnn_layer_outputs:
            self.__summarization_layer = nn.Linear(
                gnn.input_node_state_dim + sum(l.output_state_dimension for l in gnn.message_passing_layers),
                gnn.output_node_state_dim)
        dim = gnn.output_node_state_dim
        self.__localization_module = LocalizationModule(dim, buggy_samples_weight_schedule=buggy_samples_weight_schedule)
        self._buggy_samples_weight_schedule = buggy_samples_weight_schedule
        self._text_repair_module = TextRepairModule(dim, rewrite_vocabulary_size)
        self._varmisuse_module = SingleCandidateNodeSelectorModule(dim)
        self._argswap_module = CandidatePairSelectorModule(dim)

    @property
    def use_all_gnn_layer_outputs(self):
        return self.__use_all_gnn_layer_outputs

    @property
    def gnn(self):
        return self._gnn

    # ---- metrics (device-resident; one D2H when reported) ---------------------------------------
    def _reset_module_metrics(self) -> None:
        if not hasattr(self, "_epoch_idx"):
            self._epoch_idx = 0
        elif self.training and self.__num_batches > 0:
            self._epoch_idx += 1
        self.__sums = None  # [loss, repair_loss, buggy samples]
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

    # ---- forward pieces -------------------------------------------------------------------------
    def __compute_gnn_output(self, graph_data) -> GnnOutput:
        graph_data = {k: v for k, v in graph_data.items() if k != "h2d_bytes"}
        out: GnnOutput = self._gnn(**graph_data, return_all_states=self.__use_all_gnn_layer_outputs)
        if self.__use_all_gnn_layer_outputs:
            out = out._replace(output_node_representations=self.__summarization_layer(out.output_node_representations))
        return out

    def compute_localization_logprobs(self, graph_data: Dict[str, Any]):
        gnn_output = self.__compute_gnn_output(graph_data)
        candidate_reprs = gnn_output.output_node_representations[gnn_output.node_idx_references["candidate_nodes"]]
        groups, log_probs, arange = self.__localization_module.compute_localization_logprobs(
            candidate_reprs=candidate_reprs,
            candidate_to_sample_idx=gnn_output.node_graph_idx_reference["candidate_nodes"],
            num_samples=gnn_output.num_graphs)
        return groups, log_probs, gnn_output, arange

    def forward(self, *, graph_data: Dict[str, Any], correct_candidate_node_idxs, has_bug: torch.Tensor,
                target_rewrites, rewrite_to_location_group, correct_rewrite_idxs, text_rewrite_idxs,
                candidate_symbol_to_location_group, correct_candidate_symbols, candidate_rewrite_idxs,
                swapped_pair_to_call_location_group, correct_swapped_pair, pair_rewrite_idxs,
                rewrite_to_graph_id, rewrite_logprobs: Optional[torch.Tensor] = None, **_visualization_data):
        gnn_output = self.__compute_gnn_output(graph_data)
        states = gnn_output.output_node_representations
        candidate_reprs = states[gnn_output.node_idx_references["candidate_nodes"]]  # [C, H]
        candidate_to_sample = gnn_output.node_graph_idx_reference["candidate_nodes"]

        swap_lp, text_lp, misuse_lp, (swap_sel, text_sel, misuse_sel) = self._compute_repair_logprobs(
            gnn_output, target_rewrites, rewrite_to_location_group, candidate_symbol_to_location_group,
            swapped_pair_to_call_location_group)

        if rewrite_logprobs is not None:  # selector ("bug generator") training — reference gnn.py:189-219
            _, localization_logprobs, arange = self.__localization_module.compute_localization_logprobs(
                candidate_reprs=candidate_reprs, candidate_to_sample_idx=candidate_to_sample, num_samples=has_bug.shape[0])
            loss = compute_generator_loss(
                swap_lp, arange, candidate_rewrite_idxs, candidate_symbol_to_location_group, localization_logprobs,
                self.__generator_loss_type, pair_rewrite_idxs, rewrite_logprobs, rewrite_to_graph_id,
                rewrite_to_location_group, swapped_pair_to_call_location_group, text_lp, text_rewrite_idxs, misuse_lp)
            zero = torch.zeros((), device=loss.device)
            self.__accumulate(loss, zero, zero)
            return loss

        localization_loss = self.__localization_module(
            candidate_reprs=candidate_reprs, candidate_to_sample_idx=candidate_to_sample, has_bug=has_bug,
            correct_candidate_idxs=correct_candidate_node_idxs)
        repair_loss = (self._text_repair_module(text_lp, correct_rewrite_idxs, selected_fixes=text_sel).sum()
                       + self._varmisuse_module(misuse_lp, correct_candidate_symbols, selected_fixes=misuse_sel).sum()
                       + self._argswap_module(swap_lp, correct_swapped_pair, selected_fixes=swap_sel).sum())
        repair_loss = repair_loss * self._buggy_samples_weight_schedule(self._epoch_idx)
        loss = localization_loss + repair_loss / has_bug.shape[0]
        self.__accumulate(loss, repair_loss, has_bug.sum())
        return loss

    def _compute_repair_logprobs(self, gnn_output, target_rewrites, rewrite_to_location_group,
                                 candidate_symbol_to_location_group, swapped_pair_to_call_location_group):
        """Logits of the three repair heads, normalised jointly per location group (reference gnn.py:253-322)."""
        states = gnn_output.output_node_representations
        refs = gnn_output.node_idx_references
        dev = states.device
        if target_rewrites.shape[0] > 0:
            text_logits = self._text_repair_module.compute_rewrite_logits(states[refs["target_rewrite_nodes"]], target_rewrites)
        else:
            text_logits = torch.zeros(0, device=dev)
        if refs["varmisused_node_ids"].shape[0] > 0:
            misuse_logits = self._varmisuse_module.compute_per_slot_log_probability(
                states[refs["varmisused_node_ids"]], states[refs["candidate_symbol_node_ids"]])
        else:
            misuse_logits = torch.zeros(0, device=dev)
        if refs["call_node_ids"].shape[0] > 0:
            swap_logits = self._argswap_module.compute_per_pair_logits(
                states[refs["call_node_ids"]], states[refs["candidate_swapped_node_ids"]])
        else:
            swap_logits = torch.zeros(0, device=dev)

        sizes = [text_logits.shape[0], misuse_logits.shape[0], swap_logits.shape[0]]
        all_logits = torch.cat((text_logits, misuse_logits, swap_logits))
        groups = torch.cat((rewrite_to_location_group, candidate_symbol_to_location_group,
                            swapped_pair_to_call_location_group))
        if all_logits.shape[0] == 0:
            empty_b = torch.zeros(0, dtype=torch.bool, device=dev)
            return swap_logits, text_logits, misuse_logits, (empty_b, empty_b, empty_b)
        text_lp, misuse_lp, swap_lp = torch.split(scatter_log_softmax(all_logits, index=groups), sizes)
        with torch.no_grad():
            group_max = scatter_max(all_logits, groups)[0]
            text_sel, misuse_sel, swap_sel = torch.split(group_max[groups] == all_logits, sizes)
        return swap_lp, text_lp, misuse_lp, (swap_sel, text_sel, misuse_sel)


# the per-sample integer tables that extend_minibatch offsets and finalize_minibatch ships, in wire order
_INDEX_KEYS = (
    "correct_candidate_node_idxs", "target_rewrites", "rewrite_to_location_group", "correct_rewrite_idxs",
    "text_rewrite_idxs", "candidate_symbol_to_location_group", "correct_candidate_symbols", "candidate_rewrite_idxs",
    "swapped_pair_to_call_location_group", "correct_swapped_pair", "pair_rewrite_idxs", "rewrite_to_graph_id",
)


class GnnBugLabModel(AbstractNeuralModel[BugLabData, BaseTensorizedBugLabGnn, GnnBugLabModule], AbstractBugLabModel):
    def __init__(self, gnn_model: GraphNeuralNetworkModel, use_all_gnn_layer_outputs: bool = False,
                 generator_loss_type: Optional[str] = "classify-max-loss",
                 buggy_samples_weight_schedule: Callable[[int], float] = _const_one):
        super().__init__()
        self._init()
        self.__gnn_model = gnn_model
        self.__use_all_gnn_layer_outputs = use_all_gnn_layer_outputs
        self.__generator_loss_type = generator_loss_type
        self.__buggy_samples_weight_schedule = buggy_samples_weight_schedule

    @property
    def gnn_model(self) -> GraphNeuralNetworkModel:
        return self.__gnn_model

    @property
    def use_all_gnn_layer_outputs(self):
        return self.__use_all_gnn_layer_outputs

    def update_metadata_from(self, datapoint: BugLabData) -> None:
        graph_data, _ = BugLabData.as_graph_data(datapoint)
        self.__gnn_model.update_metadata_from(graph_data)

    def build_neural_module(self) -> GnnBugLabModule:
        return GnnBugLabModule(
            self.__gnn_model.build_neural_module(),
            rewrite_vocabulary_size=len(self._target_rewrite_ops),
            use_all_gnn_layer_outputs=self.__use_all_gnn_layer_outputs,
            generator_loss_type=self.__generator_loss_type,
            buggy_samples_weight_schedule=self.__buggy_samples_weight_schedule)

    # ---- one sample -----------------------------------------------------------------------------
    def tensorize(self, datapoint: BugLabData) -> Optional[BaseTensorizedBugLabGnn]:
        graph_data, target_location_node_idx = BugLabData.as_graph_data(datapoint)
        if "candidate_rewrite_logprobs" in datapoint:
            assert not self._tensorize_only_at_target_location_rewrites
        rewrite_data = self._compute_rewrite_data(datapoint, graph_data.reference_nodes["candidate_nodes"])
        self._add_rewrite_reference_nodes(graph_data.reference_nodes, rewrite_data)
        tensorized_graph = self.__gnn_model.tensorize(graph_data)
        if tensorized_graph is None:
            return None
        return self._assemble_tensorized(tensorized_graph, target_location_node_idx, rewrite_data,
                                         datapoint.get("candidate_rewrite_logprobs", None))

    @staticmethod
    def _add_rewrite_reference_nodes(refs: Dict[str, Any], rewrite_data) -> None:
        (text_nodes, _ops, _tg, text_correct, _to, misuse_nodes, _mg, misuse_candidates, misuse_correct, _mo,
         call_nodes, swapped_pairs, swap_correct, _sg, _so, _groups) = rewrite_data
        refs["target_rewrite_nodes"] = text_nodes
        refs["varmisused_node_ids"] = misuse_nodes
        refs["candidate_symbol_node_ids"] = misuse_candidates
        refs["call_node_ids"] = call_nodes
        refs["candidate_swapped_node_ids"] = swapped_pairs if len(swapped_pairs) else np.zeros((0, 2), dtype=np.int32)
        assert sum(c is not None for c in (text_correct, misuse_correct, swap_correct)) <= 1, \
            "No more than one node should be correct."

    @staticmethod
    def _assemble_tensorized(tensorized_graph, target_location_node_idx, rewrite_data,
                             rewrite_logprobs) -> BaseTensorizedBugLabGnn:
        (_tn, text_ops, text_groups, text_correct, text_orig, _mn, misuse_groups, _mc, misuse_correct, misuse_orig,
         _cn, _sp, swap_correct, swap_groups, swap_orig, location_groups) = rewrite_data
        return BaseTensorizedBugLabGnn(
            graph_data=tensorized_graph,
            target_location_node_idx=target_location_node_idx,
            target_rewrites=text_ops, target_rewrite_to_location_group=text_groups,
            correct_rewrite_target=text_correct, text_rewrite_original_idx=text_orig,
            candidate_symbol_to_varmisused_node=misuse_groups, correct_candidate_symbol_node=misuse_correct,
            candidate_rewrite_original_idx=misuse_orig,
            swapped_pair_to_call=swap_groups, correct_swapped_pair=swap_correct, pair_rewrite_original_idx=swap_orig,
            num_rewrite_locations_considered=len(location_groups),
            rewrite_logprobs=rewrite_logprobs)

    # ---- minibatch ------------------------------------------------------------------------------
    def initialize_minibatch(self) -> Dict[str, Any]:
        mb: Dict[str, Any] = {key: [] for key in _INDEX_KEYS}
        mb.update({
            "graph_data": self.__gnn_model.initialize_minibatch(),
            "has_bug": [],
            "mb_num_target_nodes": 0, "mb_num_rewrite_candidates": 0, "mb_num_repair_groups": 0,
            "num_text": 0, "num_misuse": 0, "num_swap": 0,
            "text_rewrite_original_idxs": [], "candidate_rewrite_original_idxs": [], "pair_rewrite_original_idx": [],
            "rewrite_logprobs": [], "no_bug_rewrite_logprobs": [],
        })
        return mb

    def extend_minibatch_with(self, tensorized_datapoint: BaseTensorizedBugLabGnn, partial_minibatch: Dict[str, Any]) -> bool:
        t, mb = tensorized_datapoint, partial_minibatch
        keep_extending = self.__gnn_model.extend_minibatch_with(t.graph_data, mb["graph_data"])
        graph_idx = len(mb["graph_data"]["num_nodes_per_graph"]) - 1

        # localisation target: index among the minibatch's candidate nodes (0 when the sample has no bug)
        has_bug = t.target_location_node_idx is not None
        mb["has_bug"].append(has_bug)
        mb["correct_candidate_node_idxs"].append(
            np.int32(t.target_location_node_idx + mb["mb_num_target_nodes"] if has_bug else 0))
        mb["mb_num_target_nodes"] += len(t.graph_data.reference_nodes["candidate_nodes"])

        group_offset, rewrite_offset = mb["mb_num_repair_groups"], mb["mb_num_rewrite_candidates"]

        def family(correct, correct_key, count_key, groups, groups_key, original, original_key):
            if correct is not None:
                mb[correct_key].append(np.int32(correct + mb[count_key]))
            mb[groups_key].append(np.asarray(groups, dtype=np.int32) + group_offset)
            mb[original_key].append(np.asarray(original, dtype=np.int32) + rewrite_offset)
            mb[count_key] += len(groups)
            return len(original)

        n = family(t.correct_rewrite_target, "correct_rewrite_idxs", "num_text", t.target_rewrite_to_location_group,
                   "rewrite_to_location_group", t.text_rewrite_original_idx, "text_rewrite_idxs")
        mb["target_rewrites"].append(np.asarray(t.target_rewrites, dtype=np.int32))
        n += family(t.correct_candidate_symbol_node, "correct_candidate_symbols", "num_misuse",
                    t.candidate_symbol_to_varmisused_node, "candidate_symbol_to_location_group",
                    t.candidate_rewrite_original_idx, "candidate_rewrite_idxs")
        n += family(t.correct_swapped_pair, "correct_swapped_pair", "num_swap", t.swapped_pair_to_call,
                    "swapped_pair_to_call_location_group", t.pair_rewrite_original_idx, "pair_rewrite_idxs")
        mb["mb_num_rewrite_candidates"] += n
        mb["mb_num_repair_groups"] += t.num_rewrite_locations_considered
        mb["rewrite_to_graph_id"].append(np.full(n, graph_idx, dtype=np.int32))

        # kept on the host for unpacking predictions / visualisation
        mb["text_rewrite_original_idxs"].append(t.text_rewrite_original_idx)
        mb["candidate_rewrite_original_idxs"].append(t.candidate_rewrite_original_idx)
        mb["pair_rewrite_original_idx"].append(t.pair_rewrite_original_idx)
        if t.rewrite_logprobs is not None:
            mb["rewrite_logprobs"].extend(t.rewrite_logprobs[:-1])
            mb["no_bug_rewrite_logprobs"].append(t.rewrite_logprobs[-1])
        return keep_extending

    def finalize_minibatch(self, accumulated_minibatch_data: Dict[str, Any], device: Union[str, torch.device]) -> Dict[str, Any]:
        mb = accumulated_minibatch_data
        device = torch.device(device)
        graph_data = self.__gnn_model.finalize_minibatch(mb["graph_data"], device)

        # all head index tables + has_bug in ONE pinned staging buffer / one H2D copy
        arrays = []
        for key in _INDEX_KEYS:
            chunks = mb[key]
            if not chunks:
                arrays.append(np.zeros(0, dtype=np.int32))
            elif np.ndim(chunks[0]) == 0:   # per-sample scalars (every chunk of a key has the same rank)
                arrays.append(np.asarray(chunks, dtype=np.int32))
            else:
                arrays.append(np.concatenate(chunks).astype(np.int32, copy=False))
        arrays.append(np.asarray(mb["has_bug"], dtype=np.int32))
        sizes = [a.shape[0] for a in arrays]
        staging = torch.empty(max(sum(sizes), 1), dtype=torch.int32, pin_memory=(device.type == "cuda"))
        host = staging.numpy()
        off = 0
        for a in arrays:
            host[off: off + a.shape[0]] = a
            off += a.shape[0]
        on_device = staging.to(device, non_blocking=True).long()  # the reference hands int64 tensors to the module
        views = torch.split(on_device[: sum(sizes)], sizes)
        minibatch: Dict[str, Any] = {"graph_data": graph_data}
        for key, view in zip(_INDEX_KEYS, views):
            minibatch[key] = view
        minibatch["has_bug"] = views[-1].bool()
        minibatch["text_rewrite_original_idxs"] = mb["text_rewrite_original_idxs"]
        minibatch["candidate_rewrite_original_idxs"] = mb["candidate_rewrite_original_idxs"]
        minibatch["pair_rewrite_original_idx"] = mb["pair_rewrite_original_idx"]
        graph_data["h2d_bytes"] = graph_data.get("h2d_bytes", 0) + 4 * sum(sizes)
        if mb["rewrite_logprobs"]:
            minibatch["rewrite_logprobs"] = torch.tensor(mb["rewrite_logprobs"] + mb["no_bug_rewrite_logprobs"],
                                                         dtype=torch.float32, device=device)
        return minibatch

    # ---- inference ------------------------------------------------------------------------------
    def predict(self, data: Iterator[BugLabData], trained_nn: GnnBugLabModule, device, parallelize: bool
                ) -> Iterator[Tuple[BugLabData, Dict[int, float], List[float]]]:
        """Per sample: {candidate node (or -1 = NO_BUG): log-prob} and one log-prob per candidate rewrite, scored at
        EVERY location (reference gnn.py:606-645; minibatches of <= 50 graphs)."""
        trained_nn.eval()
        with torch.no_grad(), self._tensorize_all_location_rewrites():
            if hasattr(data, "tensorized"):
                # a data source that tensorises itself (buglab_b200.shards.ShardDataset: native shard decoder, ~10x the
                # host chain) pairs every sample with its raw datapoint, as tensorize_dataset(return_input_data=True) does
                samples = data.tensorized(self, return_input_data=True,
                                          lazy_input_data=getattr(data, "lazy_input_data", False))
            else:
                samples = self.tensorize_dataset(data, return_input_data=True, parallelize=parallelize)
            for mb_data, original_datapoints in self.minibatch_iterator(
                    samples, device, max_minibatch_size=50, parallelize=parallelize):
                groups, log_probs, gnn_output, _ = trained_nn.compute_localization_logprobs(mb_data["graph_data"])
                swap_lp, text_lp, misuse_lp, _ = trained_nn._compute_repair_logprobs(
                    gnn_output, mb_data["target_rewrites"], mb_data["rewrite_to_location_group"],
                    mb_data["candidate_symbol_to_location_group"], mb_data["swapped_pair_to_call_location_group"])
                yield from self._iter_per_sample_results(
                    mb_data, groups.cpu().numpy(), log_probs.cpu().numpy(), swap_lp, gnn_output.num_graphs,
                    original_datapoints, text_lp, misuse_lp)
