"""``GraphNeuralNetworkModel`` (host: metadata, tensorise, pack) and ``GraphNeuralNetwork`` (device: embed + layers).

Semantics restated in SURVEY.md §8a P1/P3; reference call sites buglab/models/modelregistry.py:78-90 (constructor
keywords), buglab/models/gnn.py:350,354,403,433,466,547 (model hooks), :117 (``GraphNeuralNetwork.forward(**graph_data)``).

Host side is B200-first: graphs are kept as numpy int32 arrays, a minibatch is packed with vectorised offset
arithmetic into ONE pinned int32 staging buffer and crosses PCIe in one asynchronous copy; the typed-edge
plan (CSR + pair tables) is then built on the device once per minibatch and rides along with the adjacency
lists (``PlannedAdjacency``).  Deliberate, documented deviations from upstream ptgnn: edge types are ordered
by name (upstream: set iteration order) so that every data-parallel rank builds the same layer layout.
"""
import logging
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
from torch import nn

from ptgnn.baseneuralmodel import AbstractNeuralModel, ModuleWithMetrics

from .messagepassing.abstractmessagepassing import AbstractMessagePassingLayer, PlannedAdjacency, plan_for
from .structs import GnnOutput, GraphData, TensorizedGraphData

LOGGER = logging.getLogger(__name__)


class GraphNeuralNetwork(ModuleWithMetrics):
    def __init__(self, message_passing_layers: List[AbstractMessagePassingLayer], node_embedder: nn.Module,
                 edge_embedder: Optional[nn.Module] = None):
        super().__init__()
        self.__message_passing_layers = nn.ModuleList(message_passing_layers)
        self.__node_embedder = node_embedder
        self.__edge_embedder = edge_embedder

    def plan_block_nodes(self) -> int:
        """Pair-table layout every layer of this network can consume (0 = type-major), see ops.plan_block_nodes_for."""
        from buglab_b200 import ops

        dims = [(l.input_state_dimension, l.message_dimension) for l in self.__message_passing_layers
                if hasattr(l, "message_dimension")]
        return ops.plan_block_nodes_for(dims)

    @property
    def message_passing_layers(self) -> List[AbstractMessagePassingLayer]:
        return list(self.__message_passing_layers)

    @property
    def input_node_state_dim(self) -> int:
        return self.__node_embedder.embedding_layer.embedding_dim

    @property
    def output_node_state_dim(self) -> int:
        return self.__message_passing_layers[-1].output_state_dimension

    @property
    def node_embedder(self) -> nn.Module:
        return self.__node_embedder

    def forward(self, *, node_data: Dict[str, torch.Tensor], adjacency_lists, node_to_graph_idx: torch.Tensor,
                reference_node_ids: Dict[str, torch.Tensor], reference_node_graph_idx: Dict[str, torch.Tensor],
                num_graphs: int, edge_feature_data=None, return_all_states: bool = False, **_ignored) -> GnnOutput:
        initial_node_states = self.__node_embedder(**node_data)  # [N, H]
        num_nodes = initial_node_states.shape[0]
        if not isinstance(adjacency_lists, PlannedAdjacency):
            planned = PlannedAdjacency(adjacency_lists)
            planned.plan = plan_for(adjacency_lists, num_nodes, self.plan_block_nodes())
            adjacency_lists = planned
        elif adjacency_lists.plan is None:
            adjacency_lists.plan = plan_for(adjacency_lists, num_nodes, self.plan_block_nodes())
        states = [initial_node_states]
        for layer in self.__message_passing_layers:
            states.append(layer(node_states=states[-1], adjacency_lists=adjacency_lists,
                                node_to_graph_idx=node_to_graph_idx, reference_node_ids=reference_node_ids,
                                reference_node_graph_idx=reference_node_graph_idx, edge_features=None))
        output = torch.cat(states, dim=-1) if return_all_states else states[-1]
        return GnnOutput(
            input_node_representations=initial_node_states,
            output_node_representations=output,
            node_to_graph_idx=node_to_graph_idx,
            node_idx_references=reference_node_ids,
            node_graph_idx_reference=reference_node_graph_idx,
            num_graphs=num_graphs,
        )


class GraphNeuralNetworkModel(AbstractNeuralModel[GraphData, TensorizedGraphData, GraphNeuralNetwork]):
    def __init__(self, node_representation_model: AbstractNeuralModel,
                 message_passing_layer_creator: Callable[[int], List[AbstractMessagePassingLayer]],
                 edge_representation_model: Optional[AbstractNeuralModel] = None, max_nodes_per_graph: int = 80000,
                 max_graph_edges: int = 100000, introduce_backwards_edges: bool = True, add_self_edges: bool = True,
                 stop_extending_minibatch_after_num_nodes: int = 10000, edge_dropout_rate: float = 0.0):
        super().__init__()
        if edge_representation_model is not None:
            raise NotImplementedError("edge features (edge_feature_size > 0) are outside the gnn-mlp default path")
        if edge_dropout_rate:
            raise NotImplementedError("edge dropout is not part of the gnn-mlp path")
        self.__node_embedding_model = node_representation_model
        self.__message_passing_layers_creator = message_passing_layer_creator
        self.max_nodes_per_graph = max_nodes_per_graph
        self.max_graph_edges = max_graph_edges
        self.introduce_backwards_edges = introduce_backwards_edges
        self.add_self_edges = add_self_edges
        self.stop_extending_minibatch_after_num_nodes = stop_extending_minibatch_after_num_nodes
        self.__edge_types_mdata = set()
        self.__edge_types: Optional[Tuple[str, ...]] = None

    @property
    def node_representation_model(self) -> AbstractNeuralModel:
        return self.__node_embedding_model

    @property
    def edge_types(self) -> Tuple[str, ...]:
        return self.__edge_types

    @property
    def num_edge_types(self) -> int:
        """Edge types as seen by a message-passing layer: forward, backward and the self-edge type (P1)."""
        n = len(self.__edge_types)
        if self.introduce_backwards_edges:
            n *= 2
        if self.add_self_edges:
            n += 1
        return n

    # ---- metadata -------------------------------------------------------------------------------
    def update_metadata_from(self, datapoint: GraphData) -> None:
        model = self.__node_embedding_model
        if hasattr(model, "update_metadata_from_many"):
            model.update_metadata_from_many(datapoint.node_information)
        else:
            for node in datapoint.node_information:
                model.update_metadata_from(node)
        self.__edge_types_mdata.update(datapoint.edges.keys())

    def update_metadata_from_edge_types(self, edge_type_names) -> None:
        """Adds edge-type names collected elsewhere (buglab_b200.shards.NativeMetadataPass) to the metadata."""
        self.__edge_types_mdata.update(edge_type_names)

    def finalize_metadata(self) -> None:
        self.__edge_types = tuple(sorted(self.__edge_types_mdata))
        self.__edge_types_mdata = None
        LOGGER.info("Edge types (%d forward): %s", len(self.__edge_types), self.__edge_types)

    def build_neural_module(self) -> GraphNeuralNetwork:
        layers = self.__message_passing_layers_creator(self.num_edge_types)
        # (D_in, M) of every message-passing layer: finalize_minibatch picks the pair-table layout of the device plan from
        # them (it runs on the producer thread, without the module)
        self._mp_layer_dims = tuple((l.input_state_dimension, l.message_dimension) for l in layers
                                    if hasattr(l, "message_dimension"))
        return GraphNeuralNetwork(layers, node_embedder=self.__node_embedding_model.build_neural_module())

    # ---- tensorise one graph --------------------------------------------------------------------
    def tensorize(self, datapoint: GraphData) -> Optional[TensorizedGraphData]:
        num_nodes = len(datapoint.node_information)
        if num_nodes > self.max_nodes_per_graph:
            LOGGER.warning("Dropping graph with %s nodes.", num_nodes)
            return None
        empty = np.zeros(0, dtype=np.int32)
        forward = []
        for edge_type in self.__edge_types:
            adj = datapoint.edges.get(edge_type)
            if adj is None or len(adj) == 0:
                forward.append((empty, empty))
            else:
                adj = np.asarray(adj, dtype=np.int32)
                forward.append((np.ascontiguousarray(adj[:, 0]), np.ascontiguousarray(adj[:, 1])))
        node_model = self.__node_embedding_model
        if hasattr(node_model, "tensorize_many"):
            node_data = node_model.tensorize_many(datapoint.node_information)
        else:
            node_data = [node_model.tensorize(n) for n in datapoint.node_information]
        return self.tensorize_arrays(num_nodes, node_data, forward, datapoint.reference_nodes)

    def tensorize_arrays(self, num_nodes: int, node_data, forward: List[Tuple[np.ndarray, np.ndarray]],
                         reference_nodes: Dict[str, Any]) -> Optional[TensorizedGraphData]:
        """The per-graph tensors from already-packed parts: ``node_data`` as the node model tensorises it and one
        ``(sources, targets)`` int32 pair per forward edge type, in ``edge_types`` order.  Shared by ``tensorize`` and the
        native shard path (buglab_b200/shards.py), which produces these arrays without building Python objects."""
        if num_nodes > self.max_nodes_per_graph:
            LOGGER.warning("Dropping graph with %s nodes.", num_nodes)
            return None
        adjacency = list(forward)
        if self.introduce_backwards_edges:
            adjacency.extend((tgt, src) for src, tgt in forward)
        if self.add_self_edges:
            ar = np.arange(num_nodes, dtype=np.int32)
            adjacency.append((ar, ar))
        if sum(a[0].shape[0] for a in adjacency) > 2 ** 31 - 16:
            return None
        return TensorizedGraphData(
            num_nodes=num_nodes,
            node_tensorized_data=node_data,
            adjacency_lists=adjacency,
            reference_nodes={k: np.asarray(v, dtype=np.int32) for k, v in reference_nodes.items()},
        )

    # ---- minibatch packing ----------------------------------------------------------------------
    def initialize_minibatch(self) -> Dict[str, Any]:
        # Chunks are appended as they come (no per-graph arithmetic here); the node offset of every chunk is recorded and
        # applied once, vectorised, when the minibatch is finalised.
        return {
            "node_ids": [], "node_lens": [],
            "adjacency_lists": [([], [], []) for _ in range(self.num_edge_types)],  # (sources, targets, node offset) chunks
            "num_nodes_per_graph": [],
            "reference_node_ids": {},        # name -> per-graph arrays (graph-local node ids)
            "reference_node_graphs": {},     # name -> graph index of each of those arrays
            "num_nodes": 0,
        }

    def extend_minibatch_with(self, tensorized_datapoint: TensorizedGraphData, partial_minibatch: Dict[str, Any]) -> bool:
        mb = partial_minibatch
        offset = mb["num_nodes"]
        graph_idx = len(mb["num_nodes_per_graph"])
        ids, lens = tensorized_datapoint.node_tensorized_data
        mb["node_ids"].append(ids)
        mb["node_lens"].append(lens)
        for (srcs, tgts, offsets), (src, tgt) in zip(mb["adjacency_lists"], tensorized_datapoint.adjacency_lists):
            if src.shape[0]:
                srcs.append(src)
                tgts.append(tgt)
                offsets.append(offset)
        for name, nodes in tensorized_datapoint.reference_nodes.items():
            mb["reference_node_ids"].setdefault(name, []).append(nodes)
            mb["reference_node_graphs"].setdefault(name, []).append(graph_idx)
        mb["num_nodes_per_graph"].append(tensorized_datapoint.num_nodes)
        mb["num_nodes"] = offset + tensorized_datapoint.num_nodes
        return mb["num_nodes"] < self.stop_extending_minibatch_after_num_nodes

    def finalize_minibatch(self, accumulated_minibatch_data: Dict[str, Any], device: Union[str, torch.device]) -> Dict[str, Any]:
        mb = accumulated_minibatch_data
        device = torch.device(device)
        num_graphs = len(mb["num_nodes_per_graph"])
        nodes_per_graph = np.asarray(mb["num_nodes_per_graph"], dtype=np.int64)
        node_offsets = np.zeros(num_graphs, dtype=np.int32)
        if num_graphs > 1:
            np.cumsum(nodes_per_graph[:-1], out=node_offsets[1:])

        # Pass 1 lays every index table out in ONE staging buffer; pass 2 concatenates the chunks straight into it and adds
        # the node offsets in place (one vectorised add per table instead of one small numpy call per graph and table).
        layout: List[Tuple[int, Tuple[int, ...]]] = []
        fillers: List[Tuple[int, int, Any]] = []
        total = 0

        def reserve(shape: Tuple[int, ...], fill) -> int:
            nonlocal total
            n = int(np.prod(shape)) if len(shape) else 1
            layout.append((total, tuple(shape)))
            fillers.append((total, n, fill))
            total += n
            return len(layout) - 1

        def table(chunks: List[np.ndarray], trailing: Tuple[int, ...] = (), chunk_offsets=None) -> int:
            rows = sum(c.shape[0] for c in chunks)
            shape = (rows,) + tuple(trailing)

            def fill(flat: np.ndarray) -> None:
                if rows == 0:
                    return
                out = flat.reshape(shape)
                np.concatenate(chunks, out=out)
                if chunk_offsets is not None:
                    per_row = np.repeat(np.asarray(chunk_offsets, dtype=np.int32), [c.shape[0] for c in chunks])
                    out += per_row.reshape((-1,) + (1,) * len(trailing))

            return reserve(shape, fill)

        def repeated(values: np.ndarray, counts) -> int:
            counts = np.asarray(counts, dtype=np.int64)
            n = int(counts.sum())

            def fill(flat: np.ndarray) -> None:
                if n:
                    flat[:] = np.repeat(values, counts)

            return reserve((n,), fill)

        T = self.__node_embedding_model.max_num_subtokens
        h_ids = table(mb["node_ids"], (T,))
        h_lens = table(mb["node_lens"])
        h_adj = [(table(srcs, (), offsets), table(tgts, (), offsets)) for srcs, tgts, offsets in mb["adjacency_lists"]]
        h_n2g = repeated(np.arange(num_graphs, dtype=np.int32), nodes_per_graph)
        h_ref = {}
        for name, chunks in mb["reference_node_ids"].items():
            graphs = np.asarray(mb["reference_node_graphs"][name], dtype=np.int32)
            trailing = chunks[0].shape[1:] if chunks else ()
            h_ref[name] = (table(chunks, trailing, node_offsets[graphs]),
                           repeated(graphs, [c.shape[0] for c in chunks]))

        staging = torch.empty(max(total, 1), dtype=torch.int32, pin_memory=(device.type == "cuda"))
        host = staging.numpy()
        for start, n, fill in fillers:
            fill(host[start: start + n])
        on_device = staging.to(device, non_blocking=True)

        def view(handle: int) -> torch.Tensor:
            start, shape = layout[handle]
            n = int(np.prod(shape)) if len(shape) else 1
            return on_device[start: start + n].view(shape)

        adjacency = PlannedAdjacency((view(s), view(t)) for s, t in h_adj)
        adjacency.num_nodes = mb["num_nodes"]
        if device.type == "cuda":
            from buglab_b200 import ops

            adjacency.block_nodes = ops.plan_block_nodes_for(getattr(self, "_mp_layer_dims", ()))
            adjacency.plan = plan_for(adjacency, mb["num_nodes"])
        return {
            "node_data": {"token_idxs": view(h_ids), "lengths": view(h_lens)},
            "adjacency_lists": adjacency,
            "node_to_graph_idx": view(h_n2g).long(),
            "reference_node_ids": {name: view(h[0]).long() for name, h in h_ref.items()},
            "reference_node_graph_idx": {name: view(h[1]).long() for name, h in h_ref.items()},
            "num_graphs": num_graphs,
            "h2d_bytes": int(total) * 4,
        }
