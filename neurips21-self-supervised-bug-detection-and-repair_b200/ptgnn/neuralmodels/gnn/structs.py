from typing import Any, Dict, Generic, List, NamedTuple, Optional, Tuple, TypeVar

import numpy as np
import torch

TNodeData = TypeVar("TNodeData")
TTensorizedNodeData = TypeVar("TTensorizedNodeData")


class GraphData(NamedTuple):
    """One raw graph (constructed at buglab/representations/data.py:152-167)."""

    node_information: List[Any]
    edges: Dict[str, np.ndarray]  # edge type -> int32 [E_k, 2] (src, tgt)
    reference_nodes: Dict[str, Any]  # name -> node ids (1-D, or [P, 2] for pairs)
    edge_features: Optional[Dict[str, List[Any]]] = None


class TensorizedGraphData(NamedTuple):
    num_nodes: int
    node_tensorized_data: Any
    adjacency_lists: List[Tuple[np.ndarray, np.ndarray]]
    reference_nodes: Dict[str, np.ndarray]
    edge_features: Optional[List[Any]] = None


class GnnOutput(NamedTuple):
    """Fields read by the reference at buglab/models/gnn.py:128-139."""

    input_node_representations: torch.Tensor
    output_node_representations: torch.Tensor
    node_to_graph_idx: torch.Tensor
    node_idx_references: Dict[str, torch.Tensor]
    node_graph_idx_reference: Dict[str, torch.Tensor]
    num_graphs: int
