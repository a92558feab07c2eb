from abc import abstractmethod
from typing import Dict, List, Tuple

import torch
from torch import nn


class PlannedAdjacency(list):
    """``adjacency_lists`` (a list of ``(src, tgt)`` tensors) that carries the per-minibatch device plan
    (``buglab_b200.ops.EdgePlan``) so the 8 message-passing layers share one CSR / pair-table build."""

    plan = None
    num_nodes = None
    block_nodes = 0  # pair-table layout the model's layers can consume (ops.plan_block_nodes_for); 0 = type-major


def plan_for(adjacency_lists, num_nodes: int, block_nodes=None):
    from buglab_b200 import ops

    plan = getattr(adjacency_lists, "plan", None)
    if plan is None or plan.num_nodes != num_nodes:
        if block_nodes is None:
            block_nodes = getattr(adjacency_lists, "block_nodes", 0)
        plan = ops.build_edge_plan(adjacency_lists, num_nodes, block_nodes)
        if isinstance(adjacency_lists, PlannedAdjacency):
            adjacency_lists.plan = plan
    return plan


class AbstractMessagePassingLayer(nn.Module):
    @abstractmethod
    def forward(self, node_states: torch.Tensor, adjacency_lists: List[Tuple[torch.Tensor, torch.Tensor]],
                node_to_graph_idx: torch.Tensor, reference_node_ids: Dict[str, torch.Tensor],
                reference_node_graph_idx: Dict[str, torch.Tensor], edge_features: List[torch.Tensor]) -> torch.Tensor:
        ...

    @property
    @abstractmethod
    def input_state_dimension(self) -> int:
        ...

    @property
    @abstractmethod
    def output_state_dimension(self) -> int:
        ...
