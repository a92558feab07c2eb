"""Concat-residual pair used by buglab/models/gnnlayerdefs.py:24-39: the dummy layer remembers the state that
enters a block, the residual layer returns ``cat(remembered, current)``."""
from typing import Optional

import torch

from .abstractmessagepassing import AbstractMessagePassingLayer


class _PassThroughDummyLayer(AbstractMessagePassingLayer):
    def __init__(self, residual_layer: "ConcatResidualLayer", state_dimension: int):
        super().__init__()
        # not registered as a sub-module (it is the same object that appears later in the layer list)
        object.__setattr__(self, "_residual_layer", residual_layer)
        self.__state_dimension = state_dimension

    def forward(self, node_states, adjacency_lists=None, node_to_graph_idx=None, reference_node_ids=None,
                reference_node_graph_idx=None, edge_features=None):
        self._residual_layer._remember(node_states)
        return node_states

    @property
    def input_state_dimension(self) -> int:
        return self.__state_dimension

    @property
    def output_state_dimension(self) -> int:
        return self.__state_dimension


class ConcatResidualLayer(AbstractMessagePassingLayer):
    def __init__(self, input_state_dimension: int):
        super().__init__()
        self.__input_dim = input_state_dimension
        self.__saved: Optional[torch.Tensor] = None

    def pass_through_dummy_layer(self) -> _PassThroughDummyLayer:
        return _PassThroughDummyLayer(self, self.__input_dim)

    def _remember(self, node_states: torch.Tensor) -> None:
        self.__saved = node_states

    def forward(self, node_states, adjacency_lists=None, node_to_graph_idx=None, reference_node_ids=None,
                reference_node_graph_idx=None, edge_features=None):
        saved, self.__saved = self.__saved, None
        if saved is None:
            raise RuntimeError("ConcatResidualLayer used before its pass_through_dummy_layer()")
        return torch.cat((saved, node_states), dim=-1)

    def __getstate__(self):
        state = dict(self.__dict__)
        state["_ConcatResidualLayer__saved"] = None  # never pickle activations
        return state

    @property
    def input_state_dimension(self) -> int:
        return self.__input_dim

    @property
    def output_state_dimension(self) -> int:
        return 2 * self.__input_dim
