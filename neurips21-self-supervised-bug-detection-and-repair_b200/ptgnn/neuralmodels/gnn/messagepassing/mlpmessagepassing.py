"""``MlpMessagePassingLayer`` on the B200 kernels.

Constructor keywords are the ones buglab/models/gnnlayerdefs.py:6-23 passes.  Semantics:
    m_k = Linear_k(cat[h[src_k], h[tgt_k]])      per edge type k (2*D_in -> M)
    agg = scatter_max(GELU(cat_k m_k), cat_k tgt_k, dim_size=N)       (empty -> 0)
    out = Dropout(Tanh(Linear(M -> D_out, no bias, xavier)(LayerNorm(M)(agg))))
Execution: ``Linear_k([h_s;h_t]) = A_k h_s + B_k h_t + b_k`` is hoisted to the unique (type, node) pairs
(per-type fp32 GEMMs), the per-edge gather-add-GELU-segmented-max is ONE fused kernel
(``bl_edge_segmax_fwd``), LayerNorm and Tanh+Dropout are row kernels around one library GEMM.
Parameters are held by ordinary ``nn.Linear`` / ``nn.LayerNorm`` sub-modules so ``state_dict`` keys and
pickles keep the ptgnn layout; the kernels only read their ``.weight`` / ``.bias``.
"""
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from .abstractmessagepassing import AbstractMessagePassingLayer, plan_for


class MlpMessagePassingLayer(AbstractMessagePassingLayer):
    def __init__(self, input_state_dimension: int, message_dimension: int, output_state_dimension: int,
                 num_edge_types: int, message_aggregation_function: str, message_activation: Optional[nn.Module] = None,
                 use_target_state_as_message_input: bool = True, use_layer_norm: bool = True,
                 dense_activation: Optional[nn.Module] = None, dropout_rate: float = 0.0, features_dimension: int = 0,
                 use_message_bias: bool = True):
        super().__init__()
        if message_aggregation_function != "max":
            raise NotImplementedError("the fused B200 edge kernel implements the gnn-mlp aggregation ('max', "
                                      "gnnlayerdefs.py:11,20); other aggregations are not built")
        if features_dimension != 0:
            raise NotImplementedError("edge features (edge_feature_size > 0) are outside the gnn-mlp default path")
        if not (use_target_state_as_message_input and use_layer_norm) or message_activation is not None \
                or dense_activation is not None:
            raise NotImplementedError("only the gnn-mlp configuration (GELU messages, LayerNorm, Tanh) is built")
        self.__input_state_dim = input_state_dimension
        self.__output_state_dim = output_state_dimension
        self.__message_dim = message_dimension
        self.__dropout_rate = float(dropout_rate)
        self.__edge_message_transformation_layers = nn.ModuleList(
            [nn.Linear(2 * input_state_dimension, message_dimension, bias=use_message_bias) for _ in range(num_edge_types)]
        )
        dense = nn.Linear(message_dimension, output_state_dimension, bias=False)
        nn.init.xavier_uniform_(dense.weight)
        self.__state_update = nn.Sequential(nn.LayerNorm(message_dimension), dense, nn.Tanh(), nn.Dropout(p=dropout_rate))

    @property
    def input_state_dimension(self) -> int:
        return self.__input_state_dim

    @property
    def output_state_dimension(self) -> int:
        return self.__output_state_dim

    @property
    def message_dimension(self) -> int:
        return self.__message_dim

    def stacked_message_parameters(self) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        layers = self.__edge_message_transformation_layers
        weight = torch.stack([l.weight for l in layers])  # [K, M, 2*D_in]
        bias = torch.stack([l.bias for l in layers]) if layers[0].bias is not None else None
        return weight, bias

    def forward(self, node_states: torch.Tensor, adjacency_lists, node_to_graph_idx=None, reference_node_ids=None,
                reference_node_graph_idx=None, edge_features=None) -> torch.Tensor:
        from buglab_b200 import ops

        plan = plan_for(adjacency_lists, node_states.shape[0])
        weight, bias = self.stacked_message_parameters()
        aggregated = ops.typed_edge_message_max(node_states, weight, bias, plan)  # [N, M]
        norm, dense = self.__state_update[0], self.__state_update[1]
        normed = ops.layer_norm(aggregated, norm.weight, norm.bias, norm.eps)
        updated = ops.dense_linear(normed, dense.weight)  # split-fp16 tensor-core GEMM (fp32-class accuracy)
        return ops.tanh_dropout(updated, self.__dropout_rate, self.training)
