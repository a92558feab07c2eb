"""ptgnn message passing restated in plain PyTorch (per-edge formulation) — oracle, test infrastructure only.

PARITY UNPINNED: ``ptgnn`` is an unpinned third-party dependency of the reference (requirements.txt:13),
absent from /root/reference and not installable offline; no reference test pins it.  This follows the
semantics SURVEY.md §8a states for P3/P4/P5 (ptgnn ~0.8-0.10, torch_scatter 2.0.x):

  MlpMessagePassingLayer.forward, as configured by buglab/models/gnnlayerdefs.py:6-23:
      for each edge type k:  m_k = Linear_k(cat[h[src_k], h[tgt_k]])        (2*D_in -> M, with bias)
      m   = GELU(cat_k m_k)                                                 (exact erf GELU)
      agg = torch_scatter.scatter_max(m, cat_k tgt_k, dim=0, dim_size=N)[0] (empty -> 0, first max wins)
      out = Dropout(Tanh(Linear(M -> D_out, no bias, xavier)(LayerNorm(M)(agg))))
"""
from typing import List, Optional, Sequence, Tuple

import torch
from torch import nn

from .scatter_ref import scatter_max, scatter_mean, scatter_min, scatter_sum


def edge_messages_ref(h: torch.Tensor, adjacency_lists: Sequence[Tuple[torch.Tensor, torch.Tensor]],
                      weight: torch.Tensor, bias: Optional[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """All per-edge messages GELU(W_k [h_s;h_t] + b_k) in the type-major concatenation order, and their targets."""
    messages, targets = [], []
    for k, (src, tgt) in enumerate(adjacency_lists):
        src, tgt = src.long(), tgt.long()
        message_input = torch.cat([h[src], h[tgt]], dim=-1)
        m = torch.nn.functional.linear(message_input, weight[k], None if bias is None else bias[k])
        messages.append(m)
        targets.append(tgt)
    messages = torch.nn.functional.gelu(torch.cat(messages, dim=0))
    return messages, torch.cat(targets)


def typed_edge_message_max_ref(h, adjacency_lists, weight, bias):
    messages, targets = edge_messages_ref(h, adjacency_lists, weight, bias)
    return scatter_max(messages, targets, dim=0, dim_size=h.shape[0])


class MlpMessagePassingLayerRef(nn.Module):
    """Parameter names mirror the product module so ``load_state_dict`` moves weights across."""

    def __init__(self, input_state_dimension: int, message_dimension: int, output_state_dimension: int,
                 num_edge_types: int, message_aggregation_function: str = "max", dropout_rate: float = 0.0,
                 features_dimension: int = 0, use_message_bias: bool = True):
        super().__init__()
        assert features_dimension == 0, "edge features are outside the gnn-mlp default path (modelregistry.py:56)"
        self.aggregation = message_aggregation_function
        self.edge_message_transformation_layers = nn.ModuleList(
            [nn.Linear(2 * input_state_dimension, message_dimension, bias=use_message_bias) for _ in range(num_edge_types)]
        )
        self.state_update_norm = nn.LayerNorm(message_dimension)
        self.state_update_dense = nn.Linear(message_dimension, output_state_dimension, bias=False)
        nn.init.xavier_uniform_(self.state_update_dense.weight)
        self.dropout = nn.Dropout(dropout_rate)
        self.output_state_dimension = output_state_dimension

    def forward(self, node_states, adjacency_lists, **_):
        weight = torch.stack([l.weight for l in self.edge_message_transformation_layers])
        bias = None
        if self.edge_message_transformation_layers[0].bias is not None:
            bias = torch.stack([l.bias for l in self.edge_message_transformation_layers])
        messages, targets = edge_messages_ref(node_states, adjacency_lists, weight, bias)
        N = node_states.shape[0]
        if self.aggregation == "max":
            agg = scatter_max(messages, targets, dim=0, dim_size=N)[0]
        elif self.aggregation == "min":
            agg = scatter_min(messages, targets, dim=0, dim_size=N)[0]
        elif self.aggregation == "sum":
            agg = scatter_sum(messages, targets, dim=0, dim_size=N)
        elif self.aggregation == "mean":
            agg = scatter_mean(messages, targets, dim=0, dim_size=N)
        else:
            raise ValueError(self.aggregation)
        return self.dropout(torch.tanh(self.state_update_dense(self.state_update_norm(agg))))


def subtoken_maxpool_ref(embedding: torch.Tensor, ids: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """ptgnn subtoken embedder with ``subtoken_combination='max'``, dropout off:
    Embedding -> masked max over the first ``lens[n]`` subtokens."""
    emb = embedding[ids.long()]  # [N, T, H]
    T = ids.shape[1]
    mask = torch.arange(T).view(1, -1) < lens.long().view(-1, 1)
    emb = emb.masked_fill(~mask.unsqueeze(-1), -float("inf"))
    return emb.max(dim=1)[0]
