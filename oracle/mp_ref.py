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


class MlpMessagePassingLayer(nn.Module):
    """Class and attribute names mirror ptgnn's (name-mangled private attributes), so ``state_dict`` keys are
    identical to the product module's and weights move across with ``load_state_dict``."""

    def __init__(self, input_state_dimension: int, message_dimension: int, output_state_dimension: int,
                 num_edge_types: int, message_aggregation_function: str = "max", dropout_rate: float = 0.0,
                 features_dimension: int = 0, use_message_bias: bool = True, **_unused):
        super().__init__()
        assert features_dimension == 0, "edge features are outside the gnn-mlp default path (modelregistry.py:56)"
        self.__aggregation = message_aggregation_function
        self.__edge_message_transformation_layers = nn.ModuleList(
            [nn.Linear(2 * input_state_dimension, message_dimension, bias=use_message_bias) for _ in range(num_edge_types)]
        )
        dense = nn.Linear(message_dimension, output_state_dimension, bias=False)
        nn.init.xavier_uniform_(dense.weight)
        self.__state_update = nn.Sequential(nn.LayerNorm(message_dimension), dense, nn.Tanh(), nn.Dropout(p=dropout_rate))
        self.__output_state_dim = output_state_dimension
        self.__input_state_dim = input_state_dimension
        # test hook: [N, M] int64 winning edge per (node, channel) in the type-major concatenation (E = no edge).  When
        # set, the max-aggregation uses THIS routing instead of its own argmax, which makes gradients comparable
        # elementwise with an implementation whose near-tied winners differ (see oracle/parity.py).
        self.forced_winners = None
        # routing audit (filled whenever forced_winners is used): how the forced routing compares with THIS evaluation's
        # own exact segment max — lets a test verify another implementation's winners without trusting them
        self.routing_audit = None

    @property
    def output_state_dimension(self) -> int:
        return self.__output_state_dim

    @property
    def input_state_dimension(self) -> int:
        return self.__input_state_dim

    def aggregated_messages(self, node_states, adjacency_lists):
        layers = self.__edge_message_transformation_layers
        weight = torch.stack([l.weight for l in layers])
        bias = torch.stack([l.bias for l in layers]) if layers[0].bias is not None else None
        messages, targets = edge_messages_ref(node_states, adjacency_lists, weight, bias)
        N = node_states.shape[0]
        if self.__aggregation == "max" and self.forced_winners is not None:
            arg = self.forced_winners
            E = messages.shape[0]
            picked = messages.gather(0, arg.clamp(max=E - 1))
            forced = torch.where(arg >= E, torch.zeros_like(picked), picked)
            with torch.no_grad():
                own_max, own_arg = scatter_max(messages, targets, dim=0, dim_size=N)
                valid = arg < E
                node_of_winner = targets[arg.clamp(max=E - 1)]
                wrong_segment = valid & (node_of_winner != torch.arange(N).view(-1, 1))
                deficit = (own_max - forced) / (1.0 + own_max.abs())     # >= 0 up to rounding: the forced edge's message
                later_on_tie = valid & (own_arg < E) & (arg > own_arg) & (forced == own_max)
                self.routing_audit = dict(
                    decisions=int(arg.numel()), differing=int((arg != own_arg).sum()),
                    wrong_segment=int(wrong_segment.sum()), empty_mismatch=int(((arg >= E) != (own_arg >= E)).sum()),
                    max_relative_deficit=float(deficit.max()) if deficit.numel() else 0.0,
                    later_edge_on_exact_tie=int(later_on_tie.sum()))
            return forced
        if self.__aggregation == "max":
            return scatter_max(messages, targets, dim=0, dim_size=N)[0]
        if self.__aggregation == "min":
            return scatter_min(messages, targets, dim=0, dim_size=N)[0]
        if self.__aggregation == "sum":
            return scatter_sum(messages, targets, dim=0, dim_size=N)
        if self.__aggregation == "mean":
            return scatter_mean(messages, targets, dim=0, dim_size=N)
        raise ValueError(self.__aggregation)

    def forward(self, node_states, adjacency_lists, node_to_graph_idx=None, reference_node_ids=None,
                reference_node_graph_idx=None, edge_features=None):
        return self.__state_update(self.aggregated_messages(node_states, adjacency_lists))


MlpMessagePassingLayerRef = MlpMessagePassingLayer


class SubtokenUnitEmbedder(nn.Module):
    """ptgnn subtoken embedder, ``subtoken_combination='max'``: Embedding -> dropout -> masked max."""

    def __init__(self, vocabulary_size: int, embedding_size: int, dropout_rate: float, subtoken_combination: str = "max",
                 padding_idx: int = 0):
        super().__init__()
        assert subtoken_combination == "max"
        self.__embeddings = nn.Embedding(vocabulary_size, embedding_size, padding_idx=padding_idx)
        self.__dropout_layer = nn.Dropout(p=dropout_rate)

    @property
    def embedding_layer(self) -> nn.Embedding:
        return self.__embeddings

    def forward(self, token_idxs, lengths):
        emb = self.__dropout_layer(self.__embeddings(token_idxs.long()))  # [N, T, H]
        mask = torch.arange(token_idxs.shape[1]).view(1, -1) < lengths.long().view(-1, 1)
        return emb.masked_fill(~mask.unsqueeze(-1), -float("inf")).max(dim=1)[0]


def subtoken_maxpool_ref(embedding: torch.Tensor, ids: torch.Tensor, lens: torch.Tensor) -> torch.Tensor:
    """ptgnn subtoken embedder with ``subtoken_combination='max'``, dropout off:
    Embedding -> masked max over the first ``lens[n]`` subtokens."""
    emb = embedding[ids.long()]  # [N, T, H]
    T = ids.shape[1]
    mask = torch.arange(T).view(1, -1) < lens.long().view(-1, 1)
    emb = emb.masked_fill(~mask.unsqueeze(-1), -float("inf"))
    return emb.max(dim=1)[0]
