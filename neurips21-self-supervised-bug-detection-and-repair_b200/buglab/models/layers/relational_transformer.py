"""Relational transformer encoder layer of the sequence models.

Parameter names and semantics: reference buglab/models/layers/multihead_attention.py:7-85,
relational_multihead_attention.py:7-178, relational_transformer.py:18-125 (a reference ``state_dict`` loads unchanged).
What runs where: the score / softmax / weighted-sum core with the typed-edge terms is the hand-written kernel behind
``ops.seq_edge_attention`` (tensor-core path: ``bl_tma_project`` / ``bl_tma_weight_grad`` + ``bl_seq_softmax_*``; CUDA-core
path ``bl_seq_attention_fwd/_bwd``), LayerNorm is ``bl_layernorm_*``; the dense projections and the feed-forward block run
on the same split-fp16 TMA-fed tcgen05 GEMMs (``ops.dense_linear``) where they cover the shape (config 4: all of them),
else on plain library GEMMs (``torch.nn.functional.linear``).  Behaviours kept on purpose: "great" uses
the query-side vector bias (the scalar switch is never set upstream) and post-norm applies ``norm1`` after both sub-layers.
"""
from typing import Optional, Union

import torch
from torch import nn

from buglab_b200 import ops


def _linear(lin: nn.Linear, x: torch.Tensor) -> torch.Tensor:
    """``lin(x)`` for x [..., k] on the split-fp16 TMA-fed tcgen05 GEMMs (``ops.dense_linear``: fp32-class accuracy, forward
    and both backward products) where they cover the shape, else on the fp32 library GEMM."""
    y = ops.dense_linear(x.reshape(-1, x.shape[-1]), lin.weight)
    if lin.bias is not None:
        y = y + lin.bias
    return y.view(*x.shape[:-1], lin.weight.shape[0])


def _layer_norm(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    shape = x.shape
    return ops.layer_norm(x.reshape(-1, shape[-1]), norm.weight, norm.bias, norm.eps).view(shape)


class MultiheadAttention(nn.Module):
    def __init__(self, *, num_heads: int, input_state_dimension: int, key_query_dimension: int, value_dimension: int,
                 output_dimension: int, dropout_rate: float):
        super().__init__()
        self._dropout_rate = float(dropout_rate)  # on the attention probabilities, inside the attention kernel
        self._num_heads, self._key_query_dim, self._value_dim = num_heads, key_query_dimension, value_dimension
        self._selfatt_head_transforms = nn.Linear(input_state_dimension,
                                                  num_heads * (2 * key_query_dimension + value_dimension), bias=False)
        self._out_proj = nn.Linear(value_dimension * num_heads, output_dimension, bias=False)
        self._scaling = key_query_dimension ** -0.5

    def _project(self, x: torch.Tensor):
        B, L, _ = x.shape
        per_head = _linear(self._selfatt_head_transforms, x).view(B, L, self._num_heads, -1).permute(0, 2, 1, 3)
        dk = self._key_query_dim
        return ((per_head[..., :dk] * self._scaling).contiguous(), per_head[..., dk: 2 * dk].contiguous(),
                per_head[..., 2 * dk:].contiguous())

    def _merge(self, per_head_values: torch.Tensor) -> torch.Tensor:
        B, H, L, dv = per_head_values.shape
        return _linear(self._out_proj, per_head_values.permute(0, 2, 1, 3).reshape(B, L, H * dv))


class RelationalMultiheadAttention(MultiheadAttention):
    def __init__(self, *, num_heads: int, num_edge_types: int, input_state_dimension: int, key_query_dimension: int,
                 value_dimension: int, output_dimension: int, dropout_rate: float, use_edge_value_biases: bool = False,
                 edge_attention_bias_is_scalar: bool = False):
        super().__init__(num_heads=num_heads, input_state_dimension=input_state_dimension,
                         key_query_dimension=key_query_dimension, value_dimension=value_dimension,
                         output_dimension=output_dimension, dropout_rate=dropout_rate)
        if edge_attention_bias_is_scalar:
            raise NotImplementedError("scalar edge biases are never built by the model registry (seqmodel.py:93-107)")
        if key_query_dimension != value_dimension:
            raise NotImplementedError("the attention kernel assumes key and value head sizes are equal, as the registry sets them")
        self._num_edge_types = num_edge_types
        self._use_edge_value_biases = use_edge_value_biases
        self._edge_attention_biases = nn.Embedding(num_edge_types, num_heads * key_query_dimension)
        self._reverse_edge_attention_biases = nn.Embedding(num_edge_types, num_heads * key_query_dimension)
        if use_edge_value_biases:
            self._edge_value_biases = nn.Embedding(num_edge_types, num_heads * value_dimension)
            self._reverse_edge_value_biases = nn.Embedding(num_edge_types, num_heads * value_dimension)

    def forward(self, input_seq_states, masked_elements, edges, edge_types=None):
        """``edges``: an ``ops.SeqAttentionPlan`` (built once per minibatch by the module) or the reference's ``[E, 3]``."""
        B, L, _ = input_seq_states.shape
        if isinstance(edges, ops.SeqAttentionPlan):
            plan = edges
        else:
            lengths = (~masked_elements).sum(dim=1) if masked_elements is not None else \
                torch.full((B,), L, device=input_seq_states.device)
            plan = ops.build_seq_attention_plan(edges, edge_types, lengths, L, self._num_edge_types)
        q, k, v = self._project(input_seq_states)
        H = self._num_heads
        bias = torch.cat((self._edge_attention_biases.weight, self._reverse_edge_attention_biases.weight)
                         ).view(-1, H, self._key_query_dim)
        vbias = None
        if self._use_edge_value_biases:
            vbias = torch.cat((self._edge_value_biases.weight, self._reverse_edge_value_biases.weight)
                              ).view(-1, H, self._value_dim)
        return self._merge(ops.seq_edge_attention(q, k, v, bias, vbias, plan, self._dropout_rate, self.training))


class RelationalTransformerEncoderLayer(nn.Module):
    def __init__(self, d_model: int, key_query_dimension: int, value_dimension: int, nhead: int, num_edge_types: int,
                 dim_feedforward: int = 2048, dropout: float = 0.1, activation: str = "relu",
                 use_edge_value_biases: bool = False, edge_attention_bias_is_scalar: bool = False,
                 rezero_mode: str = "off", normalisation_mode: str = "postnorm"):
        super().__init__()
        if activation not in ("relu", "gelu"):
            raise RuntimeError("activation should be relu/gelu, not {}".format(activation))
        if normalisation_mode not in ("off", "prenorm", "postnorm"):
            raise ValueError(f"Unrecognized normalization mode `{normalisation_mode}`.")
        if rezero_mode not in ("off", "scalar", "vector"):
            raise ValueError(f"Unrecognized rezero mode `{rezero_mode}`.")
        self.self_attn = RelationalMultiheadAttention(
            input_state_dimension=d_model, num_heads=nhead, output_dimension=d_model, dropout_rate=dropout,
            num_edge_types=num_edge_types, key_query_dimension=key_query_dimension, value_dimension=value_dimension,
            use_edge_value_biases=use_edge_value_biases, edge_attention_bias_is_scalar=edge_attention_bias_is_scalar)
        self.linear1 = nn.Linear(d_model, dim_feedforward)
        self.dropout = nn.Dropout(dropout)
        self.linear2 = nn.Linear(dim_feedforward, d_model)
        self._normalisation_mode = normalisation_mode
        self.norm1: Optional[nn.LayerNorm] = nn.LayerNorm(d_model) if normalisation_mode != "off" else None
        self.norm2: Optional[nn.LayerNorm] = nn.LayerNorm(d_model) if normalisation_mode != "off" else None
        self.dropout1, self.dropout2 = nn.Dropout(dropout), nn.Dropout(dropout)
        self._activation_name = activation
        self._rezero_mode = rezero_mode
        if rezero_mode == "scalar":
            self._alpha1: Union[float, torch.Tensor] = nn.Parameter(torch.tensor(0.0))
            self._alpha2: Union[float, torch.Tensor] = nn.Parameter(torch.tensor(0.0))
        elif rezero_mode == "vector":
            self._alpha1 = nn.Parameter(torch.zeros(d_model))
            self._alpha2 = nn.Parameter(torch.zeros(d_model))
        else:
            self._alpha1 = self._alpha2 = 1.0

    def forward(self, src, src_mask, edges, edge_types=None):
        pre, post = self._normalisation_mode == "prenorm", self._normalisation_mode == "postnorm"
        activation = nn.functional.relu if self._activation_name == "relu" else nn.functional.gelu
        attended = self.self_attn(_layer_norm(self.norm1, src) if pre else src, src_mask, edges, edge_types)
        src = src + self.dropout1(self._alpha1 * attended)
        if post:
            src = _layer_norm(self.norm1, src)
        hidden = self.dropout(activation(_linear(self.linear1, _layer_norm(self.norm2, src) if pre else src)))
        src = src + self.dropout2(self._alpha2 * _linear(self.linear2, hidden))
        if post:
            src = _layer_norm(self.norm1, src)  # sic — the reference normalises with norm1 again (relational_transformer.py:122-123)
        return src
