from typing import List

import torch.nn as nn


class MLP(nn.Module):
    """Linear/activation stack; the ``_layers`` name keeps the reference's state_dict keys (layers/mlp.py:6-20)."""

    def __init__(self, input_dim: int, out_dim: int, hidden_layer_dims: List[int], activation=nn.ReLU()):
        super().__init__()
        dims = [input_dim] + list(hidden_layer_dims)
        modules: List[nn.Module] = []
        for d_in, d_out in zip(dims[:-1], dims[1:]):
            modules += [nn.Linear(d_in, d_out), activation]
        modules.append(nn.Linear(dims[-1], out_dim))
        self._layers = nn.Sequential(*modules)

    def forward(self, inputs):
        return self._layers(inputs)
