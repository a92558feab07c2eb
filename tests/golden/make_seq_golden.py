#!/usr/bin/env python
"""Generates tests/golden/seq_layers.npz by running the REAL reference relational-transformer layers on the CPU.

The three layer files (buglab/models/layers/{multihead_attention,relational_multihead_attention,relational_transformer}.py)
depend on torch only, so they are imported from /root/reference as they are — no shims.  For every case the fixture holds
the seeded parameters, the inputs, the layer output and the gradients of a fixed scalar functional of the output with
respect to the input and every parameter.  Run in the build container (the reference is not present on the GPU box):

    python tests/golden/make_seq_golden.py
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, "/root/reference")
from buglab.models.layers.relational_transformer import RelationalTransformerEncoderLayer  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "seq_layers.npz")

# name -> (layer kwargs, B, L, E)
CASES = {
    "great_postnorm": (dict(), 3, 12, 40),                                                       # what seq-great runs
    "rat_postnorm": (dict(use_edge_value_biases=True), 3, 12, 40),                               # what seq-rat runs
    "scalar_bias": (dict(edge_attention_bias_is_scalar=True), 2, 10, 25),                        # the GREAT paper variant
    "scalar_bias_with_values": (dict(edge_attention_bias_is_scalar=True, use_edge_value_biases=True), 2, 10, 25),
    "prenorm_gelu": (dict(normalisation_mode="prenorm", activation="gelu"), 2, 9, 30),
    "nonorm_rezero_scalar": (dict(normalisation_mode="off", rezero_mode="scalar"), 2, 8, 16),
    "postnorm_rezero_vector": (dict(rezero_mode="vector", use_edge_value_biases=True), 2, 8, 16),
    "no_edges": (dict(use_edge_value_biases=True), 2, 7, 0),
    "single_sample_dense_edges": (dict(), 1, 6, 90),                                             # many duplicate (i, j) pairs
    "medium": (dict(use_edge_value_biases=True), 2, 64, 400),
}
D_MODEL, HEADS, TYPES, FF = {"medium": 64}, 4, 5, 64


def main():
    out = {}
    for idx, (name, (kwargs, B, L, E)) in enumerate(CASES.items()):
        d_model = D_MODEL.get(name, 32)
        torch.manual_seed(100 + idx)
        layer = RelationalTransformerEncoderLayer(d_model=d_model, key_query_dimension=d_model // HEADS,
                                                  value_dimension=d_model // HEADS, nhead=HEADS, num_edge_types=TYPES,
                                                  dim_feedforward=FF, dropout=0.0, **kwargs)
        with torch.no_grad():  # ReZero starts at 0 (a no-op layer); use trained-looking values instead
            for pname, p in layer.named_parameters():
                if "_alpha" in pname:
                    p.copy_(torch.rand_like(p) + 0.5)
        layer.eval()
        g = torch.Generator().manual_seed(7 + idx)
        lengths = torch.randint(max(2, L // 2), L + 1, (B,), generator=g)
        lengths[0] = L
        mask = torch.arange(L)[None, :] >= lengths[:, None]                     # True = padding
        src = torch.randn(B, L, d_model, generator=g)
        src = (src * (~mask)[..., None]).requires_grad_(True)                  # seqmodel.py:369 zeroes padding rows
        sample = torch.randint(0, B, (E,), generator=g)
        hi = lengths[sample].float()
        e_src = (torch.rand(E, generator=g) * hi).long()
        e_tgt = (torch.rand(E, generator=g) * hi).long()                        # self loops and duplicates included
        edges = torch.stack([sample, e_src, e_tgt], dim=1)
        edge_types = torch.randint(0, TYPES, (E,), generator=g)
        weights = torch.randn(B, L, d_model, generator=g) * (~mask)[..., None]  # only unmasked positions are ever read

        y = layer(src, mask, edges, edge_types)
        (y * weights).sum().backward()
        pre = f"{name}/"
        out[pre + "kwargs"] = np.array(repr(sorted(kwargs.items())))
        out[pre + "dims"] = np.array([B, L, E, d_model, HEADS, TYPES, FF])
        for k, v in layer.state_dict().items():
            out[pre + "param/" + k] = v.detach().numpy().copy()
        for k, p in layer.named_parameters():
            if p.grad is not None:
                out[pre + "grad/" + k] = p.grad.numpy().copy()
        out[pre + "src"], out[pre + "mask"] = src.detach().numpy().copy(), mask.numpy()
        out[pre + "edges"], out[pre + "edge_types"] = edges.numpy(), edge_types.numpy()
        out[pre + "weights"], out[pre + "out"] = weights.numpy(), y.detach().numpy().copy()
        out[pre + "grad_src"] = src.grad.numpy().copy()
    np.savez_compressed(OUT, **out)
    print(f"wrote {OUT}: {len(CASES)} cases, {os.path.getsize(OUT) / 1e3:.0f} kB")


if __name__ == "__main__":
    main()
