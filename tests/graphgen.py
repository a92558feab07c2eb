"""Seeded random typed-edge graphs for the kernel tests (test helper)."""
import numpy as np
import torch


def random_adjacency(num_nodes: int, num_edge_types: int, edges_per_type, seed: int, self_edges: bool = True,
                     hub: bool = True, duplicates: bool = True):
    """Returns a list of (src, tgt) int64 tensors per type.  Includes empty types, duplicate edges, a hub
    target with many in-edges and (optionally) a final self-edge type, and leaves some nodes isolated when
    ``self_edges`` is False."""
    rng = np.random.default_rng(seed)
    adj = []
    n_rand = num_edge_types - (1 if self_edges else 0)
    for k in range(n_rand):
        e = int(edges_per_type[k % len(edges_per_type)])
        src = rng.integers(0, num_nodes, size=e)
        # 80 % local targets, 20 % uniform
        local = np.clip(src + rng.integers(-8, 9, size=e), 0, num_nodes - 1)
        tgt = np.where(rng.random(e) < 0.8, local, rng.integers(0, num_nodes, size=e))
        if not self_edges and num_nodes > 4:
            tgt = np.where(tgt >= num_nodes - 2, 0, tgt)  # last two nodes stay isolated
        if hub and e > 8 and k == 0:
            tgt[: e // 4] = num_nodes // 2
        if duplicates and e > 4:
            src[-2:] = src[:2]
            tgt[-2:] = tgt[:2]
        adj.append((torch.from_numpy(src.astype(np.int64)), torch.from_numpy(tgt.astype(np.int64))))
    if self_edges:
        ar = torch.arange(num_nodes, dtype=torch.int64)
        adj.append((ar, ar.clone()))
    return adj


def flatten(adj):
    src = torch.cat([a[0] for a in adj]).to(torch.int32)
    tgt = torch.cat([a[1] for a in adj]).to(torch.int32)
    etype = torch.cat([torch.full((a[0].shape[0],), k, dtype=torch.int32) for k, a in enumerate(adj)])
    return src, tgt, etype
