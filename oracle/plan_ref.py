"""Host (numpy) restatement of the typed-edge plan — oracle for ``bl_plan_build`` (bit-exact check).

Integer bookkeeping only.  Mirrors the table definitions in include/buglab_b200.h; the ordering rules
are the ones ptgnn's MlpMessagePassingLayer implies: edges are concatenated type-major
(``torch.cat`` over ``adjacency_lists``, reference call site buglab/models/gnnlayerdefs.py:6-23), and the
"original edge index" is the position in that concatenation.  Test infrastructure only.
"""
import numpy as np


def build_plan_ref(src: np.ndarray, tgt: np.ndarray, etype: np.ndarray, num_nodes: int, num_edge_types: int,
                   block_nodes: int = 0) -> dict:
    """``block_nodes`` = B > 0: pair tables ordered by (node // B, type, node), segment s = (block s // K, type s % K)."""
    src, tgt, etype = (np.asarray(a, dtype=np.int64) for a in (src, tgt, etype))
    E, N, K = src.shape[0], int(num_nodes), int(num_edge_types)
    assert np.all(np.diff(etype) >= 0), "input must be the type-major concatenation"
    perm = np.argsort(tgt, kind="stable")  # (tgt, type, original index) because the input is type-major
    e_src, e_tgt, e_type = src[perm], tgt[perm], etype[perm]
    row_ptr = np.searchsorted(e_tgt, np.arange(N + 1), side="left")

    B = int(block_nodes)
    S = ((N + B - 1) // B) * K if B > 0 else K

    def pairs(node_of_edge):
        if B > 0:
            keys = ((node_of_edge // B) * K + e_type) * B + node_of_edge % B
            ukeys, inv = np.unique(keys, return_inverse=True)  # (node block, type, node) order
            pair_node = (ukeys // B // K) * B + ukeys % B
            type_ptr = np.searchsorted(ukeys, np.arange(S + 1) * B, side="left")
        else:
            keys = e_type * N + node_of_edge
            ukeys, inv = np.unique(keys, return_inverse=True)  # sorted unique keys: (type, node) order
            pair_node = ukeys % N
            type_ptr = np.searchsorted(ukeys, np.arange(K + 1) * N, side="left")
        order = np.argsort(pair_node, kind="stable")
        by_node_ptr = np.searchsorted(pair_node[order], np.arange(N + 1), side="left")
        return inv.reshape(-1), pair_node, type_ptr, by_node_ptr, order

    urow, s_node, s_type_ptr, s_by_node_ptr, s_by_node_idx = pairs(e_src)
    vrow, t_node, t_type_ptr, t_by_node_ptr, t_by_node_idx = pairs(e_tgt)
    out = dict(
        e_perm=perm, e_src=e_src, e_type=e_type, row_ptr=row_ptr, urow=urow, vrow=vrow,
        s_node=s_node, s_type_ptr=s_type_ptr, s_by_node_ptr=s_by_node_ptr, s_by_node_idx=s_by_node_idx,
        t_node=t_node, t_type_ptr=t_type_ptr, t_by_node_ptr=t_by_node_ptr, t_by_node_idx=t_by_node_idx,
    )
    # S-pair -> its sorted edges (ascending) and the target of every sorted edge (by-source half of the edge backward)
    by_pair = np.argsort(urow, kind="stable")
    out["s_edge_idx"] = by_pair
    out["s_edge_ptr"] = np.searchsorted(urow[by_pair], np.arange(s_node.shape[0] + 1), side="left")
    out["e_tgt"] = e_tgt
    out["s_edge_tgt"] = e_tgt[by_pair]
    out = {k: np.asarray(v, dtype=np.int32) for k, v in out.items()}
    out["num_s_pairs"], out["num_t_pairs"] = int(s_node.shape[0]), int(t_node.shape[0])
    return out
