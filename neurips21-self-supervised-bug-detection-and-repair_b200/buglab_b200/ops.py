"""PyTorch-facing operators over the buglab_b200 C ABI (``include/buglab_b200.h``).

PyTorch is plumbing here: it owns device memory, streams and the autograd tape; every arithmetic step of
the gnn-mlp hot path runs in the kernels of ``csrc/``.  The per-type projections, their backward products and the
node-update Linear are split-fp16 ("f16x3") tensor-core GEMMs with fp32-class accuracy.  Dispatch, in order:

* the TMA-fed tcgen05 family of ``csrc/gemm_tma.cu`` (``bl_tma_project`` / ``bl_tma_weight_grad``: operands split once
  per table, CTA pairs) for every shape it covers — all of the registry's hidden-256 model, forward and backward, and the
  forward / backward-input products of hidden 128;
* the first-generation tcgen05 kernels of ``csrc/pair_project_tc.cu`` (gather + split inside the loader) for widths
  <= ``TC_MAX_WIDTH`` when ``BUGLAB_B200_TMA=0``, and split kernels + ``cublasGemmEx`` for the widths neither covers
  (hidden 128's message weight gradient);
* ``PROJECTION_MODE = "fp32"`` (plain cuBLAS SGEMM through ``torch.mm``, TF32 disabled) is kept as the exact referee for
  tests and error budgeting.

Reference semantics replaced: P4/P5 ``MlpMessagePassingLayer`` message+aggregate,
A9/A10 scatter ops (``buglab/models/utils.py:15-48``), P2 subtoken max-pool, A12 optimiser; §8(f) row 2: the relational
attention of the sequence models.
"""
import ctypes
import os
from typing import List, NamedTuple, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import check, f32, i32, stream_ptr


# ---------------------------------------------------------------------------------------------------
# Typed-edge plan
# ---------------------------------------------------------------------------------------------------
class EdgePlan(NamedTuple):
    """Device tables built once per minibatch by ``bl_plan_build`` (see the header for the layout)."""

    num_nodes: int
    num_edges: int
    num_edge_types: int
    e_perm: torch.Tensor
    e_src: torch.Tensor
    e_type: torch.Tensor
    row_ptr: torch.Tensor
    urow: torch.Tensor
    vrow: torch.Tensor
    s_node: torch.Tensor
    s_type_ptr: torch.Tensor
    s_by_node_ptr: torch.Tensor
    s_by_node_idx: torch.Tensor
    t_node: torch.Tensor
    t_type_ptr: torch.Tensor
    t_by_node_ptr: torch.Tensor
    t_by_node_idx: torch.Tensor
    num_s_pairs: int
    num_t_pairs: int
    s_type_ptr_host: Tuple[int, ...]
    t_type_ptr_host: Tuple[int, ...]
    # work-unit tables of the TMA GEMMs over the S- and T-pair segments (``Units``: tiles of tma_tile_rows() rows for the
    # projections, slabs of tma_slab_rows() rows for the weight gradient); None when the TMA path is off
    s_tiles: Optional["Units"] = None
    t_tiles: Optional["Units"] = None
    s_slabs: Optional["Units"] = None
    t_slabs: Optional["Units"] = None
    # S-pair -> its sorted edges (CSR) and the target node of every sorted edge: by-source half of the edge backward
    s_edge_ptr: Optional[torch.Tensor] = None
    s_edge_idx: Optional[torch.Tensor] = None
    e_tgt: Optional[torch.Tensor] = None
    s_edge_tgt: Optional[torch.Tensor] = None
    # pair layout: 0 = type-major (one segment per type); B > 0 = node-blocked, segments (node block, type) with
    # seg_type[s] = s % K (then s_type_ptr / t_type_ptr hold num_segs + 1 segment pointers and the *_host tuples are empty)
    block_nodes: int = 0
    num_segs: int = 0
    seg_type: Optional[torch.Tensor] = None


def build_edge_plan(
    adjacency_lists: Sequence[Tuple[torch.Tensor, torch.Tensor]], num_nodes: int, block_nodes: int = 0
) -> EdgePlan:
    """Build the plan from ptgnn-style adjacency lists ``[(src_k, tgt_k)]`` (one per edge type, any int dtype).

    The per-type pair counts come back to the host in one small D2H copy (they size the projection GEMMs).
    """
    lib = _lib.load()
    K = len(adjacency_lists)
    if K == 0:
        raise ValueError("need at least one edge type")
    device = adjacency_lists[0][0].device
    if device.type != "cuda":
        raise _lib.BuglabB200Error("build_edge_plan needs CUDA tensors; there is no CPU fallback")
    sizes = [int(a[0].shape[0]) for a in adjacency_lists]
    E = sum(sizes)
    src = torch.cat([a[0].reshape(-1) for a in adjacency_lists]).to(torch.int32)
    tgt = torch.cat([a[1].reshape(-1) for a in adjacency_lists]).to(torch.int32)
    # per-type constant runs written by fill kernels: building the counts tensor from the host list would be a
    # synchronising H2D copy in the middle of the step
    etype = torch.empty(E, device=device, dtype=torch.int32)
    start = 0
    for k, n in enumerate(sizes):
        if n:
            etype[start:start + n].fill_(k)
            start += n
    return build_edge_plan_from_flat(src, tgt, etype, num_nodes, K, block_nodes)


def build_edge_plan_from_flat(
    src: torch.Tensor, tgt: torch.Tensor, etype: torch.Tensor, num_nodes: int, num_edge_types: int, block_nodes: int = 0
) -> EdgePlan:
    """Same, from the type-major concatenation (int32 CUDA tensors).  ``block_nodes`` > 0 orders the pair tables by
    (node block, type, node) — only for models whose every layer runs on the segment-aware TMA GEMMs
    (:func:`plan_block_nodes_for`)."""
    lib = _lib.load()
    device = src.device
    E, N, K = int(src.shape[0]), int(num_nodes), int(num_edge_types)
    B = int(block_nodes) if USE_TMA else 0
    S = ((N + B - 1) // B) * K if B > 0 else K
    opts = dict(device=device, dtype=torch.int32)
    Ea = max(E, 1)
    e_perm, e_src, e_type = (torch.empty(Ea, **opts) for _ in range(3))
    urow, vrow = torch.empty(Ea, **opts), torch.empty(Ea, **opts)
    row_ptr = torch.empty(N + 1, **opts)
    s_node, s_by_node_idx = torch.empty(Ea, **opts), torch.empty(Ea, **opts)
    t_node, t_by_node_idx = torch.empty(Ea, **opts), torch.empty(Ea, **opts)
    s_by_node_ptr, t_by_node_ptr = torch.empty(N + 1, **opts), torch.empty(N + 1, **opts)
    # [counts (2) | s_type_ptr (S+1) | t_type_ptr (S+1)] in one buffer -> one D2H copy (of the counts only when blocked)
    meta = torch.empty(2 + 2 * (S + 1), **opts)
    counts, s_type_ptr, t_type_ptr = meta[:2], meta[2 : S + 3], meta[S + 3 :]
    s_edge_ptr, s_edge_idx, e_tgt, s_edge_tgt = (torch.empty(Ea + 1, **opts), torch.empty(Ea, **opts), torch.empty(Ea, **opts),
                                                 torch.empty(Ea, **opts))
    ws_bytes = lib.bl_plan_workspace_bytes(E, N, K)
    workspace = torch.empty(ws_bytes, device=device, dtype=torch.uint8)
    check(
        lib.bl_plan_build(
            i32(src.contiguous()), i32(tgt.contiguous()), i32(etype.contiguous()), E, N, K,
            i32(e_perm), i32(e_src), i32(e_type), i32(row_ptr), i32(urow), i32(vrow),
            i32(s_node), s_type_ptr.data_ptr(), i32(s_by_node_ptr), i32(s_by_node_idx),
            i32(t_node), t_type_ptr.data_ptr(), i32(t_by_node_ptr), i32(t_by_node_idx),
            counts.data_ptr(), i32(s_edge_ptr), i32(s_edge_idx), i32(e_tgt), i32(s_edge_tgt), B, workspace.data_ptr(), ws_bytes,
            stream_ptr(device),
        ),
        "bl_plan_build",
    )
    seg_type = torch.arange(S, device=device, dtype=torch.int32).remainder_(K) if B > 0 else None
    tables = [None] * 4
    if USE_TMA:
        # P_s, P_t <= E: size the unit tables by that bound so that no host value is needed before the (single) sync below
        tile_rows, slab_rows = tma_tile_rows(), tma_slab_rows()
        tables = [segment_units(s_type_ptr, seg_type, tile_rows, E), segment_units(t_type_ptr, seg_type, tile_rows, E),
                  segment_units(s_type_ptr, seg_type, slab_rows, E), segment_units(t_type_ptr, seg_type, slab_rows, E)]
    if B > 0:
        P_s, P_t = meta[:2].cpu().tolist()  # the one host sync of the plan
        s_tp = t_tp = ()
    else:
        meta_host = meta.cpu().tolist()  # the one host sync of the plan
        P_s, P_t = meta_host[0], meta_host[1]
        s_tp, t_tp = tuple(meta_host[2 : K + 3]), tuple(meta_host[K + 3 :])
    return EdgePlan(
        N, E, K, e_perm[:E], e_src[:E], e_type[:E], row_ptr, urow[:E], vrow[:E],
        s_node[:P_s], s_type_ptr, s_by_node_ptr, s_by_node_idx[:P_s],
        t_node[:P_t], t_type_ptr, t_by_node_ptr, t_by_node_idx[:P_t],
        P_s, P_t, s_tp, t_tp, *tables, s_edge_ptr[: P_s + 1], s_edge_idx[:E], e_tgt[:E], s_edge_tgt[:E], B, S, seg_type,
    )


PLAN_BLOCK_NODES = int(os.environ.get("BUGLAB_B200_PLAN_BLOCK", "8192"))


def plan_block_nodes_for(layer_dims: Sequence[Tuple[int, int]]) -> int:
    """Node-block size of the pair tables for a model whose message-passing layers have the given (D_in, M) shapes:
    ``PLAN_BLOCK_NODES`` when EVERY layer runs forward, backward and weight gradient on the segment-aware TMA GEMMs and the
    split-table edge backward, else 0 (type-major tables, which the first-generation and library paths need)."""
    if not (USE_TMA and USE_SPLIT_EDGE_BACKWARD and PROJECTION_MODE == "f16x3" and PLAN_BLOCK_NODES > 0):
        return 0
    try:
        ok = all(M in (128, 256, 512) and _tma_proj_ok(M, D) and _tma_proj_ok(D, M) and _tma_wgrad_ok(M, D) for D, M in layer_dims)
    except Exception:  # library not built / no driver: host-only use of the model classes
        return 0
    return PLAN_BLOCK_NODES if ok and len(layer_dims) > 0 else 0


# ---------------------------------------------------------------------------------------------------
# Fused typed-edge message + max aggregate (P4 + P5)
# ---------------------------------------------------------------------------------------------------
def _rows_gather(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    out = torch.empty((idx.shape[0], table.shape[1]), device=table.device, dtype=torch.float32)
    check(
        _lib.load().bl_rows_gather(f32(table), i32(idx), idx.shape[0], table.shape[1], f32(out), stream_ptr(table.device)),
        "bl_rows_gather",
    )
    return out


def _project_pairs(rows: torch.Tensor, weight: torch.Tensor, col0: int, type_ptr: Tuple[int, ...],
                   bias: Optional[torch.Tensor]) -> torch.Tensor:
    """out[p] = W_k[:, col0:col0+D] @ rows[p] (+ b_k) for the pairs p of type k — per-type fp32 GEMMs."""
    K, M, _ = weight.shape
    D = rows.shape[1]
    out = torch.empty((rows.shape[0], M), device=rows.device, dtype=torch.float32)
    for k in range(K):
        lo, hi = type_ptr[k], type_ptr[k + 1]
        if hi == lo:
            continue
        w = weight[k, :, col0 : col0 + D]
        if bias is not None:
            torch.addmm(bias[k], rows[lo:hi], w.t(), out=out[lo:hi])
        else:
            torch.mm(rows[lo:hi], w.t(), out=out[lo:hi])
    return out


# Projection arithmetic: "f16x3" = split-fp16 tensor-core GEMMs (default, see csrc/gemm.cu); "fp32" = cuBLAS SGEMM
# through torch.mm (kept as an exact referee for tests and error budgeting).
PROJECTION_MODE = "f16x3"
# With "f16x3": run the forward projections through the hand-written tcgen05 kernel (csrc/pair_project_tc.cu: gather +
# split inside the GEMM loader, no split table in HBM) whenever the shape is supported; else split kernel + cuBLAS.
USE_TCGEN05 = os.environ.get("BUGLAB_B200_TCGEN05", "1") != "0"
# Largest width that goes through the tcgen05 kernels; the 512-wide post-residual layers stay on split + cuBLAS by default
# (measured parity there: 12.6 vs 12.1 ms per table), BUGLAB_B200_TC_MAX_WIDTH=512 switches them over.
TC_MAX_WIDTH = int(os.environ.get("BUGLAB_B200_TC_MAX_WIDTH", "256"))
# Second-generation GEMMs (csrc/gemm_tma.cu: node-level fp16 split, TMA-fed tcgen05, CTA pairs) for every shape they
# support (all widths that are multiples of 256, and 128-wide projections); BUGLAB_B200_TMA=0 restores round 1's kernels.
USE_TMA = os.environ.get("BUGLAB_B200_TMA", "1") != "0"
# Edge backward that writes the gradient tables as fp16 split tables with one writer per row (bl_edge_bwd_targets/_sources)
# instead of fp32 tables + REDs + a split pass; BUGLAB_B200_SPLIT_EDGE_BWD=0 keeps round 1's bl_edge_segmax_bwd.
USE_SPLIT_EDGE_BACKWARD = os.environ.get("BUGLAB_B200_SPLIT_EDGE_BWD", "1") != "0"
# BUGLAB_B200_OVERLAP=1: run the by-source half of the edge backward on a side stream, concurrently with the T-table GEMMs.
# Off by default: measured on power-capped B200s (profiles/r2_overlap_vs_e2e.txt, same box, alternating, 3 x 20 steps each)
# it does not change the resident step (1 064 graphs/s either way — the tensor-bound and the HBM-bound kernels share one power
# budget) and costs ~1.3 % end to end (1 027 vs 1 041 graphs/s) with the loader's copy / plan stream as a third stream.
OVERLAP_EDGE_BACKWARD = os.environ.get("BUGLAB_B200_OVERLAP", "0") != "0"
# Pre-scale BOTH operands of the TMA GEMMs by powers of two (activations and weights to ~2^12) so that their fp16 lo parts
# are normal numbers (unscaled, the lo part of a typical weight is a subnormal: ~17 instead of 22 significant bits)
PRESCALE_OPERANDS = os.environ.get("BUGLAB_B200_PRESCALE", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(device) -> "torch.cuda.Stream":
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def _tma_proj_ok(n_out: int, k_in: int) -> bool:
    return USE_TMA and PROJECTION_MODE == "f16x3" and bool(_lib.load().bl_tma_gemm_supported(n_out, k_in))


def _tma_wgrad_ok(m_out: int, n_in: int) -> bool:
    return USE_TMA and PROJECTION_MODE == "f16x3" and bool(_lib.load().bl_tma_weight_grad_supported(m_out, n_in))


def _host_i32(values: Tuple[int, ...]):
    return (ctypes.c_int32 * len(values))(*values)


def _split3_rows(table: torch.Tensor, idx: Optional[torch.Tensor], amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows = int(idx.shape[0]) if idx is not None else int(table.shape[0])
    dim = int(table.shape[1])
    out = torch.empty((rows, 3 * dim + 8), device=table.device, dtype=torch.float16)
    check(_lib.load().bl_rows_split3_f16(f32(table), i32(idx) if idx is not None else None, rows, dim,
                                         f32(amax) if amax is not None else None, out.data_ptr(),
                                         stream_ptr(table.device)), "bl_rows_split3_f16")
    return out


def _split2_rows(table: torch.Tensor, idx: Optional[torch.Tensor], amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    rows = int(idx.shape[0]) if idx is not None else int(table.shape[0])
    dim = int(table.shape[1])
    out = torch.empty((rows, 2 * dim), device=table.device, dtype=torch.float16)
    check(_lib.load().bl_rows_split2_f16(f32(table), i32(idx) if idx is not None else None, rows, dim,
                                         f32(amax) if amax is not None else None, out.data_ptr(),
                                         stream_ptr(table.device)), "bl_rows_split2_f16")
    return out


def _split3_weights(weight: torch.Tensor, bias: Optional[torch.Tensor], col0: int, in_dim: int, fwd: bool, bwd: bool):
    K, M, ld = weight.shape
    w3 = torch.empty((K, M, 3 * in_dim + 8), device=weight.device, dtype=torch.float16) if fwd else None
    b3 = torch.empty((K, 3 * M, in_dim), device=weight.device, dtype=torch.float16) if bwd else None
    check(_lib.load().bl_weights_split3_f16(f32(weight), f32(bias) if bias is not None else None, K, M, in_dim, ld, col0,
                                             w3.data_ptr() if fwd else None, b3.data_ptr() if bwd else None,
                                             stream_ptr(weight.device)), "bl_weights_split3_f16")
    return w3, b3


def weight_parts(weight: torch.Tensor, n_out: int, k_in: int, col0: int, transposed: bool,
                 amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp16 hi/lo parts [K, 2, n_out, k_in] of weight[:, :, col0:...] (or of its transpose) for the tcgen05 kernels,
    pre-scaled by the power of two derived from ``amax`` (TMA kernels only: they undo it through ``amax_b``)."""
    K, _, ld = weight.shape
    parts = torch.empty((K, 2, n_out, k_in), device=weight.device, dtype=torch.float16)
    check(_lib.load().bl_weight_parts_f16(f32(weight), K, n_out, k_in, ld, col0, 1 if transposed else 0,
                                          f32(amax) if amax is not None else None, parts.data_ptr(),
                                          stream_ptr(weight.device)), "bl_weight_parts_f16")
    return parts


def absmax(x: torch.Tensor) -> torch.Tensor:
    """Device scalar max|x| (source of the power-of-two pre-scales of the split tables); no host sync."""
    out = torch.empty(1, device=x.device, dtype=torch.float32)
    check(_lib.load().bl_absmax(f32(x), x.numel(), f32(out), stream_ptr(x.device)), "bl_absmax")
    return out


def pair_project_tc(src: torch.Tensor, idx: Optional[torch.Tensor], parts: torch.Tensor, bias: Optional[torch.Tensor],
                    type_ptr_dev: torch.Tensor, num_rows: int, amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    K, _, n_out, k_in = parts.shape
    out = torch.empty((num_rows, n_out), device=src.device, dtype=torch.float32)
    check(_lib.load().bl_pair_project_tc(f32(src), i32(idx) if idx is not None else None,
                                         f32(amax) if amax is not None else None, parts.data_ptr(),
                                         f32(bias) if bias is not None else None, i32(type_ptr_dev), K, num_rows, n_out, k_in,
                                         f32(out), stream_ptr(src.device)), "bl_pair_project_tc")
    return out


def pair_weight_grad_tc(g: torch.Tensor, x: torch.Tensor, idx: torch.Tensor, amax: Optional[torch.Tensor],
                        type_ptr_dev: torch.Tensor, d_weight: torch.Tensor, col0: int) -> None:
    """d_weight[k, :, col0:col0+x.shape[1]] = sum over pair rows of type k of g[p]^T x[idx[p]] (tcgen05 kernel)."""
    K, M, ld = d_weight.shape
    check(_lib.load().bl_pair_weight_grad_tc(f32(g), f32(x), i32(idx), f32(amax) if amax is not None else None,
                                             i32(type_ptr_dev), K, int(idx.shape[0]), M, int(x.shape[1]), f32(d_weight), ld, col0,
                                             stream_ptr(g.device)), "bl_pair_weight_grad_tc")


# ---------------------------------------------------------------------------------------------------
# Second-generation GEMMs (csrc/gemm_tma.cu): split once per table, TMA-fed tcgen05, CTA pairs
# ---------------------------------------------------------------------------------------------------
def rows_split(x: torch.Tensor, idx: Optional[torch.Tensor] = None, amax: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp16 hi/lo split table ``[2, rows + 1, dim]`` of ``x`` (or of ``x[idx]``), pre-scaled by the power of two derived
    from ``amax``; the extra last row of each part is zero."""
    rows = int(idx.shape[0]) if idx is not None else int(x.shape[0])
    dim = int(x.shape[1])
    out = torch.empty((2, rows + 1, dim), device=x.device, dtype=torch.float16)
    check(_lib.load().bl_rows_split_f16(f32(x), i32(idx) if idx is not None else None, rows, dim,
                                        f32(amax) if amax is not None else None, out.data_ptr(), stream_ptr(x.device)),
          "bl_rows_split_f16")
    return out


def unit_prefix(seg_ptr: torch.Tensor, unit: int) -> torch.Tensor:
    """``prefix[s] = sum_{s' < s} ceil(rows(s') / unit)`` on the device."""
    num_segs = int(seg_ptr.shape[0]) - 1
    out = torch.empty(num_segs + 1, device=seg_ptr.device, dtype=torch.int32)
    check(_lib.load().bl_segment_unit_prefix(i32(seg_ptr), num_segs, int(unit), i32(out), stream_ptr(seg_ptr.device)),
          "bl_segment_unit_prefix")
    return out


class Units(NamedTuple):
    """Device-side work-unit table of the TMA GEMMs (``bl_segment_units``): ``table[u] = (first row, end row, weight matrix,
    segment)`` for every run of <= ``unit`` consecutive pair rows of one segment; ``count`` = number of units (device scalar);
    ``capacity`` = rows of ``table`` (an upper bound of the count known without a host sync)."""

    table: torch.Tensor
    count: torch.Tensor
    capacity: int
    unit: int


def segment_units(seg_ptr: torch.Tensor, seg_type: Optional[torch.Tensor], unit: int, max_rows: int) -> Units:
    """Work units of ``unit`` rows over the segments ``[seg_ptr[s], seg_ptr[s+1])`` (weight matrix ``seg_type[s]``, or ``s``);
    ``max_rows`` = any upper bound of the total number of pair rows."""
    num_segs = int(seg_ptr.shape[0]) - 1
    capacity = max(1, int(max_rows) // int(unit) + num_segs)
    dev = seg_ptr.device
    table = torch.empty((capacity, 4), device=dev, dtype=torch.int32)
    count = torch.empty(1, device=dev, dtype=torch.int32)
    scratch = torch.empty(num_segs + 1, device=dev, dtype=torch.int32)
    check(_lib.load().bl_segment_units(i32(seg_ptr), i32(seg_type) if seg_type is not None else None, num_segs, int(unit),
                                       capacity, i32(scratch), table.data_ptr(), i32(count), stream_ptr(dev)), "bl_segment_units")
    return Units(table, count, capacity, int(unit))


def tma_tile_rows() -> int:
    return int(_lib.load().bl_tma_tile_rows())


def tma_slab_rows() -> int:
    return int(_lib.load().bl_tma_slab_rows())


def tma_project(a_split: torch.Tensor, idx: Optional[torch.Tensor], parts: torch.Tensor, bias: Optional[torch.Tensor],
                amax: Optional[torch.Tensor], tiles: Units, num_rows: int, slabs: Optional[Units] = None,
                amax_b: Optional[torch.Tensor] = None) -> torch.Tensor:
    """``out[p] = (1/(s_a s_b)) * A[row(p)] @ W_type.T (+ bias_type)`` on the TMA-fed tcgen05 kernels; ``a_split`` from
    :func:`rows_split`, ``parts`` from :func:`weight_parts`, ``tiles`` from :func:`segment_units` (unit ``tma_tile_rows()``).
    256 x 256 products run on the weight-stationary variant when it is enabled and ``slabs`` is given."""
    num_types, _, n_out, k_in = parts.shape
    out = torch.empty((num_rows, n_out), device=a_split.device, dtype=torch.float32)
    lib = _lib.load()
    common = (a_split.data_ptr(), int(a_split.shape[1]), i32(idx) if idx is not None else None, parts.data_ptr(),
              f32(bias) if bias is not None else None, f32(amax) if amax is not None else None,
              f32(amax_b) if amax_b is not None else None)
    if slabs is not None and lib.bl_tma_project_stationary_supported(n_out, k_in):
        check(lib.bl_tma_project_stationary(*common, slabs.table.data_ptr(), i32(slabs.count), num_types, num_rows,
                                            slabs.capacity, n_out, k_in, f32(out),
                                            stream_ptr(a_split.device)), "bl_tma_project_stationary")
        return out
    check(lib.bl_tma_project(*common, tiles.table.data_ptr(), i32(tiles.count), num_types, num_rows, tiles.capacity, n_out, k_in,
                             f32(out), stream_ptr(a_split.device)), "bl_tma_project")
    return out


def tma_weight_grad(g_split: torch.Tensor, x_split: torch.Tensor, idx: torch.Tensor, amax: Optional[torch.Tensor],
                    slabs: Units, d_weight: torch.Tensor, col0: int, amax_x: Optional[torch.Tensor] = None) -> None:
    """``d_weight[type, :, col0:col0+n] = (1/(s_g s_x)) * sum_p G[p]^T X[idx[p]]`` (block zeroed first) on the TMA-fed
    tcgen05 kernel; ``slabs`` from :func:`segment_units` (unit ``tma_slab_rows()``)."""
    num_types, m_out, ld = d_weight.shape
    n_in = int(x_split.shape[2])
    num_rows = int(idx.shape[0])
    check(_lib.load().bl_tma_weight_grad(g_split.data_ptr(), int(g_split.shape[1]), x_split.data_ptr(), int(x_split.shape[1]),
                                         i32(idx), f32(amax) if amax is not None else None,
                                         f32(amax_x) if amax_x is not None else None, slabs.table.data_ptr(), i32(slabs.count),
                                         num_types, num_rows, slabs.capacity, m_out, n_in, f32(d_weight), ld, col0,
                                         stream_ptr(g_split.device)), "bl_tma_weight_grad")


def _project_pairs_f16x3(h: torch.Tensor, idx: torch.Tensor, weight: torch.Tensor, col0: int,
                          type_ptr: Tuple[int, ...], bias: Optional[torch.Tensor],
                          type_ptr_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    K, M, _ = weight.shape
    D = h.shape[1]
    # measured on B200 (scripts/bench_project.py): fused 4.6 ms vs split+cuBLAS 4.8 ms at D=M=256, 14.3 vs 11.8 ms at 512
    if USE_TCGEN05 and type_ptr_dev is not None and M <= TC_MAX_WIDTH and _lib.load().bl_pair_project_tc_supported(M, D):
        return pair_project_tc(h, idx, weight_parts(weight, M, D, col0, False), bias, type_ptr_dev, int(idx.shape[0]))
    a3 = _split3_rows(h, idx)
    w3, _ = _split3_weights(weight, bias, col0, D, True, False)
    out = torch.empty((idx.shape[0], M), device=h.device, dtype=torch.float32)
    check(_lib.load().bl_pair_project_fwd(a3.data_ptr(), w3.data_ptr(), _host_i32(type_ptr), K, M, D, f32(out),
                                          stream_ptr(h.device)), "bl_pair_project_fwd")
    return out


# Test hook: when set to a list, every forward appends the winning ORIGINAL edge index per (node, channel)
# (num_edges for nodes without in-edges) so that an oracle can be evaluated with the same max-routing.
WINNER_TRACE: Optional[list] = None


class TypedEdgeMessageMax(torch.autograd.Function):
    """agg[n] = max over in-edges (s->n, type k) of GELU(W_k [h_s; h_n] + b_k); 0 for isolated nodes.

    ``weight`` is the stack ``[K, M, 2*D]`` of the per-type ``Linear(2D -> M)`` weights, ``bias`` ``[K, M]``
    or None.  Forward keeps only ``h``, ``weight``, the winning pre-activation and winning edge per
    (node, channel); the U/V tables are transient.
    """

    @staticmethod
    def forward(ctx, h: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], plan: EdgePlan):
        lib = _lib.load()
        h = h.contiguous()
        weight = weight.contiguous()
        N, D = h.shape
        K, M, twoD = weight.shape
        if twoD != 2 * D or K != plan.num_edge_types or N != plan.num_nodes:
            raise ValueError(f"shape mismatch: h {tuple(h.shape)}, weight {tuple(weight.shape)}, plan K={plan.num_edge_types} N={plan.num_nodes}")
        bias_c = bias.contiguous() if bias is not None else None
        h_split = amax_h = amax_w = None
        with torch.no_grad():
            if _tma_proj_ok(M, D) and plan.s_tiles is not None:
                # split h ONCE per layer at node granularity; both projections gather its rows by TMA.  Both operands are
                # pre-scaled by exact powers of two (their absolute maxima brought to ~2^12) so that the lo parts are normal
                # fp16 numbers; the kernels' epilogues undo both scales.
                amax_h, amax_w = (absmax(h), absmax(weight)) if PRESCALE_OPERANDS else (None, None)
                h_split = rows_split(h, None, amax_h)
                u_rows = tma_project(h_split, plan.s_node, weight_parts(weight, M, D, 0, False, amax_w), None, amax_h,
                                     plan.s_tiles, plan.num_s_pairs, plan.s_slabs, amax_w)
                v_rows = tma_project(h_split, plan.t_node, weight_parts(weight, M, D, D, False, amax_w), bias_c, amax_h,
                                     plan.t_tiles, plan.num_t_pairs, plan.t_slabs, amax_w)
            elif plan.block_nodes > 0:
                raise _lib.BuglabB200Error(f"a node-blocked plan needs the TMA GEMMs, which do not cover (D={D}, M={M})")
            elif PROJECTION_MODE == "f16x3":
                u_rows = _project_pairs_f16x3(h, plan.s_node, weight, 0, plan.s_type_ptr_host, None, plan.s_type_ptr)
                v_rows = _project_pairs_f16x3(h, plan.t_node, weight, D, plan.t_type_ptr_host, bias_c, plan.t_type_ptr)
            else:
                hs = _rows_gather(h, plan.s_node)
                u_rows = _project_pairs(hs, weight, 0, plan.s_type_ptr_host, None)
                del hs
                ht = _rows_gather(h, plan.t_node)
                v_rows = _project_pairs(ht, weight, D, plan.t_type_ptr_host, bias_c)
                del ht
            agg = torch.empty((N, M), device=h.device, dtype=torch.float32)
            xwin = torch.empty_like(agg)
            ewin = torch.empty((N, M), device=h.device, dtype=torch.int32)
            check(
                lib.bl_edge_segmax_fwd(f32(u_rows), f32(v_rows), i32(plan.row_ptr), i32(plan.urow), i32(plan.vrow),
                                       N, M, f32(agg), f32(xwin), i32(ewin), stream_ptr(h.device)),
                "bl_edge_segmax_fwd",
            )
        if WINNER_TRACE is not None:
            e = ewin.long()
            WINNER_TRACE.append(torch.where(e >= 0, plan.e_perm.long()[e.clamp(min=0)], torch.full_like(e, plan.num_edges)).cpu())
        ctx.plan = plan
        ctx.has_bias = bias is not None
        ctx.mode = PROJECTION_MODE
        ctx.h_split = h_split if (h_split is not None and _tma_wgrad_ok(M, D)) else None  # x operand of the weight gradient
        ctx.amax_h, ctx.amax_w = (amax_h, amax_w) if ctx.h_split is not None else (None, None)
        ctx.save_for_backward(h, weight, xwin, ewin)
        return agg

    @staticmethod
    def backward(ctx, d_agg: torch.Tensor):
        lib = _lib.load()
        plan: EdgePlan = ctx.plan
        h, weight, xwin, ewin = ctx.saved_tensors
        N, D = h.shape
        K, M, _ = weight.shape
        d_agg = d_agg.contiguous()
        dev = h.device
        d_bias = torch.zeros((K, M), device=dev, dtype=torch.float32) if ctx.has_bias else None
        if (ctx.mode == "f16x3" and USE_SPLIT_EDGE_BACKWARD and M in (128, 256, 512) and _tma_proj_ok(D, M) and _tma_wgrad_ok(M, D)
                and plan.s_tiles is not None and plan.s_edge_ptr is not None):
            # Second-generation backward, end to end: the edge kernels write both gradient tables directly as pre-scaled
            # fp16 hi/lo split tables (one writer per row, no memset, no atomics on the tables) and fold the bias column
            # sums in; the TMA-fed tcgen05 kernels take it from there.
            amax_in = torch.empty(1, device=dev, dtype=torch.float32)
            check(lib.bl_absmax(f32(d_agg), d_agg.numel(), f32(amax_in), stream_ptr(dev)), "bl_absmax")
            amax = torch.empty(1, device=dev, dtype=torch.float32)
            g_rows = torch.empty((N, M), device=dev, dtype=torch.float32)
            dv_split = torch.empty((2, plan.num_t_pairs + 1, M), device=dev, dtype=torch.float16)
            check(lib.bl_edge_bwd_targets(f32(d_agg), f32(xwin), i32(ewin), i32(plan.row_ptr), i32(plan.vrow), i32(plan.e_type),
                                          N, M, K, plan.num_t_pairs, f32(amax_in), f32(amax), f32(g_rows), dv_split.data_ptr(),
                                          f32(d_bias) if d_bias is not None else None, stream_ptr(dev)), "bl_edge_bwd_targets")
            du_split = torch.empty((2, plan.num_s_pairs + 1, M), device=dev, dtype=torch.float16)
            # The by-source kernel is latency/HBM-bound and independent of the T-table GEMMs: run it on a side stream so it
            # overlaps those tensor-bound kernels (its blocks co-reside with the persistent GEMM CTAs), then join.
            main = torch.cuda.current_stream(dev)
            side = _side_stream(dev) if OVERLAP_EDGE_BACKWARD else None
            if side is not None:
                side.wait_stream(main)
                with torch.cuda.stream(side):
                    check(lib.bl_edge_bwd_sources(f32(g_rows), i32(ewin), i32(plan.s_edge_ptr), i32(plan.s_edge_idx), i32(plan.s_edge_tgt),
                                                  plan.num_s_pairs, M, f32(amax), du_split.data_ptr(), stream_ptr(dev)),
                          "bl_edge_bwd_sources")
            else:
                check(lib.bl_edge_bwd_sources(f32(g_rows), i32(ewin), i32(plan.s_edge_ptr), i32(plan.s_edge_idx), i32(plan.s_edge_tgt),
                                              plan.num_s_pairs, M, f32(amax), du_split.data_ptr(), stream_ptr(dev)), "bl_edge_bwd_sources")
            if ctx.h_split is not None:
                h_split, amax_h, amax_w = ctx.h_split, ctx.amax_h, ctx.amax_w
            else:
                amax_h, amax_w = (absmax(h), absmax(weight)) if PRESCALE_OPERANDS else (None, None)
                h_split = rows_split(h, None, amax_h)
            d_weight = torch.empty_like(weight)
            d_rows = [None, None]
            for slot, rows_idx, g_split, col0, tiles, slabs in (
                    (1, plan.t_node, dv_split, D, plan.t_tiles, plan.t_slabs),
                    (0, plan.s_node, du_split, 0, plan.s_tiles, plan.s_slabs)):
                if slot == 0 and side is not None:
                    main.wait_stream(side)  # dU is complete; g_rows / ewin are no longer read on the side stream
                d_rows[slot] = tma_project(g_split, None, weight_parts(weight, D, M, col0, True, amax_w), None, amax, tiles,
                                           int(rows_idx.shape[0]), slabs, amax_w)
                tma_weight_grad(g_split, h_split, rows_idx, amax, slabs, d_weight, col0, amax_h)
            del du_split, dv_split, g_rows
            d_h = torch.empty_like(h)
            check(lib.bl_rows_segment_sum(f32(d_rows[0]), i32(plan.s_by_node_ptr), i32(plan.s_by_node_idx),
                                          f32(d_rows[1]), i32(plan.t_by_node_ptr), i32(plan.t_by_node_idx),
                                          N, D, 0, None, f32(d_h), stream_ptr(dev)), "bl_rows_segment_sum")
            return d_h, d_weight, d_bias, None

        if plan.block_nodes > 0:
            raise _lib.BuglabB200Error(f"a node-blocked plan needs the TMA backward, which does not cover (D={D}, M={M})")
        du = torch.empty((plan.num_s_pairs, M), device=dev, dtype=torch.float32)
        dv = torch.empty((plan.num_t_pairs, M), device=dev, dtype=torch.float32)
        amax = torch.empty(1, device=dev, dtype=torch.float32) if ctx.mode == "f16x3" else None
        check(
            lib.bl_edge_segmax_bwd(f32(d_agg), f32(xwin), i32(ewin), i32(plan.row_ptr), i32(plan.urow), i32(plan.vrow),
                                   N, M, plan.num_s_pairs, plan.num_t_pairs, f32(du), f32(dv),
                                   f32(amax) if amax is not None else None, stream_ptr(dev)),
            "bl_edge_segmax_bwd",
        )
        d_rows = []
        unscaled = False
        if ctx.mode == "f16x3" and _tma_proj_ok(D, M) and plan.s_tiles is not None:
            # second-generation path: split each gradient table once (pre-scaled), then TMA-fed tcgen05 for both products
            unscaled = True  # the projection epilogue undoes the pre-scale
            d_weight = torch.empty_like(weight)
            if d_bias is not None:
                check(lib.bl_grouped_colsum(f32(dv), i32(plan.t_type_ptr), K, M, f32(d_bias), stream_ptr(dev)), "bl_grouped_colsum")
            wg_ok = _tma_wgrad_ok(M, D)
            amax_h, amax_w = ctx.amax_h, ctx.amax_w
            h_split = ctx.h_split
            if h_split is None and wg_ok:
                amax_h = absmax(h) if PRESCALE_OPERANDS else None
                h_split = rows_split(h, None, amax_h)
            if amax_w is None and PRESCALE_OPERANDS:
                amax_w = absmax(weight)
            for rows_idx, d_tab, col0, type_ptr, tiles, slabs in (
                    (plan.s_node, du, 0, plan.s_type_ptr_host, plan.s_tiles, plan.s_slabs),
                    (plan.t_node, dv, D, plan.t_type_ptr_host, plan.t_tiles, plan.t_slabs)):
                g_split = rows_split(d_tab, None, amax)
                d_rows.append(tma_project(g_split, None, weight_parts(weight, D, M, col0, True, amax_w), None, amax, tiles,
                                          int(rows_idx.shape[0]), slabs, amax_w))
                if wg_ok:
                    tma_weight_grad(g_split, h_split, rows_idx, amax, slabs, d_weight, col0, amax_h)
                else:  # widths the weight-gradient kernel does not cover (e.g. 128): round 1's split + library GEMM
                    g2 = _split2_rows(d_tab, None, amax)
                    a2 = _split2_rows(h, rows_idx)
                    tmp = torch.empty((K, 2 * M, 2 * D), device=dev, dtype=torch.float32)
                    check(lib.bl_pair_project_bwd_weight(g2.data_ptr(), 2 * M, 0, a2.data_ptr(), _host_i32(type_ptr), K, M, D,
                                                         f32(amax), f32(tmp), f32(d_weight), 2 * D, col0, stream_ptr(dev)),
                          "bl_pair_project_bwd_weight")
                    del g2, a2, tmp
                del g_split
            del du, dv
        elif ctx.mode == "f16x3":
            d_weight = torch.empty_like(weight)
            if d_bias is not None:
                check(lib.bl_grouped_colsum(f32(dv), i32(plan.t_type_ptr), K, M, f32(d_bias), stream_ptr(dev)), "bl_grouped_colsum")
            use_tc = USE_TCGEN05 and D <= TC_MAX_WIDTH and bool(lib.bl_pair_project_tc_supported(D, M))
            for rows_idx, d_tab, col0, type_ptr, type_ptr_dev in (
                    (plan.s_node, du, 0, plan.s_type_ptr_host, plan.s_type_ptr),
                    (plan.t_node, dv, D, plan.t_type_ptr_host, plan.t_type_ptr)):
                tp = _host_i32(type_ptr)
                # both results below carry the power-of-two pre-scale of the gradient table; it is undone in
                # bl_rows_segment_sum (d_in) and in the fold kernel of bl_pair_project_bwd_weight (d_weight)
                # measured (scripts/bench_project.py, B200): 5.8 ms vs 6.2 ms for split2 + cuBLAS at D=M=256; the 512-wide layers
                # would re-load operands per 128x256 output tile and stay on the library path
                use_tc_wg = USE_TCGEN05 and M <= TC_MAX_WIDTH and D <= TC_MAX_WIDTH and bool(lib.bl_pair_weight_grad_tc_supported(M, D))
                g = None
                if use_tc:
                    # d(rows) = dTable @ W_k[:, col0:col0+D] on the hand-written tcgen05 kernel (reads the fp32 table directly)
                    d_in = pair_project_tc(d_tab, None, weight_parts(weight, D, M, col0, True), None, type_ptr_dev,
                                           int(rows_idx.shape[0]), amax=amax)
                    if not use_tc_wg:
                        g, g_stride, g_col0 = _split2_rows(d_tab, None, amax), 2 * M, 0
                else:
                    g, g_stride, g_col0 = _split3_rows(d_tab, None, amax), 3 * M + 8, M
                    _, b3 = _split3_weights(weight, None, col0, D, False, True)
                    d_in = torch.empty((rows_idx.shape[0], D), device=dev, dtype=torch.float32)
                    check(lib.bl_pair_project_bwd_input(g.data_ptr(), b3.data_ptr(), tp, K, M, D, f32(d_in), stream_ptr(dev)),
                          "bl_pair_project_bwd_input")
                    del b3
                if use_tc_wg:
                    # dW_k = dTable^T h[rows] with the transposing tcgen05 loader: no split tables at all
                    pair_weight_grad_tc(d_tab, h, rows_idx, amax, type_ptr_dev, d_weight, col0)
                    a2 = tmp = None
                else:
                    a2 = _split2_rows(h, rows_idx)  # recomputed instead of kept alive since forward
                    tmp = torch.empty((K, 2 * M, 2 * D), device=dev, dtype=torch.float32)
                    check(lib.bl_pair_project_bwd_weight(g.data_ptr(), g_stride, g_col0, a2.data_ptr(), tp, K, M, D, f32(amax),
                                                         f32(tmp), f32(d_weight), 2 * D, col0, stream_ptr(dev)),
                          "bl_pair_project_bwd_weight")
                d_rows.append(d_in)
                del g, a2, tmp
            del du, dv
        else:
            d_weight = torch.zeros_like(weight)
            for rows_idx, d_tab, col0, type_ptr, is_t in (
                (plan.s_node, du, 0, plan.s_type_ptr_host, False),
                (plan.t_node, dv, D, plan.t_type_ptr_host, True),
            ):
                rows = _rows_gather(h, rows_idx)  # recomputed instead of kept alive since forward
                d_in = torch.empty_like(rows)
                for k in range(K):
                    lo, hi = type_ptr[k], type_ptr[k + 1]
                    if hi == lo:
                        continue
                    w = weight[k, :, col0 : col0 + D]
                    torch.mm(d_tab[lo:hi], w, out=d_in[lo:hi])
                    d_weight[k, :, col0 : col0 + D].copy_(torch.mm(d_tab[lo:hi].t(), rows[lo:hi]))
                    if is_t and d_bias is not None:
                        torch.sum(d_tab[lo:hi], dim=0, out=d_bias[k])
                d_rows.append(d_in)
                del rows
        d_h = torch.empty_like(h)
        check(
            lib.bl_rows_segment_sum(f32(d_rows[0]), i32(plan.s_by_node_ptr), i32(plan.s_by_node_idx),
                                    f32(d_rows[1]), i32(plan.t_by_node_ptr), i32(plan.t_by_node_idx),
                                    N, D, 0, f32(amax) if (amax is not None and not unscaled) else None, f32(d_h), stream_ptr(dev)),
            "bl_rows_segment_sum",
        )
        return d_h, d_weight, d_bias, None


def typed_edge_message_max(h, weight, bias, plan: EdgePlan) -> torch.Tensor:
    return TypedEdgeMessageMax.apply(h, weight, bias, plan)


# ---------------------------------------------------------------------------------------------------
# Dense Linear (no bias) on the tensor cores with the same split-fp16 scheme — the node-update Linear(M -> D_out)
# ---------------------------------------------------------------------------------------------------
class DenseLinearF16x3(torch.autograd.Function):
    """y = x @ weight.T for x [R, K_in], weight [N_out, K_in], fp32 in/out.

    Forward: plain fp32 library GEMM by default (``exact_forward``) — forward rounding noise is amplified ~2x per layer
    and the whole 8-layer stack must stay within 1e-4 of the reference, while this GEMM is only ~2 % of the step.
    Backward (2/3 of the work, not on the forward error path): split-fp16 tensor-core GEMMs."""

    exact_forward = True

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor):
        lib = _lib.load()
        x = x.contiguous()
        weight = weight.contiguous()
        R, K_in = x.shape
        N_out = weight.shape[0]
        if DenseLinearF16x3.exact_forward:
            y = torch.mm(x, weight.t())
        else:
            w = weight.view(1, N_out, K_in)
            a3 = _split3_rows(x, None)
            w3, _ = _split3_weights(w, None, 0, K_in, True, False)
            y = torch.empty((R, N_out), device=x.device, dtype=torch.float32)
            check(lib.bl_pair_project_fwd(a3.data_ptr(), w3.data_ptr(), _host_i32((0, R)), 1, N_out, K_in, f32(y),
                                          stream_ptr(x.device)), "bl_pair_project_fwd")
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        lib = _lib.load()
        x, weight = ctx.saved_tensors
        dy = dy.contiguous()
        R, K_in = x.shape
        N_out = weight.shape[0]
        dev = x.device
        w = weight.view(1, N_out, K_in)
        tp = _host_i32((0, R))
        amax = torch.empty(1, device=dev, dtype=torch.float32)
        check(lib.bl_absmax(f32(dy), dy.numel(), f32(amax), stream_ptr(dev)), "bl_absmax")
        g3 = _split3_rows(dy, None, amax)
        _, b3 = _split3_weights(w, None, 0, K_in, False, True)
        dx = torch.empty_like(x)
        check(lib.bl_pair_project_bwd_input(g3.data_ptr(), b3.data_ptr(), tp, 1, N_out, K_in, f32(dx), stream_ptr(dev)),
              "bl_pair_project_bwd_input")
        check(lib.bl_unscale_pow2(f32(dx), dx.numel(), f32(amax), stream_ptr(dev)), "bl_unscale_pow2")
        a2 = _split2_rows(x, None)
        dw = torch.empty_like(weight)
        tmp = torch.empty((1, 2 * N_out, 2 * K_in), device=dev, dtype=torch.float32)
        check(lib.bl_pair_project_bwd_weight(g3.data_ptr(), 3 * N_out + 8, N_out, a2.data_ptr(), tp, 1, N_out, K_in, f32(amax),
                                             f32(tmp), f32(dw), K_in, 0, stream_ptr(dev)), "bl_pair_project_bwd_weight")
        return dx, dw


def _single_segment(rows: int, device) -> torch.Tensor:
    """Device tensor ``[0, rows]`` (one segment covering the whole table) built without a host->device copy."""
    seg = torch.full((2,), int(rows), dtype=torch.int32, device=device)
    seg[0].zero_()
    return seg


# Diagnostics only (scripts/diag_precision.py): evaluate the node-update Linear's FORWARD with a plain fp32 library GEMM
DENSE_FORWARD_FP32_REFEREE = os.environ.get("BUGLAB_B200_DENSE_FP32_REFEREE", "0") == "1"


class DenseLinearTma(torch.autograd.Function):
    """y = x @ weight.T (no bias) for x [R, K_in], weight [N_out, K_in] on the TMA-fed tcgen05 kernels: the node-update
    ``Linear(M -> D_out)`` of the message-passing layer.  Same split-fp16 arithmetic as the projections (fp32-class
    accuracy), forward and both backward products; nothing runs on a library GEMM."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, weight: torch.Tensor):
        x = x.contiguous()
        weight = weight.contiguous()
        R, K_in = x.shape
        N_out = weight.shape[0]
        seg = _single_segment(R, x.device)
        amax_x, amax_w = (absmax(x), absmax(weight)) if PRESCALE_OPERANDS else (None, None)
        x_split = rows_split(x, None, amax_x)
        tiles, slabs = segment_units(seg, None, tma_tile_rows(), R), segment_units(seg, None, tma_slab_rows(), R)
        if DENSE_FORWARD_FP32_REFEREE:
            y = torch.mm(x, weight.t())
        else:
            y = tma_project(x_split, None, weight_parts(weight.view(1, N_out, K_in), N_out, K_in, 0, False, amax_w), None, amax_x,
                            tiles, R, slabs, amax_w)
        ctx.units = (tiles, slabs)
        ctx.amax = (amax_x, amax_w)
        ctx.save_for_backward(x_split, weight)
        return y

    @staticmethod
    def backward(ctx, dy: torch.Tensor):
        lib = _lib.load()
        x_split, weight = ctx.saved_tensors
        (tiles, slabs), (amax_x, amax_w) = ctx.units, ctx.amax
        dy = dy.contiguous()
        R, N_out = dy.shape
        K_in = weight.shape[1]
        dev = dy.device
        amax = torch.empty(1, device=dev, dtype=torch.float32)
        check(lib.bl_absmax(f32(dy), dy.numel(), f32(amax), stream_ptr(dev)), "bl_absmax")
        g_split = rows_split(dy, None, amax)
        dx = tma_project(g_split, None, weight_parts(weight.view(1, N_out, K_in), K_in, N_out, 0, True, amax_w), None, amax, tiles,
                         R, slabs, amax_w)
        dw = torch.empty_like(weight)
        identity = torch.arange(R, device=dev, dtype=torch.int32)
        tma_weight_grad(g_split, x_split, identity, amax, slabs, dw.view(1, N_out, K_in), 0, amax_x)
        return dx, dw


def dense_linear(x: torch.Tensor, weight: torch.Tensor) -> torch.Tensor:
    """Bias-free Linear: split-fp16 tensor-core GEMMs by default, plain fp32 library GEMM in "fp32" mode or for odd widths."""
    if not x.is_cuda:
        raise _lib.BuglabB200Error("dense_linear: buglab_b200 has no CPU path (x is a CPU tensor)")
    n_out, k_in = int(weight.shape[0]), int(x.shape[1])
    if _tma_proj_ok(n_out, k_in) and _tma_proj_ok(k_in, n_out) and _tma_wgrad_ok(n_out, k_in):
        return DenseLinearTma.apply(x, weight)
    if PROJECTION_MODE == "f16x3" and x.shape[1] % 4 == 0 and weight.shape[0] % 4 == 0:
        return DenseLinearF16x3.apply(x, weight)
    return torch.nn.functional.linear(x, weight)


# ---------------------------------------------------------------------------------------------------
# Node update pieces
# ---------------------------------------------------------------------------------------------------
LN_PARTIALS = 1184  # BL_LN_PARTIALS in include/buglab_b200.h (8 blocks per SM)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps: float):
        x = x.contiguous()
        rows, dim = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, device=x.device, dtype=torch.float32)
        rstd = torch.empty_like(mean)
        check(_lib.load().bl_layernorm_fwd(f32(x), f32(gamma.contiguous()), f32(beta.contiguous()), rows, dim, eps,
                                            f32(y), f32(mean), f32(rstd), stream_ptr(x.device)), "bl_layernorm_fwd")
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        rows, dim = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        d_gamma = torch.empty_like(gamma)
        d_beta = torch.empty_like(gamma)
        partial = torch.empty(2 * LN_PARTIALS * dim, device=x.device, dtype=torch.float32)
        check(_lib.load().bl_layernorm_bwd(f32(dy), f32(x), f32(gamma.contiguous()), f32(mean), f32(rstd), rows, dim,
                                            f32(dx), f32(d_gamma), f32(d_beta), f32(partial), stream_ptr(x.device)),
              "bl_layernorm_bwd")
        return dx, d_gamma, d_beta, None


def layer_norm(x, gamma, beta, eps: float = 1e-5):
    return LayerNormFn.apply(x, gamma, beta, eps)


class TanhDropoutFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, p_drop: float, seed: int):
        x = x.contiguous()
        y = torch.empty_like(x)
        t = torch.empty_like(x)
        check(_lib.load().bl_tanh_dropout_fwd(f32(x), x.numel(), p_drop, seed, f32(y), f32(t), stream_ptr(x.device)),
              "bl_tanh_dropout_fwd")
        ctx.save_for_backward(t)
        ctx.p_drop, ctx.seed = p_drop, seed
        return y

    @staticmethod
    def backward(ctx, dy):
        (t,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(t)
        check(_lib.load().bl_tanh_dropout_bwd(f32(dy), f32(t), t.numel(), ctx.p_drop, ctx.seed, f32(dx),
                                               stream_ptr(t.device)), "bl_tanh_dropout_bwd")
        return dx, None, None


def fresh_seed() -> int:
    """A new 63-bit dropout seed drawn from torch's CPU generator state (so torch.manual_seed controls it)."""
    return int(torch.randint(0, 2**62, (1,)).item())


def tanh_dropout(x, p_drop: float, training: bool):
    p = float(p_drop) if training else 0.0
    return TanhDropoutFn.apply(x, p, fresh_seed() if p > 0 else 0)


# ---------------------------------------------------------------------------------------------------
# Segment primitives (torch_scatter surface)
# ---------------------------------------------------------------------------------------------------
def _as_LF(src: torch.Tensor, dim: int):
    if src.dim() == 1:
        return src.contiguous().view(-1, 1), True
    if src.dim() == 2 and dim in (0, -2):
        return src.contiguous(), False
    raise NotImplementedError("buglab_b200 segment ops support 1-D src or 2-D src reduced along dim 0")


class SegmentMinMaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src2d, index, num_segments: int, is_min: bool):
        L, F = src2d.shape
        out = torch.empty((num_segments, F), device=src2d.device, dtype=torch.float32)
        arg = torch.empty((num_segments, F), device=src2d.device, dtype=torch.int32)
        check(_lib.load().bl_segment_minmax(f32(src2d), i32(index), L, F, num_segments, 1 if is_min else 0,
                                             f32(out), i32(arg), stream_ptr(src2d.device)), "bl_segment_minmax")
        ctx.save_for_backward(arg, index)
        ctx.shape = (L, F)
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, d_out, _d_arg):
        arg, index = ctx.saved_tensors
        L, F = ctx.shape
        d_src = torch.empty((L, F), device=d_out.device, dtype=torch.float32)
        check(_lib.load().bl_segment_minmax_bwd(f32(d_out.contiguous()), i32(arg), i32(index), L, F, f32(d_src),
                                                 stream_ptr(d_out.device)), "bl_segment_minmax_bwd")
        return d_src, None, None, None


def _index32(index: torch.Tensor) -> torch.Tensor:
    return index.contiguous() if index.dtype == torch.int32 else index.to(torch.int32)


def _num_segments(index: torch.Tensor, dim_size: Optional[int]) -> int:
    if dim_size is not None:
        return int(dim_size)
    return int(index.max().item()) + 1 if index.numel() > 0 else 0


# Test hook (like WINNER_TRACE): when set to a list, every DIFFERENTIABLE segment max / min appends its arg tensor
# ([S, F] positions in src; src.shape[0] for empty segments) so that an oracle can be evaluated with the same routing.
MINMAX_TRACE: Optional[list] = None


def segment_minmax(src, index, dim=-1, dim_size=None, is_min=False):
    src2d, was_1d = _as_LF(src.float(), dim)
    S = _num_segments(index, dim_size)
    out, arg = SegmentMinMaxFn.apply(src2d, _index32(index), S, is_min)
    if MINMAX_TRACE is not None and torch.is_grad_enabled() and src2d.requires_grad:
        MINMAX_TRACE.append(arg.detach().to(torch.int64).cpu())
    arg = arg.to(torch.int64)
    if was_1d:
        return out.view(-1), arg.view(-1)
    return out, arg


class SegmentSumFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src2d, index, num_segments: int):
        L, F = src2d.shape
        out = torch.empty((num_segments, F), device=src2d.device, dtype=torch.float32)
        check(_lib.load().bl_segment_sum(f32(src2d), i32(index), L, F, num_segments, f32(out),
                                          stream_ptr(src2d.device)), "bl_segment_sum")
        ctx.save_for_backward(index)
        return out

    @staticmethod
    def backward(ctx, d_out):
        (index,) = ctx.saved_tensors
        return _rows_gather_any(d_out.contiguous(), index), None, None


def _rows_gather_any(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    if table.shape[1] % 4 == 0:
        return _rows_gather(table, idx)
    return table[idx.long()]  # tiny head tensors with F not a multiple of 4 (e.g. F == 1)


def segment_sum(src, index, dim=-1, dim_size=None):
    src2d, was_1d = _as_LF(src.float(), dim)
    S = _num_segments(index, dim_size)
    out = SegmentSumFn.apply(src2d, _index32(index), S)
    return out.view(-1) if was_1d else out


class SegmentLogSoftmaxFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src, index, num_segments: int, eps: float):
        src = src.contiguous()
        L = src.shape[0]
        out = torch.empty_like(src)
        seg_max = torch.empty(num_segments, device=src.device, dtype=torch.float32)
        seg_sum = torch.empty_like(seg_max)
        check(_lib.load().bl_segment_log_softmax_fwd(f32(src), i32(index), L, num_segments, eps, f32(out), f32(seg_max),
                                                      f32(seg_sum), stream_ptr(src.device)), "bl_segment_log_softmax_fwd")
        ctx.save_for_backward(out, index)
        ctx.num_segments = num_segments
        return out

    @staticmethod
    def backward(ctx, d_out):
        out, index = ctx.saved_tensors
        L = out.shape[0]
        d_src = torch.empty_like(out)
        tmp = torch.empty(ctx.num_segments, device=out.device, dtype=torch.float32)
        check(_lib.load().bl_segment_log_softmax_bwd(f32(d_out.contiguous()), f32(out), i32(index), L, ctx.num_segments,
                                                      f32(d_src), f32(tmp), stream_ptr(out.device)), "bl_segment_log_softmax_bwd")
        return d_src, None, None, None


def segment_log_softmax(src: torch.Tensor, index: torch.Tensor, eps: float = 1e-12, num_segments: Optional[int] = None):
    """``scatter_log_softmax`` of buglab/models/utils.py:15-28 for 1-D scores."""
    if src.dim() != 1:
        raise NotImplementedError("segment_log_softmax supports 1-D scores")
    if src.numel() == 0:
        return src.float()
    S = _num_segments(index, num_segments)
    return SegmentLogSoftmaxFn.apply(src.float(), _index32(index), S, eps)


# ---------------------------------------------------------------------------------------------------
# Subtoken embedding max-pool (P2)
# ---------------------------------------------------------------------------------------------------
class SubtokenMaxPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, emb, ids, lens, p_drop: float, seed: int):
        emb = emb.contiguous()
        N, T = ids.shape
        H = emb.shape[1]
        out = torch.empty((N, H), device=emb.device, dtype=torch.float32)
        arg = torch.empty((N, H), device=emb.device, dtype=torch.int32)
        check(_lib.load().bl_subtoken_maxpool_fwd(f32(emb), i32(ids), i32(lens), N, T, H, p_drop, seed, f32(out), i32(arg),
                                                   stream_ptr(emb.device)), "bl_subtoken_maxpool_fwd")
        ctx.save_for_backward(ids, arg)
        ctx.meta = (tuple(emb.shape), p_drop, seed)
        return out

    @staticmethod
    def backward(ctx, d_out):
        ids, arg = ctx.saved_tensors
        shape, p_drop, seed = ctx.meta
        N, T = ids.shape
        d_emb = torch.zeros(shape, device=d_out.device, dtype=torch.float32)
        check(_lib.load().bl_subtoken_maxpool_bwd(f32(d_out.contiguous()), i32(ids), i32(arg), N, T, shape[1], p_drop, seed,
                                                   f32(d_emb), stream_ptr(d_out.device)), "bl_subtoken_maxpool_bwd")
        return d_emb, None, None, None, None


def subtoken_maxpool(emb, ids, lens, p_drop: float = 0.0, training: bool = False):
    p = float(p_drop) if training else 0.0
    return SubtokenMaxPoolFn.apply(emb, _index32(ids), _index32(lens), p, fresh_seed() if p > 0 else 0)


# ---------------------------------------------------------------------------------------------------
# Edge-biased attention of the sequence models
# ---------------------------------------------------------------------------------------------------
class SeqAttentionPlan(NamedTuple):
    """The typed edges of one minibatch as attention "entries" (see include/buglab_b200.h): both directions of every edge,
    grouped by query row (``row_*``) and by key (``col_*``).  Built once per minibatch, shared by all layers."""

    num_samples: int
    max_len: int
    num_tables: int          # 2 * relation kinds
    lengths: torch.Tensor    # [B] int32
    row_ptr: torch.Tensor    # [B*L+1] int32
    row_key: torch.Tensor    # [entries] int32
    row_tab: torch.Tensor
    col_ptr: torch.Tensor
    col_query: torch.Tensor
    col_tab: torch.Tensor


def build_seq_attention_plan(edges: torch.Tensor, edge_types: torch.Tensor, lengths: torch.Tensor, max_len: int,
                             num_edge_types: int) -> SeqAttentionPlan:
    """``edges`` [E, 3] = (sample, source position, target position), ``edge_types`` [E] (seqmodel.py:780-782).
    Plain torch index arithmetic (runs where the tensors live); the kernels only read the result."""
    B, L, T = int(lengths.shape[0]), int(max_len), int(num_edge_types)
    sample, src, tgt = edges[:, 0].long(), edges[:, 1].long(), edges[:, 2].long()
    kinds = edge_types.long()
    query = torch.cat((sample * L + src, sample * L + tgt))      # global query row of each entry
    key = torch.cat((tgt, src))
    table = torch.cat((kinds, kinds + T))

    def grouped(major: torch.Tensor, minor: torch.Tensor):
        order = torch.argsort(major * L + minor, stable=True)
        counts = torch.bincount(major, minlength=B * L)
        ptr = torch.zeros(B * L + 1, dtype=torch.int64, device=edges.device)
        torch.cumsum(counts, dim=0, out=ptr[1:])
        return order, ptr.to(torch.int32)

    row_order, row_ptr = grouped(query, key)
    key_row = (query // L) * L + key                             # global key row of each entry
    col_order, col_ptr = grouped(key_row, query % L)
    i32 = lambda t: t.to(torch.int32).contiguous()  # noqa: E731
    return SeqAttentionPlan(B, L, 2 * T, i32(lengths), row_ptr, i32(key[row_order]), i32(table[row_order]),
                            col_ptr, i32((query % L)[col_order]), i32(table[col_order]))


def _seq_attention_backend():
    """(forward, backward) callables with the C-ABI argument order minus the stream.  The product has exactly one backend,
    the CUDA library; tests substitute the host emulation of the same kernel source (tests/emul)."""
    lib = _lib.load()

    def fwd(*args):
        check(lib.bl_seq_attention_fwd(*args, stream_ptr(None)), "bl_seq_attention_fwd")

    def bwd(*args):
        check(lib.bl_seq_attention_bwd(*args, stream_ptr(None)), "bl_seq_attention_bwd")

    return fwd, bwd, f32, i32


class SeqEdgeAttentionFn(torch.autograd.Function):
    """out[b, h, i] = sum_j softmax_j(<q_i, k_j> + edge terms) (v_j + edge value terms); q, k, v: [B, H, L, D] fp32."""

    @staticmethod
    def forward(ctx, q, k, v, bias, vbias, plan: SeqAttentionPlan, p_drop: float = 0.0, seed: int = 0):
        fwd, _, fp, ip = _seq_attention_backend()
        q, k, v, bias = q.contiguous(), k.contiguous(), v.contiguous(), bias.contiguous()
        vbias = vbias.contiguous() if vbias is not None else None
        B, H, L, D = q.shape
        if (B, L) != (plan.num_samples, plan.max_len) or bias.shape != (plan.num_tables, H, D):
            raise ValueError(f"shape mismatch: q {tuple(q.shape)}, bias {tuple(bias.shape)}, plan B={plan.num_samples} "
                             f"L={plan.max_len} tables={plan.num_tables}")
        out = torch.empty_like(q)
        lse = torch.empty((B, H, L), device=q.device, dtype=torch.float32)
        fwd(fp(q), fp(k), fp(v), ip(plan.lengths), fp(bias), fp(vbias) if vbias is not None else None, ip(plan.row_ptr),
            ip(plan.row_key), ip(plan.row_tab), B, H, L, D, plan.num_tables, float(p_drop), int(seed), fp(out), fp(lse))
        ctx.plan = plan
        ctx.dropout = (float(p_drop), int(seed))
        ctx.has_vbias = vbias is not None
        ctx.save_for_backward(q, k, v, bias, vbias if vbias is not None else bias.new_zeros(0), out, lse)
        return out

    @staticmethod
    def backward(ctx, d_out):
        _, bwd, fp, ip = _seq_attention_backend()
        q, k, v, bias, vbias, out, lse = ctx.saved_tensors
        vbias = vbias if ctx.has_vbias else None
        plan: SeqAttentionPlan = ctx.plan
        B, H, L, D = q.shape
        d_out = d_out.contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        entries = int(plan.row_key.shape[0])
        d_entry_bias = torch.empty((max(entries, 1), H, D), device=q.device, dtype=torch.float32)
        d_entry_vbias = torch.empty_like(d_entry_bias) if vbias is not None else None
        delta = torch.empty((B, H, L), device=q.device, dtype=torch.float32)
        bwd(fp(q), fp(k), fp(v), ip(plan.lengths), fp(bias), fp(vbias) if vbias is not None else None, ip(plan.row_ptr),
            ip(plan.row_key), ip(plan.row_tab), ip(plan.col_ptr), ip(plan.col_query), ip(plan.col_tab), B, H, L, D,
            plan.num_tables, ctx.dropout[0], ctx.dropout[1], fp(out), fp(lse), fp(d_out), fp(dq), fp(dk), fp(dv),
            fp(d_entry_bias),
            fp(d_entry_vbias) if d_entry_vbias is not None else None, fp(delta))
        tabs = plan.row_tab.long()
        d_bias = torch.zeros_like(bias).index_add_(0, tabs, d_entry_bias[:entries])
        d_vbias = torch.zeros_like(vbias).index_add_(0, tabs, d_entry_vbias[:entries]) if vbias is not None else None
        return dq, dk, dv, d_bias, d_vbias, None, None, None


# Tensor-core path (csrc/seq_attention_tc.cu + csrc/gemm_tma.cu).  BUGLAB_B200_SEQ_TC=0 keeps the fp32 CUDA-core kernels.
SEQ_ATTENTION_TC = os.environ.get("BUGLAB_B200_SEQ_TC", "1") != "0"
_SEQ_HEAD = 64  # the tensor-core path's head size (smaller heads are zero-padded)


def _seq_tc_ok(q: torch.Tensor) -> bool:
    if not (SEQ_ATTENTION_TC and USE_TMA and q.is_cuda):
        return False
    lib = _lib.load()
    L, D = int(q.shape[2]), int(q.shape[3])
    return bool(lib.bl_seq_attention_tc_supported(D, L)) and bool(lib.bl_tma_gemm_supported(_SEQ_HEAD, _SEQ_HEAD)) and \
        bool(lib.bl_tma_weight_grad_supported(128, _SEQ_HEAD))


def _pad_heads(x: torch.Tensor, Lp: int) -> torch.Tensor:
    """[B, H, L, D] -> contiguous [B, H, Lp, 64], zero-padded."""
    B, H, L, D = x.shape
    if L == Lp and D == _SEQ_HEAD:
        return x.contiguous()
    return torch.nn.functional.pad(x, (0, _SEQ_HEAD - D, 0, Lp - L)).contiguous()


class SeqEdgeAttentionTcFn(torch.autograd.Function):
    """The same function as :class:`SeqEdgeAttentionFn` with the four GEMM-shaped products on the TMA-fed tcgen05 kernels
    (split-fp16, fp32-class accuracy), one segment per (sample, head) whose "weight matrix" is that head's K / V / Q / dO,
    and warp-per-row kernels for everything between them (entry terms, mask, softmax, dropout, split tables, entry
    gradients).  The [B*H*Lp, Lp] score tile goes through HBM (537 MB at B=64, H=8, Lp=512: ~0.1 ms per pass)."""

    @staticmethod
    def forward(ctx, q, k, v, bias, vbias, plan: SeqAttentionPlan, p_drop: float = 0.0, seed: int = 0):
        lib = _lib.load()
        B, H, L, D = q.shape
        if (B, L) != (plan.num_samples, plan.max_len) or bias.shape != (plan.num_tables, H, D):
            raise ValueError(f"shape mismatch: q {tuple(q.shape)}, bias {tuple(bias.shape)}, plan B={plan.num_samples} "
                             f"L={plan.max_len} tables={plan.num_tables}")
        dev = q.device
        Lp = 128 if L <= 128 else (256 if L <= 256 else 512)
        G, R = B * H, B * H * Lp
        qp, kp, vp = _pad_heads(q, Lp), _pad_heads(k, Lp), _pad_heads(v, Lp)
        pad_d = (0, _SEQ_HEAD - D)
        bias_p = torch.nn.functional.pad(bias, pad_d).contiguous()
        vbias_p = torch.nn.functional.pad(vbias, pad_d).contiguous() if vbias is not None else None
        seg = torch.arange(G + 1, device=dev, dtype=torch.int32) * Lp
        tiles = segment_units(seg, None, tma_tile_rows(), R)
        slabs = segment_units(seg, None, tma_slab_rows(), R)
        amax_q, amax_k, amax_v = absmax(qp), absmax(kp), absmax(vp)
        q_split = rows_split(qp.view(R, _SEQ_HEAD), None, amax_q)
        scores = tma_project(q_split, None, weight_parts(kp.view(G, Lp, _SEQ_HEAD), Lp, _SEQ_HEAD, 0, False, amax_k), None,
                             amax_q, tiles, R, None, amax_k)                                    # S = Q K^T  [R, Lp]
        amax_p = torch.full((1,), 1.0 / (1.0 - float(p_drop)), device=dev, dtype=torch.float32)
        p_split = torch.empty((2, R + 1, Lp), device=dev, dtype=torch.float16)
        lse = torch.empty(R, device=dev, dtype=torch.float32)
        o_extra = torch.empty((R, _SEQ_HEAD), device=dev, dtype=torch.float32) if vbias is not None else None
        check(lib.bl_seq_softmax_fwd(f32(scores), f32(qp), i32(plan.lengths), f32(bias_p),
                                     f32(vbias_p) if vbias_p is not None else None, i32(plan.row_ptr), i32(plan.row_key),
                                     i32(plan.row_tab), B, H, L, Lp, plan.num_tables, float(p_drop), int(seed), f32(amax_p),
                                     f32(lse), p_split.data_ptr(), f32(o_extra) if o_extra is not None else None,
                                     stream_ptr(dev)), "bl_seq_softmax_fwd")
        out_p = tma_project(p_split, None, weight_parts(vp.view(G, Lp, _SEQ_HEAD), _SEQ_HEAD, Lp, 0, True, amax_v), None, amax_p,
                            tiles, R, None, amax_v)                                             # O = P' V  [R, 64]
        if o_extra is not None:
            out_p += o_extra
        ctx.plan, ctx.dims = plan, (B, H, L, D, Lp)
        ctx.dropout = (float(p_drop), int(seed))
        ctx.has_vbias = vbias is not None
        ctx.units = (tiles, slabs)
        ctx.save_for_backward(qp, kp, vp, bias_p, vbias_p if vbias_p is not None else bias_p.new_zeros(0), scores, lse, out_p,
                              amax_q, amax_k, amax_v, amax_p)
        return out_p.view(B, H, Lp, _SEQ_HEAD)[:, :, :L, :D].contiguous()

    @staticmethod
    def backward(ctx, d_out):
        lib = _lib.load()
        qp, kp, vp, bias_p, vbias_p, scores, lse, out_p, amax_q, amax_k, amax_v, amax_p = ctx.saved_tensors
        vbias_p = vbias_p if ctx.has_vbias else None
        plan: SeqAttentionPlan = ctx.plan
        B, H, L, D, Lp = ctx.dims
        tiles, slabs = ctx.units
        dev = qp.device
        G, R = B * H, B * H * Lp
        gp = _pad_heads(d_out, Lp).view(R, _SEQ_HEAD)
        amax_g = absmax(gp)
        g_split = rows_split(gp, None, amax_g)
        d_scores = tma_project(g_split, None, weight_parts(vp.view(G, Lp, _SEQ_HEAD), Lp, _SEQ_HEAD, 0, False, amax_v), None,
                               amax_g, tiles, R, None, amax_v)                                  # dP = dO V^T  [R, Lp]
        entries = int(plan.row_key.shape[0])
        p_split = torch.empty((2, R + 1, Lp), device=dev, dtype=torch.float16)
        dq_extra = torch.empty((R, _SEQ_HEAD), device=dev, dtype=torch.float32)
        d_entry_bias = torch.empty((max(entries, 1), H, D), device=dev, dtype=torch.float32)
        d_entry_vbias = torch.empty_like(d_entry_bias) if vbias_p is not None else None
        check(lib.bl_seq_softmax_bwd(f32(scores), f32(lse), f32(qp), i32(plan.lengths), f32(bias_p),
                                     f32(vbias_p) if vbias_p is not None else None, i32(plan.row_ptr), i32(plan.row_key),
                                     i32(plan.row_tab), B, H, L, Lp, plan.num_tables, ctx.dropout[0], ctx.dropout[1], f32(amax_p),
                                     f32(out_p), f32(gp), f32(d_scores), p_split.data_ptr(), f32(dq_extra), f32(d_entry_bias),
                                     f32(d_entry_vbias) if d_entry_vbias is not None else None, D, stream_ptr(dev)),
              "bl_seq_softmax_bwd")                                                             # d_scores now holds dS
        amax_ds = absmax(d_scores)
        ds_split = rows_split(d_scores, None, amax_ds)
        del d_scores
        dq = tma_project(ds_split, None, weight_parts(kp.view(G, Lp, _SEQ_HEAD), _SEQ_HEAD, Lp, 0, True, amax_k), None, amax_ds,
                         tiles, R, None, amax_k)                                                # dQ = dS K  [R, 64]
        dq += dq_extra
        identity = torch.arange(R, device=dev, dtype=torch.int32)
        q_split = rows_split(qp.view(R, _SEQ_HEAD), None, amax_q)
        dk = torch.empty((G, Lp, _SEQ_HEAD), device=dev, dtype=torch.float32)
        tma_weight_grad(ds_split, q_split, identity, amax_ds, slabs, dk, 0, amax_q)             # dK = dS^T Q
        dv = torch.empty((G, Lp, _SEQ_HEAD), device=dev, dtype=torch.float32)
        tma_weight_grad(p_split, g_split, identity, amax_p, slabs, dv, 0, amax_g)               # dV = P'^T dO

        def unpad(x):
            return x.view(B, H, Lp, _SEQ_HEAD)[:, :, :L, :D].contiguous()

        tabs = plan.row_tab.long()
        d_bias = torch.zeros((plan.num_tables, H, D), device=dev, dtype=torch.float32).index_add_(0, tabs, d_entry_bias[:entries])
        d_vbias = (torch.zeros((plan.num_tables, H, D), device=dev, dtype=torch.float32).index_add_(0, tabs, d_entry_vbias[:entries])
                   if vbias_p is not None else None)
        return unpad(dq), unpad(dk), unpad(dv), d_bias, d_vbias, None, None, None


def seq_edge_attention(q, k, v, bias, vbias, plan: SeqAttentionPlan, p_drop: float = 0.0, training: bool = False) -> torch.Tensor:
    """``p_drop``: dropout on the attention probabilities (active when ``training``); the mask is a pure function of a fresh
    seed and the (sample, head, query, key) index, recomputed in backward."""
    p = float(p_drop) if training else 0.0
    fn = SeqEdgeAttentionTcFn if _seq_tc_ok(q) else SeqEdgeAttentionFn
    return fn.apply(q, k, v, bias, vbias, plan, p, fresh_seed() if p > 0 else 0)


# ---------------------------------------------------------------------------------------------------
# Flat-buffer optimiser (A12)
# ---------------------------------------------------------------------------------------------------
def grad_sqnorm(flat_grad: torch.Tensor, out: torch.Tensor, partial: torch.Tensor) -> None:
    check(_lib.load().bl_grad_sqnorm(f32(flat_grad), flat_grad.numel(), f32(out), f32(partial),
                                      stream_ptr(flat_grad.device)), "bl_grad_sqnorm")


def adam_step(param, grad, exp_avg, exp_avg_sq, lr, beta1, beta2, eps, step, max_norm, sqnorm, grad_scale=1.0) -> None:
    check(_lib.load().bl_adam_step(f32(param), f32(grad), f32(exp_avg), f32(exp_avg_sq), param.numel(), lr, beta1, beta2,
                                    eps, step, max_norm, f32(sqnorm) if sqnorm is not None else None, grad_scale,
                                    stream_ptr(param.device)), "bl_adam_step")
