#!/usr/bin/env python
"""Micro-benchmark of the projection GEMMs at the c2 shape (BASELINE configs[1]: N = 564 508 nodes, ~5 M pair rows per
table, 17 edge types): first-generation tcgen05 kernels (csrc/pair_project_tc.cu) vs the TMA-fed kernels
(csrc/gemm_tma.cu, CTA pairs unless BUGLAB_B200_TMA_CG=1).  One JSON line per shape; times are CUDA-event means over
``--iters`` launches after ``--warmup``, inputs (> 1 GB) far larger than L2.

    python scripts/bench_tma_gemm.py [--rows 5000000] [--nodes 564508] [--types 17] [--iters 10]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def timed(fn, warmup, iters):
    import torch

    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(iters):
        fn()
    end.record()
    end.synchronize()
    return start.elapsed_time(end) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=5_000_000)
    ap.add_argument("--nodes", type=int, default=564_508)
    ap.add_argument("--types", type=int, default=17)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--shapes", default="256x256,512x512")
    ap.add_argument("--skip-old", action="store_true")
    ap.add_argument("--block-nodes", type=int, default=0, help="order the pair rows by (node block, type, node) like the model's plan")
    args = ap.parse_args()

    import torch

    from buglab_b200 import _lib, ops

    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    K, P, N = args.types, args.rows, args.nodes
    # pair rows ordered by (type, node) like the plan's tables: long-tailed type sizes, sorted node ids inside a type
    w = torch.tensor([1.0 / (k + 1) ** 0.7 for k in range(K)])
    counts = (w / w.sum() * P).long()
    counts[0] += P - int(counts.sum())
    type_ptr_host = [0]
    for c in counts.tolist():
        type_ptr_host.append(type_ptr_host[-1] + c)
    idx = torch.cat([torch.sort(torch.randint(0, N, (int(c),), dtype=torch.int32))[0] for c in counts.tolist()]).to(dev)
    type_ptr = torch.tensor(type_ptr_host, dtype=torch.int32, device=dev)
    if args.block_nodes > 0:
        raise SystemExit("--block-nodes: use bench.py's roofline legs (they time the kernels on the model's blocked plan)")
    cg = os.environ.get("BUGLAB_B200_TMA_CG", "2")

    for shape in args.shapes.split(","):
        D, M = (int(v) for v in shape.split("x"))
        h = torch.randn(N, D, device=dev)
        weight = torch.randn(K, M, 2 * D, device=dev) / (2 * D) ** 0.5
        bias = torch.randn(K, M, device=dev)
        out = {"shape": f"D={D} M={M}", "rows": P, "types": K, "cta_group": int(cg)}
        flops = 2.0 * P * D * M * 3

        # ---- forward projection (gathered rows) ----
        parts = ops.weight_parts(weight, M, D, 0, False)
        h_split = ops.rows_split(h)
        tiles = ops.segment_units(type_ptr, None, ops.tma_tile_rows(), P)
        slabs = ops.segment_units(type_ptr, None, ops.tma_slab_rows(), P)
        new = ops.tma_project(h_split, idx, parts, bias, None, tiles, P)
        out["split_ms"] = timed(lambda: ops.rows_split(h), args.warmup, args.iters)
        out["fwd_tma_ms"] = timed(lambda: ops.tma_project(h_split, idx, parts, bias, None, tiles, P), args.warmup, args.iters)
        out["fwd_tma_tflops"] = flops / out["fwd_tma_ms"] / 1e9
        if _lib.load().bl_tma_project_stationary_supported(M, D):
            stat = ops.tma_project(h_split, idx, parts, bias, None, tiles, P, slabs)
            out["fwd_stationary_max_abs_diff_vs_streaming"] = float((stat - new).abs().max())
            out["fwd_tma_stationary_ms"] = timed(lambda: ops.tma_project(h_split, idx, parts, bias, None, tiles, P, slabs),
                                                 args.warmup, args.iters)
            out["fwd_tma_stationary_tflops"] = flops / out["fwd_tma_stationary_ms"] / 1e9
            del stat
        if not args.skip_old and _lib.load().bl_pair_project_tc_supported(M, D):
            old = ops.pair_project_tc(h, idx, parts, bias, type_ptr, P)
            out["fwd_max_abs_diff_vs_gen1"] = float((old - new).abs().max())
            out["fwd_gen1_ms"] = timed(lambda: ops.pair_project_tc(h, idx, parts, bias, type_ptr, P), args.warmup, args.iters)
            del old
        sample = torch.randint(0, P, (2048,), device=dev)
        seg_of = torch.searchsorted(type_ptr[1:].long(), sample, right=True)
        ref = torch.stack([h[idx[p].long()].double() @ weight[k, :, :D].double().t() + bias[k].double()
                           for p, k in zip(sample.tolist(), seg_of.tolist())])
        out["fwd_max_abs_err_vs_fp64_sample"] = float((new[sample].double() - ref).abs().max())
        del new

        # ---- backward w.r.t. the rows (contiguous gradient table, transposed weights) ----
        g = torch.randn(P, M, device=dev) * 1e-4
        amax = torch.empty(1, device=dev)
        _lib.check(_lib.load().bl_absmax(_lib.f32(g), g.numel(), _lib.f32(amax), _lib.stream_ptr(dev)), "bl_absmax")
        parts_t = ops.weight_parts(weight, D, M, 0, True)
        g_split = ops.rows_split(g, None, amax)
        out["gsplit_ms"] = timed(lambda: ops.rows_split(g, None, amax), args.warmup, args.iters)
        d_in = ops.tma_project(g_split, None, parts_t, None, amax, tiles, P)
        out["bwd_in_tma_ms"] = timed(lambda: ops.tma_project(g_split, None, parts_t, None, amax, tiles, P), args.warmup, args.iters)
        if _lib.load().bl_tma_project_stationary_supported(D, M):
            out["bwd_in_tma_stationary_ms"] = timed(lambda: ops.tma_project(g_split, None, parts_t, None, amax, tiles, P, slabs),
                                                    args.warmup, args.iters)
        if not args.skip_old and _lib.load().bl_pair_project_tc_supported(D, M):
            old = ops.pair_project_tc(g, None, parts_t, None, type_ptr, P, amax=amax)
            scale = float(torch.exp2(12 - torch.ceil(torch.log2(amax))))
            out["bwd_in_rel_diff_vs_gen1"] = float((old / scale - d_in).abs().max() / d_in.abs().max())
            out["bwd_in_gen1_ms"] = timed(lambda: ops.pair_project_tc(g, None, parts_t, None, type_ptr, P, amax=amax),
                                          args.warmup, args.iters)
            del old
        del d_in

        # ---- weight gradient ----
        d_weight = torch.zeros(K, M, 2 * D, device=dev)
        ops.tma_weight_grad(g_split, h_split, idx, amax, slabs, d_weight, 0)
        out["wgrad_tma_ms"] = timed(lambda: ops.tma_weight_grad(g_split, h_split, idx, amax, slabs, d_weight, 0), args.warmup, args.iters)
        k_small = K - 1  # the smallest type: cheap exact reference
        lo, hi = type_ptr_host[k_small], type_ptr_host[k_small + 1]
        ref = g[lo:hi].double().t() @ h[idx[lo:hi].long()].double()
        out["wgrad_rel_err_vs_fp64_smallest_type"] = float((d_weight[k_small, :, :D].double() - ref).abs().max() / ref.abs().max())
        if not args.skip_old and _lib.load().bl_pair_weight_grad_tc_supported(M, D):
            d_old = torch.zeros_like(d_weight)
            ops.pair_weight_grad_tc(g, h, idx, amax, type_ptr, d_old, 0)
            out["wgrad_rel_diff_vs_gen1"] = float((d_old[:, :, :D] - d_weight[:, :, :D]).abs().max() / d_weight[:, :, :D].abs().max())
            out["wgrad_gen1_ms"] = timed(lambda: ops.pair_weight_grad_tc(g, h, idx, amax, type_ptr, d_old, 0), args.warmup, args.iters)
            del d_old
        print(json.dumps(out), flush=True)
        del h, weight, g, g_split, h_split, d_weight


if __name__ == "__main__":
    main()
