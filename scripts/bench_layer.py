"""One message-passing layer on B200: this repo's path vs PyTorch-eager (index_select + Linear + GELU + scatter_reduce amax),
for BASELINE.json configs[1] (c2 shape) and configs[2] (c3: 1M nodes / 10M edges / 14 edge kinds packed batch).
Prints one JSON line per config; numbers are CUDA-event timings after warm-up on inputs far larger than L2."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import numpy as np
import torch
from buglab_b200 import _lib, ops
from buglab_b200.synthetic import packed_edge_batch

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0


def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(reps): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / reps


def run(name, N, E, K, D, M, seed=0):
    src, tgt, etype = packed_edge_batch(N, E, K, seed)
    src_d, tgt_d, et_d = (torch.from_numpy(a).to(dev) for a in (src, tgt, etype))
    plan = ops.build_edge_plan_from_flat(src_d, tgt_d, et_d, N, K)
    g = torch.Generator(device=dev).manual_seed(seed)
    h = torch.randn(N, D, device=dev, generator=g)
    w = (torch.randn(K, M, 2 * D, device=dev, generator=g) / (2 * D) ** 0.5)
    b = torch.randn(K, M, device=dev, generator=g) * 0.1
    lib = _lib.load()
    # --- fused edge kernel alone (the roofline kernel) ---
    u = torch.randn(plan.num_s_pairs, M, device=dev, generator=g); v = torch.randn(plan.num_t_pairs, M, device=dev, generator=g)
    agg = torch.empty(N, M, device=dev); xw = torch.empty_like(agg); ew = torch.empty(N, M, device=dev, dtype=torch.int32)
    st = torch.cuda.current_stream().cuda_stream
    def edge_fwd():
        _lib.check(lib.bl_edge_segmax_fwd(u.data_ptr(), v.data_ptr(), plan.row_ptr.data_ptr(), plan.urow.data_ptr(), plan.vrow.data_ptr(), N, M, agg.data_ptr(), xw.data_ptr(), ew.data_ptr(), st), "f")
    ms_edge = timeit(edge_fwd, reps=10, warm=3)
    algo = E * (2 * M * 4 + 12) + N * M * 4
    # --- whole typed-edge message+aggregate step, forward and forward+backward (ours) ---
    hq, wq, bq = h.clone().requires_grad_(), w.clone().requires_grad_(), b.clone().requires_grad_()
    dout = torch.randn(N, M, device=dev, generator=g)
    def ours_fwd():
        with torch.no_grad():
            return ops.typed_edge_message_max(hq, wq, bq, plan)
    def ours_fwd_bwd():
        hq.grad = wq.grad = bq.grad = None
        ops.typed_edge_message_max(hq, wq, bq, plan).backward(dout)
    ms_ours_f, ms_ours_fb = timeit(ours_fwd), timeit(ours_fwd_bwd)
    # --- PyTorch eager on the same GPU (the reference formulation: per-edge gather + Linear, then scatter amax) ---
    type_ptr = np.concatenate(([0], np.cumsum(np.bincount(etype, minlength=K))))
    srcl, tgtl = src_d.long(), tgt_d.long()
    def eager(hh, ww, bb):
        msgs = []
        for k in range(K):
            lo, hi = int(type_ptr[k]), int(type_ptr[k + 1])
            if hi > lo:
                x = torch.cat((hh.index_select(0, srcl[lo:hi]), hh.index_select(0, tgtl[lo:hi])), dim=1)
                msgs.append(torch.nn.functional.linear(x, ww[k], bb[k]))
        m = torch.nn.functional.gelu(torch.cat(msgs, 0))
        out = torch.zeros(N, M, device=dev).scatter_reduce(0, tgtl.view(-1, 1).expand(-1, M), m, "amax", include_self=False)
        return out
    def eager_fwd():
        with torch.no_grad():
            return eager(hq, wq, bq)
    def eager_fwd_bwd():
        hq.grad = wq.grad = bq.grad = None
        eager(hq, wq, bq).backward(dout)
    ms_eager_f, ms_eager_fb = timeit(eager_fwd, reps=3, warm=1), timeit(eager_fwd_bwd, reps=3, warm=1)
    ok = float((ours_fwd() - eager_fwd()).abs().max())
    print(json.dumps({
        "config": name, "nodes": N, "edges": E, "edge_kinds": K, "D_in": D, "M": M, "s_pairs": plan.num_s_pairs, "t_pairs": plan.num_t_pairs,
        "edge_kernel": {"ms": ms_edge, "algorithmic_bytes": algo, "achieved_GBs": algo / ms_edge / 1e6, "peak_GBs": PEAK, "frac": algo / ms_edge / 1e6 / PEAK},
        "message_aggregate_fwd_ms": {"ours": ms_ours_f, "torch_eager_scatter_reduce": ms_eager_f, "speedup": ms_eager_f / ms_ours_f},
        "message_aggregate_fwd_bwd_ms": {"ours": ms_ours_fb, "torch_eager_scatter_reduce": ms_eager_fb, "speedup": ms_eager_fb / ms_ours_fb},
        "max_abs_diff_ours_vs_eager": ok}), flush=True)
    del u, v, agg, xw, ew


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("c2", "all"):
        run("c2-like: H->H layer of configs[1] (256 graphs)", 564_508, 6_410_926, 17, 256, 256)
    if which in ("c3", "all"):
        run("c3: configs[2] 1M nodes / 10M edges / 14 edge kinds, hidden 256", 1_000_000, 10_000_000, 14, 256, 256)
