"""Micro-benchmark (GPU): tcgen05 fused projection vs split kernel + cuBLAS, realistic c2 shapes."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch
from buglab_b200 import ops, _lib
dev = torch.device("cuda:0")
lib = _lib.load()
def timeit(fn, reps=5):
    for _ in range(2): fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(reps): fn()
    e.record(); e.synchronize()
    return s.elapsed_time(e) / reps
for (P, N, D, M, K) in ((5_000_000, 564_508, 256, 256, 17), (5_000_000, 564_508, 512, 512, 17)):
    g = torch.Generator(device=dev).manual_seed(0)
    h = torch.randn(N, D, device=dev, generator=g)
    idx = torch.randint(0, N, (P,), device=dev, generator=g, dtype=torch.int32)
    # realistic locality: mostly ascending node ids per type
    idx = torch.sort(idx.view(K, -1) if P % K == 0 else idx[:P - P % K].view(K, -1), dim=1)[0].reshape(-1).contiguous()
    P = idx.shape[0]
    w = torch.randn(K, M, 2 * D, device=dev, generator=g) / (2 * D) ** 0.5
    b = torch.randn(K, M, device=dev, generator=g)
    tp = tuple(int(i * (P // K)) for i in range(K)) + (P,)
    tp_dev = torch.tensor(tp, dtype=torch.int32, device=dev)
    def cublas_path():
        a3 = ops._split3_rows(h, idx)
        w3, _ = ops._split3_weights(w, b, D, D, True, False)
        out = torch.empty(P, M, device=dev)
        _lib.check(lib.bl_pair_project_fwd(a3.data_ptr(), w3.data_ptr(), ops._host_i32(tp), K, M, D, out.data_ptr(), torch.cuda.current_stream().cuda_stream), "f")
        return out
    def tc_path():
        parts = ops.weight_parts(w, M, D, D, False)
        return ops.pair_project_tc(h, idx, parts, b, tp_dev, P)
    o1, o2 = cublas_path(), tc_path()
    print(f"P={P} D={D} M={M}: max|diff| {float((o1-o2).abs().max()):.2e}; split+cuBLAS {timeit(cublas_path):.2f} ms; tcgen05 fused {timeit(tc_path):.2f} ms; "
          f"useful flops 2*P*3D*M = {2*P*3*D*M/1e12:.2f} TF")
    del o1, o2

# ---- weight gradient: tcgen05 kernel vs split2 + cuBLAS ----
for (P, N, D, M, K) in ((5_000_000, 564_508, 256, 256, 17),):
    g = torch.Generator(device=dev).manual_seed(1)
    h = torch.randn(N, D, device=dev, generator=g)
    idx = torch.sort(torch.randint(0, N, (P - P % K,), device=dev, generator=g, dtype=torch.int32).view(K, -1), dim=1)[0].reshape(-1).contiguous()
    P = idx.shape[0]
    grad = torch.randn(P, M, device=dev, generator=g) * 1e-3
    tp = tuple(int(i * (P // K)) for i in range(K)) + (P,)
    tp_dev = torch.tensor(tp, dtype=torch.int32, device=dev)
    amax = torch.empty(1, device=dev)
    _lib.check(lib.bl_absmax(grad.data_ptr(), grad.numel(), amax.data_ptr(), torch.cuda.current_stream().cuda_stream), "a")
    dw1 = torch.zeros(K, M, 2 * D, device=dev); dw2 = torch.zeros(K, M, 2 * D, device=dev)
    def cublas_wg():
        g2 = ops._split2_rows(grad, None, amax)
        a2 = ops._split2_rows(h, idx)
        tmp = torch.empty(K, 2 * M, 2 * D, device=dev)
        _lib.check(lib.bl_pair_project_bwd_weight(g2.data_ptr(), 2 * M, 0, a2.data_ptr(), ops._host_i32(tp), K, M, D, amax.data_ptr(), tmp.data_ptr(), dw1.data_ptr(), 2 * D, 0, torch.cuda.current_stream().cuda_stream), "w")
    def tc_wg():
        ops.pair_weight_grad_tc(grad, h, idx, amax, tp_dev, dw2, 0)
    cublas_wg(); tc_wg()
    print(f"weight grad P={P} D={D} M={M}: max|diff| {float((dw1[:, :, :D]-dw2[:, :, :D]).abs().max()):.2e} (scale {float(dw1.abs().max()):.2e}); "
          f"split2+cuBLAS {timeit(cublas_wg):.2f} ms; tcgen05 {timeit(tc_wg):.2f} ms")
