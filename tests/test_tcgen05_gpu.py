"""GPU tests of the hand-written tcgen05/TMEM grouped projection kernel (csrc/pair_project_tc.cu) and of the split-fp16
dense linear, against fp64 references.  Tolerance: 1e-4 abs+rel as everywhere (errors are ~1e-6)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ragged_type_ptr(counts):
    tp = [0]
    for c in counts:
        tp.append(tp[-1] + c)
    return tp


@pytest.mark.parametrize("counts,k_in,n_out,use_idx,use_bias", [
    ([5], 64, 128, False, False),                      # a single partial tile
    ([128, 0, 300, 1], 64, 128, True, True),           # exact tile, empty type, ragged tails
    ([700, 33, 0, 260], 128, 256, True, True),         # 2 K chunks (both smem stages), N = 256
    ([1000, 515], 256, 256, True, False),              # 4 K chunks: stage reuse + mbarrier phase flips
    ([400, 77], 256, 512, True, True),                 # N = 512 -> two 256-column work items per row tile
    ([5000, 3000, 2500], 512, 256, True, True),        # more tiles than SMs would need per type; 8 K chunks
])
def test_pair_project_tc_matches_fp64(cuda_device, counts, k_in, n_out, use_idx, use_bias):
    from buglab_b200 import ops

    g = torch.Generator().manual_seed(sum(counts) + k_in + n_out)
    K = len(counts)
    tp = _ragged_type_ptr(counts)
    P = tp[-1]
    n_src = 3000
    src = torch.randn(n_src if use_idx else P, k_in, generator=g)
    idx = torch.randint(0, n_src, (P,), generator=g, dtype=torch.int32) if use_idx else None
    col0, ld = 8, k_in + 24
    weight = torch.randn(K, n_out, ld, generator=g) / k_in ** 0.5
    bias = torch.randn(K, n_out, generator=g) if use_bias else None
    rows = src[idx.long()] if use_idx else src
    ref = torch.empty(P, n_out, dtype=torch.float64)
    for k in range(K):
        lo, hi = tp[k], tp[k + 1]
        ref[lo:hi] = rows[lo:hi].double() @ weight[k, :, col0:col0 + k_in].double().t()
        if use_bias:
            ref[lo:hi] += bias[k].double()

    dev = cuda_device
    parts = ops.weight_parts(weight.to(dev), n_out, k_in, col0, transposed=False)
    out = ops.pair_project_tc(src.to(dev), idx.to(dev) if use_idx else None, parts, bias.to(dev) if use_bias else None,
                              torch.tensor(tp, dtype=torch.int32, device=dev), P)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu().double(), ref, atol=1e-4, rtol=1e-4)
    assert float((out.cpu().double() - ref).abs().max()) < 2e-5  # fp32-class accuracy, not merely inside the budget


def test_pair_project_tc_transposed_with_prescale(cuda_device):
    """Backward-with-respect-to-rows use: src = a tiny-valued gradient table (needs the pow2 pre-scale), W transposed."""
    from buglab_b200 import _lib, ops

    g = torch.Generator().manual_seed(3)
    counts, M, D = [900, 0, 450], 256, 128
    tp = _ragged_type_ptr(counts)
    P = tp[-1]
    grad = torch.randn(P, M, generator=g) * 3e-6
    weight = torch.randn(len(counts), M, 2 * D, generator=g) / (2 * D) ** 0.5   # Linear_k.weight = [A_k | B_k]
    ref = torch.empty(P, D, dtype=torch.float64)
    for k in range(len(counts)):
        ref[tp[k]:tp[k + 1]] = grad[tp[k]:tp[k + 1]].double() @ weight[k, :, D:].double()   # d(rows) = dV @ B_k
    dev = cuda_device
    gd = grad.to(dev)
    amax = torch.empty(1, device=dev)
    _lib.check(_lib.load().bl_absmax(gd.data_ptr(), gd.numel(), amax.data_ptr(), torch.cuda.current_stream().cuda_stream), "bl_absmax")
    assert abs(float(amax) - float(grad.abs().max())) < 1e-12
    parts = ops.weight_parts(weight.to(dev), D, M, D, transposed=True)   # [K, 2, n_out=D, k_in=M]
    out = ops.pair_project_tc(gd, None, parts, None, torch.tensor(tp, dtype=torch.int32, device=dev), P, amax=amax)
    _lib.check(_lib.load().bl_unscale_pow2(out.data_ptr(), out.numel(), amax.data_ptr(), torch.cuda.current_stream().cuda_stream), "unscale")
    err = (out.cpu().double() - ref).abs().max() / ref.abs().max()
    assert float(err) < 1e-5, float(err)


@pytest.mark.parametrize("counts,M,D,col0", [
    ([40], 128, 256, 0),                       # a single partial chunk
    ([64, 0, 1000, 129], 128, 256, 256),       # exact chunk, empty type, several chunks (both stages), column offset
    ([9000, 300], 256, 256, 0),                # > 4096 rows: three slabs accumulate into the same block (REDs), 2 m tiles
    ([2500, 1200, 700], 512, 512, 512),        # the wide layer: 4 m tiles x 2 n tiles
])
def test_pair_weight_grad_tc_matches_fp64(cuda_device, counts, M, D, col0):
    from buglab_b200 import _lib, ops

    g_ = torch.Generator().manual_seed(sum(counts) + M)
    K = len(counts)
    tp = _ragged_type_ptr(counts)
    P = tp[-1]
    n_src = 2000
    grad = torch.randn(P, M, generator=g_) * 1e-3
    x = torch.randn(n_src, D, generator=g_)
    idx = torch.randint(0, n_src, (P,), generator=g_, dtype=torch.int32)
    ref = torch.zeros(K, M, D, dtype=torch.float64)
    for k in range(K):
        lo, hi = tp[k], tp[k + 1]
        ref[k] = grad[lo:hi].double().t() @ x[idx[lo:hi].long()].double()
    dev = cuda_device
    gd = grad.to(dev)
    amax = torch.empty(1, device=dev)
    _lib.check(_lib.load().bl_absmax(gd.data_ptr(), gd.numel(), amax.data_ptr(), torch.cuda.current_stream().cuda_stream), "bl_absmax")
    ld = 2 * D + (256 if col0 else 0)
    d_weight = torch.full((K, M, ld), 7.0, device=dev)  # untouched columns must keep their contents
    ops.pair_weight_grad_tc(gd, x.to(dev), idx.to(dev), amax, torch.tensor(tp, dtype=torch.int32, device=dev), d_weight, col0)
    torch.cuda.synchronize()
    got = d_weight[:, :, col0:col0 + D].cpu().double()
    scale = float(ref.abs().max())
    # tensor-core accumulation truncates: error grows with the chain length (<= 4096 rows = 768 MMAs per partial)
    assert float((got - ref).abs().max()) < 3e-5 * scale + 1e-7, (float((got - ref).abs().max()), scale)
    outside = torch.ones(ld, dtype=torch.bool); outside[col0:col0 + D] = False
    assert torch.all(d_weight[:, :, outside.to(dev)] == 7.0)


@pytest.mark.parametrize("R,K_in,N_out", [(1000, 256, 256), (777, 512, 256), (64, 32, 16)])
@pytest.mark.parametrize("exact_forward", [True, False])
def test_dense_linear_f16x3(cuda_device, monkeypatch, exact_forward, R, K_in, N_out):
    from buglab_b200 import ops
    from oracle import parity

    monkeypatch.setattr(ops.DenseLinearF16x3, "exact_forward", exact_forward)

    g = torch.Generator().manual_seed(R)
    x = torch.randn(R, K_in, generator=g)
    w = torch.randn(N_out, K_in, generator=g) / K_in ** 0.5
    dy = torch.randn(R, N_out, generator=g) * 1e-3
    xr, wr = x.double().requires_grad_(), w.double().requires_grad_()
    (xr @ wr.t()).backward(dy.double())
    xg, wg = x.to(cuda_device).requires_grad_(), w.to(cuda_device).requires_grad_()
    y = ops.DenseLinearF16x3.apply(xg, wg)
    y.backward(dy.to(cuda_device))
    torch.testing.assert_close(y.detach().cpu().double(), (x.double() @ w.double().t()), atol=1e-4, rtol=1e-4)
    assert float((y.detach().cpu().double() - x.double() @ w.double().t()).abs().max()) < 2e-5
    parity.assert_grad_close(xg.grad, xr.grad, "dx", max_frac_bad=0.0, max_rel_l2=1e-5)
    parity.assert_grad_close(wg.grad, wr.grad, "dw", max_frac_bad=0.0, max_rel_l2=1e-5)


def test_layer_with_tcgen05_projection_matches_oracle(cuda_device, monkeypatch):
    """The whole typed-edge layer with the tcgen05 forward projection switched on."""
    from buglab_b200 import ops
    from oracle.mp_ref import typed_edge_message_max_ref
    from tests.graphgen import random_adjacency

    monkeypatch.setattr(ops, "PROJECTION_MODE", "f16x3")
    monkeypatch.setattr(ops, "USE_TCGEN05", True)
    N, D, M, K = 2000, 128, 256, 5
    adj = random_adjacency(N, K, [9000, 3000, 50], seed=77, self_edges=False)
    g = torch.Generator().manual_seed(9)
    h = torch.randn(N, D, generator=g)
    w = torch.randn(K, M, 2 * D, generator=g) / (2 * D) ** 0.5
    b = torch.randn(K, M, generator=g) * 0.1
    agg_ref, _ = typed_edge_message_max_ref(h.double(), adj, w.double(), b.double())
    plan = ops.build_edge_plan([(s.to(cuda_device), t.to(cuda_device)) for s, t in adj], N)
    agg = ops.typed_edge_message_max(h.to(cuda_device), w.to(cuda_device), b.to(cuda_device), plan)
    torch.testing.assert_close(agg.cpu(), agg_ref.float(), atol=1e-4, rtol=1e-4)
