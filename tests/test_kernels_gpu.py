"""GPU parity tests: every C-ABI kernel against the CPU oracle on seeded inputs (run with -m gpu).

Tolerances: indices bit-exact; fp32 values atol=rtol=1e-4 (BASELINE.json north_star), usually far tighter.
"""
import numpy as np
import pytest
import torch

from tests.graphgen import flatten, random_adjacency

pytestmark = pytest.mark.gpu

TOL = dict(atol=1e-4, rtol=1e-4)


def _plan(adj, N, dev):
    from buglab_b200 import ops

    return ops.build_edge_plan([(s.to(dev), t.to(dev)) for s, t in adj], N)


@pytest.mark.parametrize("N,K,epk,self_edges", [
    (1, 1, [0], True), (7, 3, [5, 0, 9], True), (300, 5, [700, 0, 1500, 3, 64], False),
    (5000, 17, [9000, 4000, 100, 0, 12000], True), (40000, 9, [90000, 30000], True),
])
def test_plan_bit_exact(cuda_device, N, K, epk, self_edges):
    from oracle.plan_ref import build_plan_ref

    adj = random_adjacency(N, K, epk, seed=N + K, self_edges=self_edges)
    src, tgt, et = flatten(adj)
    ref = build_plan_ref(src.numpy(), tgt.numpy(), et.numpy(), N, K)
    plan = _plan(adj, N, cuda_device)
    assert plan.num_s_pairs == ref["num_s_pairs"] and plan.num_t_pairs == ref["num_t_pairs"]
    for name in ("e_perm", "e_src", "e_type", "row_ptr", "urow", "vrow", "s_node", "s_type_ptr", "s_by_node_ptr",
                 "s_by_node_idx", "t_node", "t_type_ptr", "t_by_node_ptr", "t_by_node_idx", "s_edge_ptr", "s_edge_idx",
                 "e_tgt", "s_edge_tgt"):
        got = getattr(plan, name).cpu().numpy()
        assert got.dtype == np.int32
        np.testing.assert_array_equal(got, ref[name], err_msg=name)
    assert plan.s_type_ptr_host == tuple(ref["s_type_ptr"].tolist())
    # node-blocked layout: same edges, pair tables ordered by (node block, type, node)
    from buglab_b200 import ops

    for block in (64, 1 << 20):
        ref = build_plan_ref(src.numpy(), tgt.numpy(), et.numpy(), N, K, block_nodes=block)
        plan = ops.build_edge_plan([(s.to(cuda_device), t.to(cuda_device)) for s, t in adj], N, block_nodes=block)
        assert plan.block_nodes == (block if ops.USE_TMA else 0)
        if plan.block_nodes == 0:
            break
        assert plan.num_segs == int(ref["s_type_ptr"].shape[0]) - 1 and plan.num_s_pairs == ref["num_s_pairs"]
        for name in ("e_perm", "e_src", "e_type", "row_ptr", "urow", "vrow", "s_node", "s_type_ptr", "s_by_node_ptr",
                     "s_by_node_idx", "t_node", "t_type_ptr", "t_by_node_ptr", "t_by_node_idx", "s_edge_ptr", "s_edge_idx",
                     "e_tgt", "s_edge_tgt"):
            np.testing.assert_array_equal(getattr(plan, name).cpu().numpy(), ref[name], err_msg=f"{name} (block {block})")
        np.testing.assert_array_equal(plan.seg_type.cpu().numpy(), np.arange(plan.num_segs) % K)


def test_plan_no_edges(cuda_device):
    from buglab_b200 import ops

    z = torch.zeros(0, dtype=torch.int64, device=cuda_device)
    plan = ops.build_edge_plan([(z, z), (z, z)], 5)
    assert plan.num_edges == 0 and plan.num_s_pairs == 0 and plan.num_t_pairs == 0
    assert plan.row_ptr.cpu().tolist() == [0] * 6


@pytest.mark.parametrize("N,D,M,K,epk,self_edges,use_bias", [
    (9, 8, 12, 3, [11, 0, 4], False, True),      # generic kernels, isolated nodes, empty type
    (257, 32, 32, 4, [900, 300, 5], True, False),  # generic, no bias
    (1500, 64, 128, 6, [4000, 2500, 0, 800], True, True),   # warp kernel ITER=1
    (2000, 128, 256, 5, [9000, 3000, 50], False, True),     # ITER=2, isolated nodes
    (1200, 256, 512, 7, [5000, 2000, 700], True, True),     # ITER=4 (the wide post-residual layer)
    (3000, 256, 256, 9, [9000, 6000, 0, 2500, 40], True, True),   # the bench's H->H layer shape
    (1500, 512, 512, 4, [6000, 3000, 900], True, False),          # the bench's wide layer shape
])
@pytest.mark.parametrize("mode", ["f16x3", "fp32", "f16x3-blocked", "f16x3-blocked-overlap"])
def test_typed_edge_message_max_fwd_bwd(cuda_device, monkeypatch, mode, N, D, M, K, epk, self_edges, use_bias):
    from buglab_b200 import ops
    from oracle.mp_ref import typed_edge_message_max_ref

    block_nodes = 0
    if mode == "f16x3-blocked-overlap":   # the optional side-stream schedule of the by-source backward (BUGLAB_B200_OVERLAP=1)
        mode = "f16x3-blocked"
        monkeypatch.setattr(ops, "OVERLAP_EDGE_BACKWARD", True)
    if mode == "f16x3-blocked":  # node-blocked pair tables (many small segments): only where every product runs on the TMA GEMMs
        mode = "f16x3"
        if ops.plan_block_nodes_for([(D, M)]) == 0:
            pytest.skip("shape not covered by the segment-aware TMA path")
        block_nodes = 192
    monkeypatch.setattr(ops, "PROJECTION_MODE", mode)  # split-fp16 tensor-core GEMMs (default) or fp32 SGEMM

    adj = random_adjacency(N, K, epk, seed=3 * N + M, self_edges=self_edges)
    g = torch.Generator().manual_seed(N + D)
    h = torch.randn(N, D, generator=g)
    w = torch.randn(K, M, 2 * D, generator=g) / (2 * D) ** 0.5
    b = torch.randn(K, M, generator=g) * 0.1 if use_bias else None
    d_out = torch.randn(N, M, generator=g)

    h_ref, w_ref = h.double().requires_grad_(), w.double().requires_grad_()
    b_ref = b.double().requires_grad_() if use_bias else None
    agg_ref, arg_ref = typed_edge_message_max_ref(h_ref, adj, w_ref, b_ref)

    plan = ops.build_edge_plan([(s.to(cuda_device), t.to(cuda_device)) for s, t in adj], N, block_nodes=block_nodes)
    assert plan.block_nodes == block_nodes
    h_g, w_g = h.to(cuda_device).requires_grad_(), w.to(cuda_device).requires_grad_()
    b_g = b.to(cuda_device).requires_grad_() if use_bias else None
    monkeypatch.setattr(ops, "WINNER_TRACE", [])
    agg = ops.typed_edge_message_max(h_g, w_g, b_g, plan)
    winners = ops.WINNER_TRACE[0]          # [N, M] original edge index of the winner (E = no in-edge)
    agg.backward(d_out.to(cuda_device))

    from oracle import parity
    from oracle.mp_ref import edge_messages_ref

    torch.testing.assert_close(agg.cpu(), agg_ref.float(), **TOL)
    # ROUTING, checked against the exact (fp64) messages: the winner the kernel chose must be an in-edge of its node and
    # attain the exact segment maximum (up to 1e-5 relative: two messages closer than fp32 resolution may swap); then the
    # GRADIENTS are compared with the fp64 gradient under that routing — elementwise, no allowance for flips.
    messages, targets = edge_messages_ref(h_ref, adj, w_ref, b_ref)
    E = messages.shape[0]
    valid = winners < E
    assert torch.equal(valid, arg_ref < E)
    picked = messages.gather(0, winners.clamp(max=max(E - 1, 0)))
    assert bool((targets[winners.clamp(max=max(E - 1, 0))] == torch.arange(N).view(-1, 1))[valid].all())
    deficit = ((agg_ref - torch.where(valid, picked, torch.zeros_like(picked))) / (1.0 + agg_ref.abs())).detach()
    assert float(deficit.max()) <= 1e-5, f"a chosen winner misses the exact maximum by {float(deficit.max()):.2e} (relative)"
    print(f"winners differing from the exact argmax: {int((winners != arg_ref).sum())} of {winners.numel()}, "
          f"worst relative deficit {float(deficit.max()):.1e}")
    torch.where(valid, picked, torch.zeros_like(picked)).backward(d_out.double())
    parity.assert_grad_close_to_scale(h_g.grad, h_ref.grad, "d_h")
    parity.assert_grad_close_to_scale(w_g.grad, w_ref.grad, "d_weight")
    if use_bias:
        parity.assert_grad_close_to_scale(b_g.grad, b_ref.grad, "d_bias")
    # isolated nodes aggregate to exactly 0
    deg = torch.zeros(N, dtype=torch.int64).index_add_(0, torch.cat([a[1] for a in adj]), torch.ones(sum(a[1].numel() for a in adj), dtype=torch.int64))
    assert torch.all(agg.cpu()[deg == 0] == 0)


def test_edge_winner_is_first_max_edge(cuda_device):
    """Argmax bookkeeping: duplicate edges tie exactly; the lowest original edge index must win (torch_scatter CPU)."""
    from buglab_b200 import _lib, ops
    from oracle.mp_ref import edge_messages_ref
    from oracle.scatter_ref import scatter_max

    N, D, M, K = 400, 16, 128, 3
    adj = random_adjacency(N, K, [1500, 600], seed=11, self_edges=True, duplicates=True)
    g = torch.Generator().manual_seed(5)
    h = torch.randn(N, D, generator=g)
    w = torch.randn(K, M, 2 * D, generator=g) / (2 * D) ** 0.5
    plan = _plan(adj, N, cuda_device)
    hg, wg = h.to(cuda_device), w.to(cuda_device)
    hs, ht = ops._rows_gather(hg, plan.s_node), ops._rows_gather(hg, plan.t_node)
    u = ops._project_pairs(hs, wg, 0, plan.s_type_ptr_host, None)
    v = ops._project_pairs(ht, wg, D, plan.t_type_ptr_host, None)
    agg = torch.empty(N, M, device=cuda_device); xwin = torch.empty_like(agg)
    ewin = torch.empty(N, M, device=cuda_device, dtype=torch.int32)
    _lib.check(_lib.load().bl_edge_segmax_fwd(u.data_ptr(), v.data_ptr(), plan.row_ptr.data_ptr(), plan.urow.data_ptr(),
                                               plan.vrow.data_ptr(), N, M, agg.data_ptr(), xwin.data_ptr(), ewin.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "fwd")
    # oracle winners from the SAME pre-activations (so only the tie-break / selection logic is compared)
    x_sorted = (u[plan.urow.long()] + v[plan.vrow.long()]).cpu()
    msgs_sorted = torch.nn.functional.gelu(x_sorted)
    perm = plan.e_perm.long().cpu()
    msgs_orig = torch.empty_like(msgs_sorted); msgs_orig[perm] = msgs_sorted
    tgt_orig = torch.cat([a[1] for a in adj])
    val_ref, arg_ref = scatter_max(msgs_orig, tgt_orig, dim=0, dim_size=N)
    win_orig = perm[ewin.long().cpu().clamp(min=0)]
    win_orig[ewin.cpu() < 0] = msgs_orig.shape[0]
    same = win_orig == arg_ref
    # GELU rounding can reorder two messages within an ulp; such rows must still carry the same value
    torch.testing.assert_close(agg.cpu(), val_ref, atol=1e-6, rtol=1e-6)
    assert same.float().mean() > 0.999


@pytest.mark.parametrize("rows,dim", [(1, 4), (37, 128), (1000, 256), (513, 512), (64, 1024), (10, 36)])
def test_layernorm(cuda_device, rows, dim):
    from buglab_b200 import ops

    g = torch.Generator().manual_seed(rows)
    x = torch.randn(rows, dim, generator=g) * 2 + 0.5
    gamma, beta = torch.rand(dim, generator=g) + 0.5, torch.randn(dim, generator=g)
    dy = torch.randn(rows, dim, generator=g)
    xr, gr, br = (t.double().requires_grad_() for t in (x, gamma, beta))
    torch.nn.functional.layer_norm(xr, (dim,), gr, br, 1e-5).backward(dy.double())
    xg, gg, bg = (t.to(cuda_device).requires_grad_() for t in (x, gamma, beta))
    y = ops.layer_norm(xg, gg, bg, 1e-5)
    y.backward(dy.to(cuda_device))
    torch.testing.assert_close(y.cpu(), torch.nn.functional.layer_norm(x, (dim,), gamma, beta, 1e-5), **TOL)
    torch.testing.assert_close(xg.grad.cpu(), xr.grad.float(), **TOL)
    torch.testing.assert_close(gg.grad.cpu(), gr.grad.float(), atol=1e-3, rtol=1e-4)
    torch.testing.assert_close(bg.grad.cpu(), br.grad.float(), atol=1e-3, rtol=1e-4)


def test_tanh_dropout(cuda_device):
    from buglab_b200 import ops

    x = torch.randn(1000, 64)
    xg = x.to(cuda_device).requires_grad_()
    y = ops.tanh_dropout(xg, 0.2, training=False)
    y.backward(torch.ones_like(y))
    torch.testing.assert_close(y.cpu(), torch.tanh(x), **TOL)
    torch.testing.assert_close(xg.grad.cpu(), 1 - torch.tanh(x) ** 2, **TOL)
    # training: kept fraction ~ 0.8, kept values scaled by 1/0.8, backward uses the same mask
    xg2 = x.to(cuda_device).requires_grad_()
    torch.manual_seed(0)
    y2 = ops.tanh_dropout(xg2, 0.2, training=True)
    y2.backward(torch.ones_like(y2))
    kept = (y2 != 0).cpu()
    assert abs(kept.float().mean().item() - 0.8) < 0.01
    torch.testing.assert_close(y2.cpu()[kept], (torch.tanh(x) / 0.8)[kept], **TOL)
    torch.testing.assert_close(xg2.grad.cpu(), torch.where(kept, (1 - torch.tanh(x) ** 2) / 0.8, torch.zeros(())), **TOL)


@pytest.mark.parametrize("L,F,S,is_min", [(0, 1, 3, False), (50, 1, 7, False), (50, 1, 7, True), (400, 128, 33, False), (1000, 5, 1, True)])
def test_segment_minmax(cuda_device, L, F, S, is_min):
    from buglab_b200 import ops
    from oracle import scatter_ref

    g = torch.Generator().manual_seed(L + F)
    src = torch.randn(L, F, generator=g).round(decimals=1)  # rounding creates exact ties
    index = torch.randint(0, max(S - 1, 1), (L,), generator=g)  # last segment stays empty when S > 1
    src1 = src[:, 0] if F == 1 else src
    ref_fn = scatter_ref.scatter_min if is_min else scatter_ref.scatter_max
    sr = src1.clone().requires_grad_()
    out_ref, arg_ref = ref_fn(sr, index, dim=0, dim_size=S)
    d = torch.randn(out_ref.shape, generator=g)
    sg = src1.to(cuda_device).requires_grad_()
    out, arg = ops.segment_minmax(sg, index.to(cuda_device), dim=0, dim_size=S, is_min=is_min)
    if L > 0:
        out_ref.backward(d)
        out.backward(d.to(cuda_device))
    torch.testing.assert_close(out.cpu(), out_ref, atol=0, rtol=0)
    assert torch.equal(arg.cpu(), arg_ref)
    if L > 0:
        torch.testing.assert_close(sg.grad.cpu(), sr.grad, atol=0, rtol=0)


def test_segment_sum_and_log_softmax(cuda_device):
    from buglab_b200 import ops
    from oracle import scatter_ref

    g = torch.Generator().manual_seed(0)
    L, S = 777, 41
    src = torch.randn(L, generator=g) * 3
    index = torch.randint(0, S, (L,), generator=g)
    d = torch.randn(L, generator=g)
    sr = src.double().requires_grad_()
    ref = scatter_ref.scatter_log_softmax(sr, index)
    ref.backward(d.double())
    sg = src.to(cuda_device).requires_grad_()
    out = ops.segment_log_softmax(sg, index.to(cuda_device))
    out.backward(d.to(cuda_device))
    torch.testing.assert_close(out.cpu(), ref.float(), **TOL)
    torch.testing.assert_close(sg.grad.cpu(), sr.grad.float(), **TOL)
    # per group it is log_softmax
    for s in range(3):
        m = index == s
        torch.testing.assert_close(out.cpu()[m], torch.log_softmax(src[m], 0), **TOL)
    s2 = torch.randn(L, 8, generator=g)
    torch.testing.assert_close(ops.segment_sum(s2.to(cuda_device), index.to(cuda_device), dim=0, dim_size=S).cpu(),
                               scatter_ref.scatter_sum(s2, index, dim=0, dim_size=S), **TOL)


def test_subtoken_maxpool(cuda_device):
    from buglab_b200 import ops
    from oracle.mp_ref import subtoken_maxpool_ref

    g = torch.Generator().manual_seed(1)
    V, H, N, T = 500, 128, 3000, 6
    emb = torch.randn(V, H, generator=g)
    ids = torch.randint(0, V, (N, T), generator=g)
    lens = torch.randint(1, T + 1, (N,), generator=g)
    d = torch.randn(N, H, generator=g)
    er = emb.double().requires_grad_()
    ref = subtoken_maxpool_ref(er, ids, lens)
    ref.backward(d.double())
    eg = emb.to(cuda_device).requires_grad_()
    out = ops.subtoken_maxpool(eg, ids.to(cuda_device), lens.to(cuda_device))
    out.backward(d.to(cuda_device))
    torch.testing.assert_close(out.cpu(), ref.float(), atol=0, rtol=0)
    torch.testing.assert_close(eg.grad.cpu(), er.grad.float(), **TOL)


def test_flat_adam_with_clip(cuda_device):
    from buglab_b200 import ops

    g = torch.Generator().manual_seed(2)
    n = 100003
    p0 = torch.randn(n, generator=g)
    ref_p = torch.nn.Parameter(p0.clone().double())
    opt = torch.optim.Adam([ref_p], lr=1e-3)
    p = p0.to(cuda_device).clone(); m = torch.zeros_like(p); v = torch.zeros_like(p)
    sq = torch.zeros(1, device=cuda_device); partial = torch.empty(1024, device=cuda_device)
    for step in range(1, 6):
        grad = torch.randn(n, generator=g) * (3.0 if step % 2 else 0.001)
        ref_p.grad = grad.clone().double()
        torch.nn.utils.clip_grad_norm_([ref_p], 0.5)
        opt.step()
        gg = grad.to(cuda_device)
        ops.grad_sqnorm(gg, sq, partial)
        torch.testing.assert_close(sq.cpu()[0], (grad.double() ** 2).sum().float(), rtol=1e-5, atol=1e-6)
        ops.adam_step(p, gg, m, v, 1e-3, 0.9, 0.999, 1e-8, step, 0.5, sq)
        torch.testing.assert_close(p.cpu(), ref_p.detach().float(), atol=1e-6, rtol=1e-5)


def test_full_size_properties_c3(cuda_device):
    """BASELINE.json configs[2] at FULL size (1M nodes, 10M edges, 14 edge kinds, hidden 256): size-independent properties
    of the plan (permutation, sortedness, pair-table consistency) and of the fused edge kernel (equals a max over the
    per-edge messages evaluated by library ops on the same tables; winners are edges of the right segment)."""
    from buglab_b200 import _lib, ops
    from buglab_b200.synthetic import packed_edge_batch

    N, E, K, M = 1_000_000, 10_000_000, 14, 256
    src, tgt, etype = (torch.from_numpy(a).to(cuda_device) for a in packed_edge_batch(N, E, K, seed=3))
    plan = ops.build_edge_plan_from_flat(src, tgt, etype, N, K)
    perm = plan.e_perm.long()
    # e_perm is a permutation; sorted by (tgt, type, original index)
    assert int(torch.bincount(perm, minlength=E).max()) == 1
    t_sorted, k_sorted = tgt[perm].long(), etype[perm].long()
    key = (t_sorted * K + k_sorted) * E + perm
    assert bool((key[1:] > key[:-1]).all())
    assert torch.equal(plan.e_src, src[perm]) and torch.equal(plan.e_type, etype[perm])
    # CSR: row_ptr is the exclusive prefix sum of the in-degrees
    deg = torch.bincount(tgt.long(), minlength=N)
    assert torch.equal(plan.row_ptr.long(), torch.cat((torch.zeros(1, dtype=torch.int64, device=cuda_device), deg.cumsum(0))))
    # pair tables: every edge points at the pair that holds its (type, node); pairs are unique and (type, node)-sorted
    assert torch.equal(plan.s_node[plan.urow.long()], plan.e_src) and torch.equal(plan.t_node[plan.vrow.long()].long(), t_sorted)
    for node, type_ptr, rows in ((plan.s_node, plan.s_type_ptr, plan.urow), (plan.t_node, plan.t_type_ptr, plan.vrow)):
        tp = type_ptr.long()
        pair_type = torch.bucketize(torch.arange(node.shape[0], device=cuda_device), tp[1:], right=True)
        pkey = pair_type * N + node.long()
        assert bool((pkey[1:] > pkey[:-1]).all())                      # unique and sorted
        assert torch.equal(pair_type[rows.long()], k_sorted)            # an edge's pair has the edge's type
    assert plan.num_s_pairs <= E and plan.num_t_pairs <= E
    # edge kernel at full size vs library ops on the SAME U/V tables
    g = torch.Generator(device=cuda_device).manual_seed(0)
    u = torch.randn(plan.num_s_pairs, M, device=cuda_device, generator=g)
    v = torch.randn(plan.num_t_pairs, M, device=cuda_device, generator=g)
    agg = torch.empty(N, M, device=cuda_device); xwin = torch.empty_like(agg)
    ewin = torch.empty(N, M, device=cuda_device, dtype=torch.int32)
    _lib.check(_lib.load().bl_edge_segmax_fwd(u.data_ptr(), v.data_ptr(), plan.row_ptr.data_ptr(), plan.urow.data_ptr(),
                                               plan.vrow.data_ptr(), N, M, agg.data_ptr(), xwin.data_ptr(), ewin.data_ptr(),
                                               torch.cuda.current_stream().cuda_stream), "bl_edge_segmax_fwd")
    expect = torch.full((N, M), -float("inf"), device=cuda_device)
    step = 2_000_000  # chunked so the [E, M] message tensor never exists at once
    for lo in range(0, E, step):
        hi = min(E, lo + step)
        m = torch.nn.functional.gelu(u[plan.urow[lo:hi].long()] + v[plan.vrow[lo:hi].long()])
        expect = expect.scatter_reduce(0, t_sorted[lo:hi].view(-1, 1).expand(-1, M), m, "amax", include_self=True)
        del m
    expect = torch.where(torch.isinf(expect), torch.zeros_like(expect), expect)  # nodes without in-edges aggregate to 0
    torch.testing.assert_close(agg, expect, atol=1e-6, rtol=1e-6)
    assert bool((agg[deg == 0] == 0).all()) and bool((ewin[deg == 0] == -1).all())
    # winners lie inside their target's segment and reproduce the stored pre-activation
    has = deg > 0
    e = ewin[has].long()
    lo_ = plan.row_ptr[:-1].long()[has].view(-1, 1); hi_ = plan.row_ptr[1:].long()[has].view(-1, 1)
    assert bool(((e >= lo_) & (e < hi_)).all())
    sample = torch.randint(0, int(has.sum()), (4096,), device=cuda_device, generator=g)
    es = e[sample]
    cols = torch.arange(M, device=cuda_device).view(1, -1).expand_as(es)
    x_re = u[plan.urow.long()[es], cols] + v[plan.vrow.long()[es], cols]
    torch.testing.assert_close(x_re, xwin[has][sample], atol=0, rtol=0)
