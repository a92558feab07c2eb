"""The seq-attention CUDA kernels without a GPU: csrc/seq_attention_core.h (the per-row arithmetic the kernels execute, one
thread per row) is compiled for the host and driven by loops (tests/emul/seq_attention_emul.cpp), then compared with the
oracle (oracle/seq_ref.py, itself pinned to the real reference).  Covers the arithmetic, the entry (CSR) construction in
``ops.build_seq_attention_plan`` and the autograd wiring of ``ops.SeqEdgeAttentionFn`` — everything except the CUDA launch."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

from buglab_b200 import ops  # noqa: E402
from oracle import seq_ref  # noqa: E402


def attention_through_the_kernels(att, x, mask, edges, edge_types, num_edge_types):
    """RelationalMultiheadAttention.forward with the score/softmax/value part done by the kernel source."""
    B, L, _ = x.shape
    lengths = (~mask).sum(dim=1) if mask is not None else torch.full((B,), L)
    plan = ops.build_seq_attention_plan(edges, edge_types, lengths, L, num_edge_types)
    q, k, v = att.project(x)
    H, dk = att._num_heads, att._key_query_dim
    bias = torch.cat((att._edge_attention_biases.weight, att._reverse_edge_attention_biases.weight)).view(-1, H, dk)
    vbias = None
    if att._use_edge_value_biases:
        vbias = torch.cat((att._edge_value_biases.weight, att._reverse_edge_value_biases.weight)).view(-1, H, att._value_dim)
    return att.merge(ops.seq_edge_attention(q, k, v, bias, vbias, plan))


def random_case(seed, B, L, E, d_model, heads, types, value_biases):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    att = seq_ref.RelationalMultiheadAttention(num_heads=heads, num_edge_types=types, input_state_dimension=d_model,
                                               key_query_dimension=d_model // heads, value_dimension=d_model // heads,
                                               output_dimension=d_model, dropout_rate=0.0,
                                               use_edge_value_biases=value_biases).eval()
    lengths = torch.randint(max(1, L // 2), L + 1, (B,), generator=g)
    lengths[0] = L
    mask = torch.arange(L)[None, :] >= lengths[:, None]
    x = (torch.randn(B, L, d_model, generator=g) * (~mask)[..., None]).requires_grad_(True)
    sample = torch.randint(0, B, (E,), generator=g)
    hi = lengths[sample].float()
    edges = torch.stack([sample, (torch.rand(E, generator=g) * hi).long(), (torch.rand(E, generator=g) * hi).long()], dim=1)
    edge_types = torch.randint(0, types, (E,), generator=g)
    weights = torch.randn(B, L, d_model, generator=g) * (~mask)[..., None]
    return att, x, mask, edges, edge_types, weights


CASES = [  # seed, B, L, E, d_model, heads, relation kinds, value biases
    (1, 3, 12, 40, 32, 4, 5, False),      # seq-great shape (head dim 8)
    (2, 3, 12, 40, 32, 4, 5, True),       # seq-rat
    (3, 1, 6, 90, 16, 2, 3, True),        # dense duplicates: many entries on the same (i, j)
    (4, 2, 9, 0, 32, 2, 4, True),         # no edges at all (head dim 16)
    (5, 2, 33, 150, 64, 2, 6, True),      # head dim 32, length not a multiple of anything
    (6, 2, 20, 60, 128, 2, 4, False),     # head dim 64 (config 4's)
    (7, 4, 7, 25, 32, 4, 2, True),        # short sequences, heavy padding
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"seed{c[0]}")
def test_kernel_source_matches_oracle(host_backend, case):
    seed, B, L, E, d_model, heads, types, value_biases = case
    att, x, mask, edges, edge_types, weights = random_case(*case)
    keep = ~mask

    expected = att(x, mask, edges, edge_types)
    (expected * weights).sum().backward()
    ref_grads = {n: p.grad.clone() for n, p in att.named_parameters()}
    ref_dx = x.grad.clone()
    x.grad = None
    att.zero_grad()

    got = attention_through_the_kernels(att, x, mask, edges, edge_types, types)
    assert float((got - expected).detach()[keep].abs().max()) < 3e-6
    (got * weights).sum().backward()
    scale = float(ref_dx.abs().max()) + 1e-12
    assert float((x.grad - ref_dx).abs().max()) <= 2e-5 * scale + 2e-6
    for name, p in att.named_parameters():
        ref = ref_grads[name]
        tol = 2e-5 * (float(ref.abs().max()) + 1e-12) + 2e-6
        assert float((p.grad - ref).abs().max()) <= tol, (name, float((p.grad - ref).abs().max()), tol)


def test_plan_lists_both_directions_sorted():
    edges = torch.tensor([[0, 2, 1], [0, 2, 1], [1, 0, 3], [0, 1, 1]])
    types = torch.tensor([1, 0, 2, 1])
    plan = ops.build_seq_attention_plan(edges, types, torch.tensor([3, 4]), 4, 3)
    assert plan.num_tables == 6 and plan.row_key.shape[0] == 8
    rp, rk, rt = plan.row_ptr.tolist(), plan.row_key.tolist(), plan.row_tab.tolist()
    rows = {r: sorted(zip(rk[rp[r]: rp[r + 1]], rt[rp[r]: rp[r + 1]])) for r in range(8) if rp[r + 1] > rp[r]}
    # sample 0: (2->1, kind 1), (2->1, kind 0), (1->1 self loop, kind 1); sample 1: (0->3, kind 2); reverse tables are 3 + kind
    assert rows == {1: [(1, 1), (1, 4), (2, 3), (2, 4)], 2: [(1, 0), (1, 1)], 4: [(3, 2)], 7: [(0, 5)]}
    for r in range(8):
        assert rk[rp[r]: rp[r + 1]] == sorted(rk[rp[r]: rp[r + 1]])          # keys ascending within a row
    cp, cq, ct = plan.col_ptr.tolist(), plan.col_query.tolist(), plan.col_tab.tolist()
    cols = {c: sorted(zip(cq[cp[c]: cp[c + 1]], ct[cp[c]: cp[c + 1]])) for c in range(8) if cp[c + 1] > cp[c]}
    assert cols == {1: [(1, 1), (1, 4), (2, 0), (2, 1)], 2: [(1, 3), (1, 4)], 4: [(3, 5)], 7: [(0, 2)]}
    for c in range(8):
        assert cq[cp[c]: cp[c + 1]] == sorted(cq[cp[c]: cp[c + 1]])
    empty = ops.build_seq_attention_plan(torch.zeros((0, 3), dtype=torch.long), torch.zeros(0, dtype=torch.long),
                                         torch.tensor([2]), 3, 2)
    assert empty.row_ptr.tolist() == [0, 0, 0, 0] and empty.row_key.numel() == 0


def test_golden_layer_cases_through_the_kernel_source(host_backend):
    """The reference-generated layer fixtures (tests/golden/seq_layers.npz), attention computed by the kernel source."""
    import numpy as np

    import test_seq_oracle as layer_tests

    golden = np.load(layer_tests.GOLDEN)
    for name in ("great_postnorm", "rat_postnorm", "prenorm_gelu", "postnorm_rezero_vector", "medium"):
        layer, src, mask, edges, edge_types = layer_tests.build(golden, name)
        types = int(golden[f"{name}/dims"][5])
        att = layer.self_attn
        att.forward = lambda x, m, e, t, att=att, types=types: attention_through_the_kernels(att, x, m, e, t, types)
        y = layer(src, mask, edges, edge_types)
        keep = ~mask
        expected = torch.from_numpy(golden[f"{name}/out"])
        assert float((y.detach() - expected)[keep].abs().max()) < 5e-6, name
        (y * torch.from_numpy(golden[f"{name}/weights"])).sum().backward()
        g = torch.from_numpy(golden[f"{name}/grad_src"])
        assert float((src.grad - g).abs().max()) <= 2e-5 * float(g.abs().max()) + 2e-6, name
        for pname, p in layer.named_parameters():
            key = f"{name}/grad/{pname}"
            if key in golden.files:
                ref = torch.from_numpy(golden[key])
                assert float((p.grad - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 2e-6, (name, pname)


def dropout_scale_table(seed: int, B: int, H: int, L: int, p_drop: float) -> torch.Tensor:
    """The kernels' counter-based mask (csrc/seq_attention_core.h::dropout_scale) restated with numpy uint64 arithmetic."""
    import numpy as np

    idx = np.arange(B * H * L * L, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15) + np.uint64(0x632BE59BD9B4E019)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    u = ((z >> np.uint64(32)).astype(np.uint32) >> np.uint32(8)).astype(np.float32) * np.float32(1.0 / 16777216.0)
    scale = np.where(u >= np.float32(p_drop), np.float32(1.0) / (np.float32(1.0) - np.float32(p_drop)), np.float32(0.0))
    return torch.from_numpy(scale.reshape(B, H, L, L).astype(np.float32))


@pytest.mark.parametrize("value_biases", [False, True])
def test_attention_probability_dropout(host_backend, value_biases):
    """Dropout on the probabilities (multihead_attention.py:77), fused into the kernels with a counter-based mask: same
    result as masking the oracle's probabilities with that mask, forward and backward; about p_drop of the mass is dropped."""
    seed, p_drop = 123456789012345, 0.3
    att, x, mask, edges, edge_types, weights = random_case(11, 2, 10, 30, 32, 4, 3, value_biases)
    B, L, _ = x.shape
    H, dk = att._num_heads, att._key_query_dim
    keep = ~mask
    lengths = keep.sum(dim=1)
    plan = ops.build_seq_attention_plan(edges, edge_types, lengths, L, 3)
    scale = dropout_scale_table(seed, B, H, L, p_drop)
    assert 0.2 < float((scale == 0).float().mean()) < 0.4

    def tables():
        bias = torch.cat((att._edge_attention_biases.weight, att._reverse_edge_attention_biases.weight)).view(-1, H, dk)
        vbias = (torch.cat((att._edge_value_biases.weight, att._reverse_edge_value_biases.weight)).view(-1, H, dk)
                 if value_biases else None)
        return bias, vbias

    # oracle arithmetic with the explicit mask
    q, k, v = att.project(x)
    sample, src, tgt = edges[:, 0], edges[:, 1], edges[:, 2]
    fwd, rev = att.edge_score_terms(sample, src, tgt, edge_types, q, k)
    scores = seq_ref._accumulate_scores(q @ k.transpose(-1, -2), sample, src, tgt, fwd, rev)
    probs = torch.softmax(scores.masked_fill(mask[:, None, None, :], float("-inf")), dim=-1) * scale
    out = probs @ v
    if value_biases:
        _, vbias = tables()
        T = 3
        out = seq_ref._accumulate_rows(out, sample, src, tgt, probs[sample, :, src, tgt].unsqueeze(-1) * vbias[edge_types],
                                       probs[sample, :, tgt, src].unsqueeze(-1) * vbias[T + edge_types])
    expected = att.merge(out)
    (expected * weights).sum().backward()
    ref_dx = x.grad.clone()
    ref_grads = {n: p.grad.clone() for n, p in att.named_parameters()}
    x.grad = None
    att.zero_grad()

    q, k, v = att.project(x)
    bias, vbias = tables()
    got = att.merge(ops.SeqEdgeAttentionFn.apply(q, k, v, bias, vbias, plan, p_drop, seed))
    assert float((got - expected).detach()[keep].abs().max()) < 3e-6
    (got * weights).sum().backward()
    assert float((x.grad - ref_dx).abs().max()) <= 2e-5 * float(ref_dx.abs().max()) + 2e-6
    for name, p in att.named_parameters():
        ref = ref_grads[name]
        assert float((p.grad - ref).abs().max()) <= 2e-5 * float(ref.abs().max()) + 2e-6, name

    # eval mode / p = 0 is the undropped attention; training mode draws a fresh seed per call
    plain = ops.seq_edge_attention(q.detach(), k.detach(), v.detach(), bias.detach(), None if vbias is None else vbias.detach(),
                                   plan, p_drop, training=False)
    again = ops.seq_edge_attention(q.detach(), k.detach(), v.detach(), bias.detach(), None if vbias is None else vbias.detach(),
                                   plan, 0.0, training=True)
    assert torch.equal(plain, again)
    torch.manual_seed(1)
    a = ops.seq_edge_attention(q.detach(), k.detach(), v.detach(), bias.detach(), None, plan, p_drop, training=True)
    b = ops.seq_edge_attention(q.detach(), k.detach(), v.detach(), bias.detach(), None, plan, p_drop, training=True)
    assert not torch.equal(a, b)


def test_plan_against_brute_force_on_random_edge_sets():
    g = torch.Generator().manual_seed(5)
    for trial in range(20):
        B, L, T = int(torch.randint(1, 5, (1,), generator=g)), int(torch.randint(2, 40, (1,), generator=g)), int(torch.randint(1, 6, (1,), generator=g))
        E = int(torch.randint(0, 300, (1,), generator=g))
        edges = torch.stack([torch.randint(0, B, (E,), generator=g), torch.randint(0, L, (E,), generator=g),
                             torch.randint(0, L, (E,), generator=g)], dim=1)
        kinds = torch.randint(0, T, (E,), generator=g)
        plan = ops.build_seq_attention_plan(edges, kinds, torch.full((B,), L), L, T)
        expected_rows = {}
        for (b, s, t), kind in zip(edges.tolist(), kinds.tolist()):
            expected_rows.setdefault(b * L + s, []).append((t, kind))
            expected_rows.setdefault(b * L + t, []).append((s, T + kind))
        rp, rk, rt = plan.row_ptr.tolist(), plan.row_key.tolist(), plan.row_tab.tolist()
        cp, cq, ct = plan.col_ptr.tolist(), plan.col_query.tolist(), plan.col_tab.tolist()
        assert rp[-1] == cp[-1] == 2 * E
        expected_cols = {}
        for row, entries in expected_rows.items():
            b, i = divmod(row, L)
            for key, tab in entries:
                expected_cols.setdefault(b * L + key, []).append((i, tab))
        for r in range(B * L):
            got = list(zip(rk[rp[r]: rp[r + 1]], rt[rp[r]: rp[r + 1]]))
            assert sorted(got) == sorted(expected_rows.get(r, [])) and [k for k, _ in got] == sorted(k for k, _ in got)
            got = list(zip(cq[cp[r]: cp[r + 1]], ct[cp[r]: cp[r + 1]]))
            assert sorted(got) == sorted(expected_cols.get(r, [])) and [q for q, _ in got] == sorted(q for q, _ in got)
