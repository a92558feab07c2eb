"""Known-answer tests that pin the UNPINNED part of the oracle (ptgnn / torch_scatter semantics):
hand-computed tiny cases, per-group log_softmax identities, and the GELU property the fused kernel relies on."""
import math

import torch

from oracle import scatter_ref
from oracle.mp_ref import MlpMessagePassingLayer, edge_messages_ref, typed_edge_message_max_ref


def test_scatter_max_first_wins_and_empty_is_zero():
    src = torch.tensor([1.0, 3.0, 3.0, -1.0, 5.0, -2.0])
    idx = torch.tensor([0, 1, 1, 0, 3, 4])
    out, arg = scatter_ref.scatter_max(src, idx, dim_size=6)
    assert out.tolist() == [1.0, 3.0, 0.0, 5.0, -2.0, 0.0]      # segments 2 and 5 are empty -> 0
    assert arg.tolist() == [0, 1, 6, 4, 5, 6]                    # tie in segment 1 -> first index; empty -> len(src)
    out, arg = scatter_ref.scatter_min(src, idx, dim_size=5)
    assert out.tolist() == [-1.0, 3.0, 0.0, 5.0, -2.0] and arg.tolist() == [3, 1, 6, 4, 5]


def test_scatter_max_backward_routes_to_arg_only():
    src = torch.tensor([[1.0, 5.0], [2.0, 5.0], [0.5, 7.0]], requires_grad=True)
    out, arg = scatter_ref.scatter_max(src, torch.tensor([0, 0, 1]), dim=0, dim_size=2)
    (out * torch.tensor([[1.0, 10.0], [100.0, 1000.0]])).sum().backward()
    assert arg.tolist() == [[1, 0], [2, 2]]
    assert src.grad.tolist() == [[0.0, 10.0], [1.0, 0.0], [100.0, 1000.0]]


def test_scatter_log_softmax_is_log_softmax_per_group():
    g = torch.Generator().manual_seed(0)
    src = torch.randn(40, generator=g) * 4
    idx = torch.randint(0, 6, (40,), generator=g)
    out = scatter_ref.scatter_log_softmax(src, idx)
    for s in idx.unique():
        m = idx == s
        torch.testing.assert_close(out[m], torch.log_softmax(src[m], 0), atol=1e-6, rtol=1e-6)


def test_no_bug_slot_identity():
    """with all-equal candidate scores s and the virtual NO_BUG logit 1.0, the NO_BUG log-prob is
    1 - logsumexp([s]*C + [1])  (localizationmodule.py:66-77)."""
    C, s = 7, 0.3
    scores = torch.cat((torch.full((C,), s), torch.ones(1)))
    out = scatter_ref.scatter_log_softmax(scores, torch.zeros(C + 1, dtype=torch.int64))
    expected = 1.0 - math.log(C * math.exp(s) + math.exp(1.0))
    assert abs(float(out[-1]) - expected) < 1e-6


def test_hand_computed_typed_edge_layer():
    """4 nodes, 2 edge types, D=1, M=1 — every number below is computed by hand.
       type 0 (W=[1, 0], b=0):   edges 0->2, 1->2   messages GELU(h_s)   = GELU(1), GELU(-3)
       type 1 (W=[0, 2], b=1):   edge  3->2, 0->1   messages GELU(2*h_t+1) = GELU(2*0.5+1)=GELU(2) at node 2, GELU(-5) at node 1
       node 0 and node 3 have no in-edge -> 0."""
    h = torch.tensor([[1.0], [-3.0], [0.5], [4.0]])
    adj = [(torch.tensor([0, 1]), torch.tensor([2, 2])), (torch.tensor([3, 0]), torch.tensor([2, 1]))]
    W = torch.tensor([[[1.0, 0.0]], [[0.0, 2.0]]])
    b = torch.tensor([[0.0], [1.0]])
    gelu = lambda x: 0.5 * x * (1 + math.erf(x / math.sqrt(2)))  # noqa: E731
    msgs, tgt = edge_messages_ref(h, adj, W, b)
    assert tgt.tolist() == [2, 2, 2, 1]
    torch.testing.assert_close(msgs.view(-1), torch.tensor([gelu(1.0), gelu(-3.0), gelu(2.0), gelu(-5.0)]), atol=1e-6, rtol=1e-6)
    agg, arg = typed_edge_message_max_ref(h, adj, W, b)
    torch.testing.assert_close(agg.view(-1), torch.tensor([0.0, gelu(-5.0), gelu(2.0), 0.0]), atol=1e-6, rtol=1e-6)
    assert arg.view(-1).tolist() == [4, 3, 2, 4]  # edge ids in the type-major concatenation; 4 == "no edge"


def test_gelu_is_quasi_convex_so_max_sits_at_an_extreme():
    """The fused kernel keeps only min(x) and max(x) per segment: max_i GELU(x_i) == max(GELU(min x), GELU(max x))."""
    g = torch.Generator().manual_seed(1)
    for scale in (0.3, 1.0, 3.0, 8.0):
        x = torch.randn(2000, 16, generator=g, dtype=torch.float64) * scale - 0.5
        direct = torch.nn.functional.gelu(x).max(dim=0)[0]
        via_extremes = torch.maximum(torch.nn.functional.gelu(x.min(dim=0)[0]), torch.nn.functional.gelu(x.max(dim=0)[0]))
        torch.testing.assert_close(direct, via_extremes, atol=0, rtol=0)
    # strictly decreasing left of x0 ~ -0.7518 and increasing right of it; checked on [-5, 12] — further left erf
    # saturates (GELU rounds to -0.0 with ulp noise), where any choice among the saturated messages has value 0 and
    # derivative < 1e-6, so the selection is immaterial
    grid = torch.linspace(-5, 12, 170001, dtype=torch.float64)
    y = torch.nn.functional.gelu(grid)
    k = int(y.argmin())
    assert abs(float(grid[k]) + 0.7518) < 1e-3
    assert bool((y[:k].diff() < 0).all()) and bool((y[k:].diff() > 0).all())


def test_layer_module_matches_functional_and_isolated_nodes():
    torch.manual_seed(0)
    layer = MlpMessagePassingLayer(6, 10, 4, num_edge_types=2, message_aggregation_function="max")
    h = torch.randn(5, 6)
    adj = [(torch.tensor([0, 1, 1]), torch.tensor([1, 2, 2])), (torch.zeros(0, dtype=torch.int64), torch.zeros(0, dtype=torch.int64))]
    agg = layer.aggregated_messages(h, adj)
    assert torch.all(agg[[0, 3, 4]] == 0)
    out = layer(h, adj)
    sd = layer.state_dict()
    ln_w, ln_b = sd["_MlpMessagePassingLayer__state_update.0.weight"], sd["_MlpMessagePassingLayer__state_update.0.bias"]
    dense = sd["_MlpMessagePassingLayer__state_update.1.weight"]
    expect = torch.tanh(torch.nn.functional.layer_norm(agg, (10,), ln_w, ln_b) @ dense.t())
    torch.testing.assert_close(out, expect)


def test_routing_audit_accepts_the_exact_argmax_and_rejects_wrong_edges():
    """The audit used by the GPU whole-model tests (oracle/parity.py::assert_routing_is_valid): forcing the layer's own
    first-max-wins routing audits clean; a winner from another node's segment, a non-maximal edge of the right segment
    and a wrong claim about an empty segment are each caught."""
    import pytest
    import torch

    from oracle import parity
    from oracle.mp_ref import MlpMessagePassingLayer, edge_messages_ref
    from oracle.scatter_ref import scatter_max

    torch.manual_seed(0)
    N, D, K = 7, 8, 3
    layer = MlpMessagePassingLayer(D, D, D, K).double()
    h = torch.randn(N, D, dtype=torch.float64)
    adjacency = [(torch.randint(0, N - 1, (9,)), torch.randint(0, N - 1, (9,))) for _ in range(K)]  # node N-1: no in-edges
    layers = layer._MlpMessagePassingLayer__edge_message_transformation_layers
    weight, bias = torch.stack([l.weight for l in layers]), torch.stack([l.bias for l in layers])
    messages, targets = edge_messages_ref(h, adjacency, weight, bias)
    E = messages.shape[0]
    own_max, own_arg = scatter_max(messages, targets, dim=0, dim_size=N)
    assert (own_arg[N - 1] == E).all()

    def audit(arg):
        layer.forced_winners = arg
        out = layer.aggregated_messages(h, adjacency)
        return out, [layer.routing_audit]

    out, audits = audit(own_arg)
    assert torch.equal(out, own_max)
    worst = parity.assert_routing_is_valid(audits, "own routing")
    assert worst["differing_frac"] == 0.0 and worst["max_relative_deficit"] == 0.0

    node = int(targets[0])
    other = next(e for e in range(E) if int(targets[e]) != node)
    bad = own_arg.clone(); bad[node, 0] = other
    with pytest.raises(AssertionError, match="not in-edges"):
        parity.assert_routing_is_valid(audit(bad)[1], "foreign edge")

    seg = [e for e in range(E) if int(targets[e]) == node]
    if len(seg) > 1:
        loser = next(e for e in seg if e != int(own_arg[node, 1]))
        bad = own_arg.clone(); bad[node, 1] = loser
        with pytest.raises(AssertionError, match="falls short"):
            parity.assert_routing_is_valid(audit(bad)[1], "non-maximal edge")

    bad = own_arg.clone(); bad[N - 1, 0] = 0
    with pytest.raises(AssertionError):
        parity.assert_routing_is_valid(audit(bad)[1], "edge claimed for an empty segment")


def test_localization_summary_forced_routing_and_audit():
    """The oracle's localisation module under forced candidate-summary routing: its own argmax reproduces the unforced
    log-probabilities and gradients and audits clean; a winner from another sample is flagged; a near-tie swap is accepted
    with a tiny deficit (oracle/model_ref.py::LocalizationModule._summary)."""
    import pytest

    from oracle import parity
    from oracle.model_ref import LocalizationModule

    torch.manual_seed(4)
    dim, C, S = 8, 11, 3
    mod = LocalizationModule(dim).double()
    reprs = torch.randn(C, dim, dtype=torch.float64, requires_grad=True)
    to_sample = torch.tensor([0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 2])
    _, lp, _ = mod.compute_localization_logprobs(reprs, to_sample, S)
    lp.sum().backward()
    grad_own = reprs.grad.clone()
    own_arg = scatter_ref.scatter_max(mod._summary_repr(reprs), to_sample, dim=0)[1]

    mod.forced_summary_args = own_arg
    reprs.grad = None
    _, lp_forced, _ = mod.compute_localization_logprobs(reprs, to_sample, S)
    lp_forced.sum().backward()
    assert torch.equal(lp_forced, lp) and torch.equal(reprs.grad, grad_own)
    assert parity.assert_routing_is_valid([mod.routing_audit], "own argmax")["differing_frac"] == 0.0

    wrong = own_arg.clone()
    wrong[0, 0] = 5  # a candidate of sample 1 routed into sample 0
    mod.forced_summary_args = wrong
    mod.compute_localization_logprobs(reprs, to_sample, S)
    with pytest.raises(AssertionError, match="not in-edges"):
        parity.assert_routing_is_valid([mod.routing_audit], "wrong sample")


def test_forced_relu_kinks_and_their_audit():
    """oracle/model_ref.py::AuditedReLU + parity.relu_trace: a module evaluated under its own traced kink decisions
    reproduces itself exactly and audits clean; a forced decision on the wrong side of a pre-activation far from zero is
    reported with that distance."""
    import pytest

    from oracle import parity
    from oracle.model_ref import MLP, AuditedReLU

    torch.manual_seed(9)
    mlp = MLP(6, 1, [5]).double()
    x = torch.randn(7, 6, dtype=torch.float64, requires_grad=True)
    with parity.relu_trace(mlp) as masks:
        y = mlp(x)
    assert list(masks) == ["_layers.0"] and masks["_layers.0"][0].shape == (7, 5)
    y.sum().backward()
    grad_own = x.grad.clone()
    relu = next(m for m in mlp.modules() if isinstance(m, AuditedReLU))
    relu.forced_masks, relu._calls = masks["_layers.0"], 0
    x.grad = None
    y_forced = mlp(x)
    y_forced.sum().backward()
    assert torch.equal(y_forced, y) and torch.equal(x.grad, grad_own)
    assert relu.routing_audit["differing"] == 0 and relu.routing_audit["decisions"] == 35
    parity.assert_routing_is_valid([relu.routing_audit], "own kinks")

    pre = mlp._layers[0](x).detach()
    far = int(pre.abs().argmax())
    wrong = masks["_layers.0"][0].clone()
    wrong.view(-1)[far] = ~wrong.view(-1)[far]
    relu.forced_masks, relu._calls, relu.routing_audit = [wrong], 0, None
    mlp(x)
    assert relu.routing_audit["differing"] == 1
    assert abs(relu.routing_audit["max_relative_deficit"] - float(pre.abs().max())) < 1e-12
    with pytest.raises(AssertionError, match="falls short"):
        parity.assert_routing_is_valid([relu.routing_audit], "wrong kink")


def test_scatter_ref_values_agree_with_torch_scatter_reduce():
    """An independent witness for the restated torch_scatter reductions (the package itself is absent):
    PyTorch's own ``scatter_reduce`` (amax / amin / sum / mean, ``include_self=False``) must give the same VALUES on
    non-empty segments; torch_scatter's conventions that PyTorch does not share — empty segments -> 0 and the arg output with
    first-extreme-wins — stay pinned by the hand-computed cases above."""
    g = torch.Generator().manual_seed(3)
    src = torch.randn(200, 7, generator=g, dtype=torch.float64)
    index = torch.randint(0, 23, (200,), generator=g)
    index[index == 5] = 6   # segment 5 is empty
    S = 25                  # 23, 24 empty too
    idx2 = index.view(-1, 1).expand_as(src)
    nonempty = torch.zeros(S, dtype=torch.bool).index_fill_(0, index, True)
    for name, ref_fn in (("amax", scatter_ref.scatter_max), ("amin", scatter_ref.scatter_min)):
        ours = ref_fn(src, index, dim=0, dim_size=S)[0]
        theirs = torch.zeros(S, 7, dtype=torch.float64).scatter_reduce(0, idx2, src, reduce=name, include_self=False)
        assert torch.equal(ours[nonempty], theirs[nonempty])
        assert torch.all(ours[~nonempty] == 0)
    assert torch.allclose(scatter_ref.scatter_sum(src, index, dim=0, dim_size=S),
                          torch.zeros(S, 7, dtype=torch.float64).scatter_reduce(0, idx2, src, reduce="sum", include_self=False))
    mean_theirs = torch.zeros(S, 7, dtype=torch.float64).scatter_reduce(0, idx2, src, reduce="mean", include_self=False)
    assert torch.allclose(scatter_ref.scatter_mean(src, index, dim=0, dim_size=S)[nonempty], mean_theirs[nonempty])
