"""B200 parity of the seq-attention kernels through the C ABI, against the CPU oracle.

First B200 run: round 2, 11/11 passed (gpurun_out/r2c1_seq_tests.txt); part of the default ``-m gpu`` suite since then
(BUGLAB_B200_SEQ_GPU=0 skips them)."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("BUGLAB_B200_SEQ_GPU", "1") == "0",
                                 reason="seq-attention GPU parity switched off by BUGLAB_B200_SEQ_GPU=0")]


LARGE_CASES = [  # seed, B, L, E, d_model, heads, relation kinds, value biases  (GPU only; fp64 oracle)
    (8, 4, 512, 6000, 512, 8, 6, True),     # BASELINE config 4's layer: 8 heads of 64, 512 tokens
    (9, 3, 300, 2500, 512, 8, 5, False),    # padded to 512 keys
    (10, 5, 200, 1500, 256, 4, 4, True),    # padded to 256 keys
]


def _check_case(case, dev, backend, monkeypatch, double_oracle=False):
    import copy

    import test_seq_attention_emul as emul
    from buglab_b200 import ops

    monkeypatch.setattr(ops, "SEQ_ATTENTION_TC", backend == "tensor-core")
    att, x, mask, edges, edge_types, weights = emul.random_case(*case)
    types = case[6]
    keep = ~mask
    if double_oracle:  # exact evaluation of the same layer (oracle/seq_ref.py) in fp64
        att_ref, x_ref = copy.deepcopy(att).double(), x.detach().double().requires_grad_(True)
        expected = att_ref(x_ref, mask, edges, edge_types)
        (expected * weights.double()).sum().backward()
        ref_grads = {n: p.grad.float() for n, p in att_ref.named_parameters()}
        ref_dx, expected = x_ref.grad.float(), expected.float()
    else:
        expected = att(x, mask, edges, edge_types)
        (expected * weights).sum().backward()
        ref_grads = {n: p.grad.clone() for n, p in att.named_parameters()}
        ref_dx = x.grad.clone()

    B, L, _ = x.shape
    lengths = (~mask).sum(dim=1)
    plan = ops.build_seq_attention_plan(edges.to(dev), edge_types.to(dev), lengths.to(dev), L, types)
    H, dk = att._num_heads, att._key_query_dim
    xg = x.detach().to(dev).requires_grad_(True)
    params = {n: p.detach().to(dev).requires_grad_(True) for n, p in att.named_parameters()}
    per_head = (xg @ params["_selfatt_head_transforms.weight"].t()).view(B, L, H, -1).permute(0, 2, 1, 3)
    q, k, v = per_head[..., :dk] * dk ** -0.5, per_head[..., dk: 2 * dk], per_head[..., 2 * dk:]
    bias = torch.cat((params["_edge_attention_biases.weight"], params["_reverse_edge_attention_biases.weight"])).view(-1, H, dk)
    vbias = None
    if att._use_edge_value_biases:
        vbias = torch.cat((params["_edge_value_biases.weight"], params["_reverse_edge_value_biases.weight"])).view(-1, H, dk)
    assert ops._seq_tc_ok(q) == (backend == "tensor-core")
    out = ops.seq_edge_attention(q, k, v, bias, vbias, plan)
    got = out.permute(0, 2, 1, 3).reshape(B, L, -1) @ params["_out_proj.weight"].t()
    out_tol = 1e-5 if not double_oracle else 1e-5 * max(1.0, float(expected.abs().max()))
    assert float((got.detach().cpu() - expected.detach())[keep].abs().max()) < out_tol
    (got * weights.to(dev)).sum().backward()
    assert float((xg.grad.cpu() - ref_dx).abs().max()) <= 5e-5 * float(ref_dx.abs().max()) + 5e-6
    for name, ref in ref_grads.items():
        err = float((params[name].grad.cpu() - ref).abs().max())
        assert err <= 5e-5 * float(ref.abs().max()) + 5e-6, (name, err)


@pytest.mark.parametrize("backend", ["tensor-core", "cuda-core"])
@pytest.mark.parametrize("case_index", range(7))
def test_kernels_match_oracle(cuda_device, monkeypatch, case_index, backend):
    """``tensor-core``: QK^T / PV / dP / dQ / dK / dV on the TMA-fed tcgen05 GEMMs + the warp-per-row softmax kernels
    (csrc/seq_attention_tc.cu); ``cuda-core``: the fp32 one-thread-per-row kernels (csrc/seq_attention.cu)."""
    import test_seq_attention_emul as emul

    _check_case(emul.CASES[case_index], cuda_device, backend, monkeypatch)


@pytest.mark.parametrize("case", LARGE_CASES, ids=lambda c: f"seed{c[0]}-L{c[2]}")
def test_tensor_core_attention_matches_fp64_oracle_at_full_size(cuda_device, monkeypatch, case):
    """The layer of BASELINE config 4 (d_model 512, 8 heads of 64, up to 512 tokens, thousands of typed-edge entries) on the
    tensor-core path against the fp64 evaluation of the oracle layer: outputs and every gradient."""
    _check_case(case, cuda_device, "tensor-core", monkeypatch, double_oracle=True)


def test_tensor_core_attention_at_config4_shape(cuda_device, monkeypatch):
    """BASELINE config 4's attention shape (8 heads of 64, sequences up to 512 tokens; 6 samples here) with dropout on the
    probabilities and value biases ("rat"): the tensor-core path against the fp32 CUDA-core kernels, which the cases above
    pin to the oracle — outputs and all gradients."""
    from buglab_b200 import ops

    dev = cuda_device
    g = torch.Generator().manual_seed(11)
    B, H, L, D, T = 6, 8, 512, 64, 5
    lengths = torch.tensor([512, 301, 17, 448, 129, 256])
    n_edges = 4000
    eb = torch.randint(0, B, (n_edges,), generator=g)
    es = (torch.rand(n_edges, generator=g) * lengths[eb]).long()
    et = (torch.rand(n_edges, generator=g) * lengths[eb]).long()
    edges = torch.stack((eb, es, et), dim=1)   # [E, 3] = (sample, source position, target position)
    edge_types = torch.randint(0, T, (n_edges,), generator=g)
    plan = ops.build_seq_attention_plan(edges.to(dev), edge_types.to(dev), lengths.to(dev), L, T)

    def run(tc: bool):
        monkeypatch.setattr(ops, "SEQ_ATTENTION_TC", tc)
        gen = torch.Generator().manual_seed(5)
        leaves = [torch.randn(B, H, L, D, generator=gen).mul_(s).to(dev).requires_grad_(True) for s in (0.35, 1.0, 1.0)]
        leaves += [(torch.randn(2 * T, H, D, generator=gen) * 0.3).to(dev).requires_grad_(True) for _ in range(2)]
        q, k, v, bias, vbias = leaves
        fn = ops.SeqEdgeAttentionTcFn if tc else ops.SeqEdgeAttentionFn
        assert ops._seq_tc_ok(q) == tc
        out = fn.apply(q, k, v, bias, vbias, plan, 0.1, 1234)
        w = torch.randn(B, H, L, D, generator=gen).to(dev)
        keep = (torch.arange(L, device=dev).view(1, 1, L, 1) < lengths.to(dev).view(B, 1, 1, 1)).float()
        (out * w * keep).sum().backward()
        return [out.detach() * keep] + [t.grad for t in leaves]

    got, ref = run(True), run(False)
    for name, a, e in zip(("out", "dq", "dk", "dv", "d_bias", "d_vbias"), got, ref):
        scale = float(e.abs().max())
        err = float((a - e).abs().max())
        assert err <= 2e-5 * scale + 1e-6, (name, err, scale)


@pytest.mark.parametrize("layer_type", ["great", "rat", "transformer", "gru"])
def test_module_matches_reference_goldens(cuda_device, layer_type):
    """The whole sequence module on the B200 path against the real reference's loss and gradients
    (tests/golden/seq_model.npz) — the GPU twin of test_seq_golden.test_mirror_module_reproduces_reference_on_cpu_kernels."""
    import logging
    from pathlib import Path

    import numpy as np

    import test_seq_golden as sg
    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_msgpack_l_gz

    golden = np.load(os.path.join(sg.GOLDEN_DIR, "seq_model.npz"), allow_pickle=True)
    samples = lambda: list(load_msgpack_l_gz(os.path.join(sg.GOLDEN_DIR, "seq_samples.msgpack.l.gz")))  # noqa: E731
    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)
    model, _, _ = load_model(dict(sg.SPEC, modelName=f"seq-{layer_type}"), Path("/tmp/_seq_gpu.pkl.gz"))
    model.compute_metadata(iter(samples()))
    nn = model.build_neural_module()
    nn._argswap_module._input_dim = sg.SPEC["hidden_state_size"]
    order = list(golden[f"{layer_type}/edge_types_in_reference_order"])
    perm = torch.tensor([order.index(kind) for kind in model.edge_types])
    inverse = torch.argsort(perm)
    prefix = f"{layer_type}/param/"
    state = {k[len(prefix):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith(prefix)}
    rows = state["_SeqBugLabModule__positional_encoding"].shape[1]
    full = nn.state_dict()["_SeqBugLabModule__positional_encoding"].clone()
    full[:, :rows] = state["_SeqBugLabModule__positional_encoding"]
    state["_SeqBugLabModule__positional_encoding"] = full
    for k in list(state):
        if "edge_attention_biases" in k or "edge_value_biases" in k:
            state[k] = state[k][perm]
    nn.load_state_dict(state)
    nn.to(cuda_device).train()
    packed = model.initialize_minibatch()
    for t in (model.tensorize(dp) for dp in samples()):
        if t is not None:
            model.extend_minibatch_with(t, packed)
    mb = model.finalize_minibatch(packed, cuda_device)
    loss = nn(**mb)
    assert abs(float(loss.detach()) - float(golden[f"{layer_type}/loss"])) < 1e-4      # north_star's fp32 tolerance
    loss.backward()
    for name, p in nn.named_parameters():
        key = f"{layer_type}/grad/{name}"
        if key not in golden.files:
            continue
        ref = torch.from_numpy(golden[key])
        grad = p.grad.cpu()
        if "positional_encoding" in name:
            grad = grad[:, :rows]
        if "edge_attention_biases" in name or "edge_value_biases" in name:
            grad = grad[inverse]
        assert float((grad - ref).abs().max()) <= 1e-4 * (float(ref.abs().max()) + 1e-12) + 1e-5, name


def test_train_and_evaluate_entry_points_run_for_seq_great(cuda_device, tmp_path):
    """python -m buglab.models.train seq-great TRAIN VALID MODEL.pkl.gz on synthetic program shards (host-language decode:
    the native tensoriser serves the graph models), then evaluate: the trainer, the sequence model, the tensor-core
    attention and the TMA-GEMM dense layers together (hidden 128 / 2 heads of 64: every product takes the tcgen05 path)."""
    import logging

    from buglab.models import evaluate, train
    from buglab_b200 import ops
    from buglab_b200.synthetic import write_shards

    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)
    write_shards(str(tmp_path / "train"), 2, 10, seed=1, programs=True, statements=10)
    write_shards(str(tmp_path / "valid"), 1, 6, seed=2, programs=True, statements=10)
    model_path = tmp_path / "model.pkl.gz"
    calls = {"attention": 0}
    original = ops.SeqEdgeAttentionTcFn.forward

    def counting(ctx, *args, **kwargs):
        calls["attention"] += 1
        return original(ctx, *args, **kwargs)

    ops.SeqEdgeAttentionTcFn.forward = staticmethod(counting)
    try:
        train.main(["seq-great", str(tmp_path / "train"), str(tmp_path / "valid"), str(model_path), "--max-num-epochs=2",
                    "--minibatch-size=5", "--quiet", "--model-spec",
                    '{"hidden_state_size": 128, "num_heads": 2, "num_layers": 2, "max_seq_size": 256, "intermediate_dimension_size": 256}'])
    finally:
        ops.SeqEdgeAttentionTcFn.forward = staticmethod(original)
    assert model_path.exists() and calls["attention"] > 0
    args = {"MODEL_FILENAME": str(model_path), "TEST_DATA_PATH": str(tmp_path / "valid"), "--limit-num-elements": None,
            "--sequential": True, "--azure-info": None}
    metrics = evaluate.run(args)
    assert metrics["num_samples"] > 0 and 0.0 <= metrics["localization_accuracy"] <= 1.0
