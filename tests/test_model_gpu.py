"""GPU parity of the whole gnn-mlp step against the CPU oracle, through the reference-facing operator surface
(ptgnn classes + buglab model mirror).  fp32 tolerance 1e-4 (BASELINE.json north_star); indices bit-exact."""
import copy
import os
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = dict(atol=1e-4, rtol=1e-4)


def _setup(hidden, device, n_graphs=10, mean_nodes=250, seed=0, dropout=0.0, make_ref=True):
    from buglab.models.modelregistry import load_model
    from buglab_b200.synthetic import SyntheticBugLabGenerator
    from oracle import model_ref

    torch.manual_seed(seed)
    gen = SyntheticBugLabGenerator(seed=seed, mean_nodes=mean_nodes, min_nodes=40)
    data = [gen.sample() for _ in range(n_graphs)]
    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": dropout},
                             Path("/tmp/buglab_b200_test.pkl.gz"))
    model.gnn_model.node_representation_model.dropout_rate = dropout
    model.compute_metadata(iter(copy.deepcopy(data)))
    nn = model.build_neural_module().to(device)
    ref = None
    if make_ref:
        ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types,
                                        len(model.gnn_model.node_representation_model.vocabulary),
                                        len(model._target_rewrite_ops))
        ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
    tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
    return model, nn, ref, data, tensors


def _check_steps(model, nn, ref, tensors, device, minibatch_size, what, max_batches=2):
    """Forward + gradient parity of whole train steps against the CPU oracle (fp32 and fp64), with the effective
    tolerances and the routing audit printed (run pytest with -s to see them)."""
    from buglab_b200 import ops
    from oracle import model_ref, parity

    report = []
    for b, (mb, _raw) in enumerate(model.minibatch_iterator(iter(tensors), device, minibatch_size, parallelize=False)):
        if b >= max_batches:
            break
        nn.zero_grad(); ref.zero_grad()
        nn.train()
        ops.WINNER_TRACE, ops.MINMAX_TRACE = [], []
        with parity.relu_trace(nn) as relu_masks:   # the heads' ReLU kink decisions (oracle/model_ref.py::AuditedReLU)
            loss = nn(**mb)
        winners, ops.WINNER_TRACE = ops.WINNER_TRACE, None
        head_args, ops.MINMAX_TRACE = ops.MINMAX_TRACE, None   # the localisation module's max over the candidates
        loss.backward()
        mb_cpu = model_ref.minibatch_to_cpu(mb)
        ref.force_routing(None, None)
        loss_ref, det = ref(**mb_cpu, return_details=True)

        ref64 = copy.deepcopy(ref).double()  # fp64 referee: exact evaluation of the same semantics (oracle/parity.py)
        loss64, det64 = ref64(**mb_cpu, return_details=True)
        parity.assert_forward_close_deep(loss, loss_ref, loss64, f"{what}: loss")
        groups, lp, gnn_out, _ = nn.compute_localization_logprobs(mb["graph_data"])
        states = gnn_out.output_node_representations
        parity.assert_forward_close_deep(states, det["node_states"], det64["node_states"], f"{what}: node states")
        parity.assert_forward_close_deep(lp, det["localization_logprobs"], det64["localization_logprobs"],
                                         f"{what}: localization log-probs")
        assert torch.equal(groups.cpu(), det["localization_groups"])
        slack = float((det["node_states"].double() - det64["node_states"]).abs().max())
        dist64 = float((states.detach().cpu().double() - det64["node_states"]).abs().max())
        dist32 = float((states.detach().cpu() - det["node_states"]).abs().max())

        # ROUTING, verified independently of the conditioning below: the exact (fp64) oracle audits every winner the
        # GPU path chose against its own segment maxima (oracle/parity.py::assert_routing_is_valid)
        ref64.force_routing(winners, head_args, relu_masks)
        ref64(**mb_cpu)
        routing = parity.assert_routing_is_valid(ref64.routing_audits(), what)
        # GRADIENTS: oracle re-run with the GPU path's max-routing forced (message-passing layers AND the localisation
        # module's candidate summary) -> elementwise comparable
        ref.force_routing(winners, head_args, relu_masks)
        ref.zero_grad()
        ref(**mb_cpu).backward()
        ref.force_routing(None, None)
        ref_params = dict(ref.named_parameters())
        worst_frac = worst_l2 = 0.0
        for name, p in nn.named_parameters():
            g_ref = ref_params[name].grad
            if g_ref is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
                continue
            parity.assert_grad_close(p.grad, g_ref, f"{what}: {name}")
            frac, l2, _ = parity.grad_mismatch(p.grad, g_ref)
            worst_frac, worst_l2 = max(worst_frac, frac), max(worst_l2, l2)
        n_nodes = int(states.shape[0])
        report.append(f"{what} batch {b}: {n_nodes} nodes; node states: |gpu-fp64|max {dist64:.2e}, |gpu-fp32 oracle|max "
                      f"{dist32:.2e}, oracle fp32-vs-fp64 noise floor (slack) {slack:.2e} -> effective tolerance "
                      f"{1e-4 + slack:.2e}; winners differing from the exact argmax {routing['differing_frac']:.2e}, "
                      f"worst relative deficit {routing['max_relative_deficit']:.2e}; gradients (forced routing): worst "
                      f"share of entries beyond 1e-4 {worst_frac:.2e}, worst relative L2 {worst_l2:.2e}")
    assert report
    print("\n" + "\n".join(report))


@pytest.mark.parametrize("hidden", [32, 128])
def test_step_matches_oracle(cuda_device, hidden):
    model, nn, ref, data, tensors = _setup(hidden, cuda_device)
    _check_steps(model, nn, ref, tensors, cuda_device, 5, f"H={hidden}")


def test_step_matches_oracle_at_bench_width(cuda_device):
    """BASELINE configs[1] shape in small: H = 256 (the 512-wide post-residual layers and every <256>/<512> kernel
    instantiation the bench uses), 8 graphs of ~2 000 nodes in one minibatch."""
    model, nn, ref, data, tensors = _setup(256, cuda_device, n_graphs=8, mean_nodes=2000, seed=3)
    _check_steps(model, nn, ref, tensors, cuda_device, 8, "c2-shaped H=256", max_batches=1)


def test_step_matches_oracle_config1(cuda_device):
    """BASELINE configs[0]: H = 128, ~2 000-node graphs, minibatches cut by the reference's 30 000-node budget
    (modelregistry.py:53-54) rather than by the graph count."""
    model, nn, ref, data, tensors = _setup(128, cuda_device, n_graphs=24, mean_nodes=2000, seed=5)
    sizes = []
    for mb, _ in model.minibatch_iterator(iter(tensors), cuda_device, 300, parallelize=False):
        sizes.append(int(mb["graph_data"]["node_to_graph_idx"].shape[0]))
    assert len(sizes) >= 2 and max(sizes) < 30000 + 35000  # the budget, not --minibatch-size 300, ends a minibatch
    _check_steps(model, nn, ref, tensors, cuda_device, 300, "config 1 (H=128, 30k-node budget)", max_batches=1)


def test_step_matches_oracle_without_message_bias(cuda_device, monkeypatch):
    """The message Linear_k bias is an assumption about upstream ptgnn (parity unpinned); the bias-free
    variant must be just as exact through the whole model (different state_dict keys, no bias terms in V)."""
    import functools

    from buglab.models import gnnlayerdefs
    from oracle import model_ref

    monkeypatch.setattr(gnnlayerdefs, "MlpMessagePassingLayer",
                        functools.partial(gnnlayerdefs.MlpMessagePassingLayer, use_message_bias=False))
    model, nn, _ref, data, tensors = _setup(64, cuda_device, n_graphs=6, make_ref=False)
    assert not any("edge_message_transformation_layers" in k and k.endswith("bias") for k in nn.state_dict())
    ref = model_ref.GnnBugLabModule(64, model.gnn_model.num_edge_types,
                                    len(model.gnn_model.node_representation_model.vocabulary),
                                    len(model._target_rewrite_ops), use_message_bias=False)
    ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
    _check_steps(model, nn, ref, tensors, cuda_device, 6, "no message bias, H=64", max_batches=1)


def test_step_matches_oracle_with_all_layer_outputs(cuda_device):
    """``use_all_gnn_layer_outputs=True`` (modelregistry.py:55, gnn.py:65-69,109-114): the heads read a Linear over the
    concatenation of the embedding and every layer's output, so every layer receives gradient from two places."""
    from buglab.models.modelregistry import load_model
    from buglab_b200.synthetic import SyntheticBugLabGenerator
    from oracle import model_ref

    hidden = 64
    torch.manual_seed(2)
    gen = SyntheticBugLabGenerator(seed=2, mean_nodes=250, min_nodes=40)
    data = [gen.sample() for _ in range(6)]
    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": 0.0,
                              "use_all_gnn_layer_outputs": True}, Path("/tmp/buglab_b200_test_all.pkl.gz"))
    model.gnn_model.node_representation_model.dropout_rate = 0.0
    model.compute_metadata(iter(copy.deepcopy(data)))
    nn = model.build_neural_module().to(cuda_device)
    assert any("summarization_layer" in k for k in nn.state_dict())
    ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types,
                                    len(model.gnn_model.node_representation_model.vocabulary),
                                    len(model._target_rewrite_ops), use_all_gnn_layer_outputs=True)
    ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
    tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
    _check_steps(model, nn, ref, tensors, cuda_device, 6, "all layer outputs, H=64", max_batches=1)


def test_optimizer_trajectory_matches_torch_adam(cuda_device):
    """5 real training steps: the fused flat Adam + global-norm clip + linear warm-up must move the weights exactly like
    torch.optim.Adam + clip_grad_norm_(0.5) + LambdaLR (the reference's optimiser stack, utils.py:51-66 / train.py:104)
    fed with the SAME gradients.  (Comparing whole trajectories against the CPU oracle instead would measure max-winner
    flips: Adam turns one re-routed gradient into an O(lr) weight change.)"""
    from buglab.models.utils import LinearWarmupScheduler, optimizer

    model, nn, ref, data, tensors = _setup(32, cuda_device, n_graphs=8)
    shadow = [torch.nn.Parameter(p.detach().clone()) for p in nn.parameters()]
    opt = optimizer(nn.parameters(), lr=1e-3)
    opt.max_grad_norm = 0.5
    sched = LinearWarmupScheduler(opt, num_warmup_steps=3)
    opt_ref = torch.optim.Adam(shadow, lr=1e-3)
    sched_ref = torch.optim.lr_scheduler.LambdaLR(opt_ref, lambda s: min(1.0, s / 3.0))
    mbs = list(model.minibatch_iterator(iter(tensors), cuda_device, 4, parallelize=False))
    losses = []
    for step in range(5):
        mb = mbs[step % len(mbs)][0]
        opt.zero_grad()
        loss = nn(**mb)
        loss.backward()
        for q, p in zip(shadow, nn.parameters()):
            q.grad = p.grad.detach().clone()
        opt.step()
        sched.step(0, step)
        torch.nn.utils.clip_grad_norm_(shadow, 0.5)
        opt_ref.step()
        sched_ref.step()
        losses.append(float(loss.detach()))
        for (name, p), q in zip(nn.named_parameters(), shadow):
            torch.testing.assert_close(p.detach(), q.detach(), atol=2e-6, rtol=1e-5, msg=lambda m, n=name: f"step {step} {n}: {m}")
    assert losses[-1] < losses[0] + 0.5  # training is not diverging
    # parameters still live in the flat buffer and the module still sees them
    assert all(p.data_ptr() >= opt.flat_param.data_ptr() for p in nn.parameters())


def test_predict_and_checkpoint_roundtrip(cuda_device, tmp_path):
    from buglab.models.evaluate import evaluate_predictions
    from buglab.models.gnn import GnnBugLabModel
    from oracle import model_ref

    model, nn, ref, data, tensors = _setup(32, cuda_device, n_graphs=7)
    preds = list(model.predict(iter(copy.deepcopy(data)), nn, cuda_device, parallelize=False))
    assert len(preds) == len(data)
    for point, loc, rewrites in preds:
        assert abs(sum(torch.tensor(list(loc.values())).exp()).item() - 1.0) < 1e-4
        assert len(rewrites) == len(point["candidate_rewrites"]) and None not in rewrites
    metrics = evaluate_predictions(preds)
    assert metrics["num_samples"] == len(data)
    path = tmp_path / "m.pkl.gz"
    model.save(path, nn)
    model2, nn2 = GnnBugLabModel.restore_model(path, cuda_device)
    preds2 = list(model2.predict(iter(copy.deepcopy(data)), nn2, cuda_device, parallelize=False))
    for (p1, l1, r1), (p2, l2, r2) in zip(preds, preds2):
        assert l1.keys() == l2.keys() and all(abs(l1[k] - l2[k]) < 1e-6 for k in l1)
        assert all(abs(a - b) < 1e-6 for a, b in zip(r1, r2))


def test_train_entry_point_runs(cuda_device, tmp_path):
    """python -m buglab.models.train gnn-mlp TRAIN VALID MODEL.pkl.gz on two tiny synthetic shards, then evaluate."""
    from buglab.models import evaluate, train
    from buglab_b200.synthetic import write_shards

    write_shards(str(tmp_path / "train"), 2, 12, seed=1, mean_nodes=150, min_nodes=40)
    write_shards(str(tmp_path / "valid"), 1, 6, seed=2, mean_nodes=150, min_nodes=40)
    model_path = tmp_path / "model.pkl.gz"
    train.main(["gnn-mlp", str(tmp_path / "train"), str(tmp_path / "valid"), str(model_path), "--max-num-epochs=2",
                "--minibatch-size=6", "--quiet", "--model-spec", '{"hidden_state_size": 32}'])
    assert model_path.exists()
    args = {"MODEL_FILENAME": str(model_path), "TEST_DATA_PATH": str(tmp_path / "valid"), "--limit-num-elements": None,
            "--sequential": True, "--azure-info": None}
    metrics = evaluate.run(args)
    assert metrics["num_samples"] == 6 and 0.0 <= metrics["localization_accuracy"] <= 1.0
