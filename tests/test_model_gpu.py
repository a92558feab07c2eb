"""GPU parity of the whole gnn-mlp step against the CPU oracle, through the reference-facing operator surface
(ptgnn classes + buglab model mirror).  fp32 tolerance 1e-4 (BASELINE.json north_star); indices bit-exact."""
import copy
import os
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
TOL = dict(atol=1e-4, rtol=1e-4)


def _setup(hidden, device, n_graphs=10, mean_nodes=250, seed=0, dropout=0.0):
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
    ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types,
                                    len(model.gnn_model.node_representation_model.vocabulary),
                                    len(model._target_rewrite_ops))
    ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
    tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
    return model, nn, ref, data, tensors


@pytest.mark.parametrize("hidden", [32, 128])
def test_step_matches_oracle(cuda_device, hidden):
    from oracle import model_ref

    model, nn, ref, data, tensors = _setup(hidden, cuda_device)
    for mb, _raw in model.minibatch_iterator(iter(tensors), cuda_device, 5, parallelize=False):
        nn.zero_grad(); ref.zero_grad()
        nn.train()
        loss = nn(**mb)
        loss.backward()
        mb_cpu = model_ref.minibatch_to_cpu(mb)
        loss_ref, det = ref(**mb_cpu, return_details=True)
        loss_ref.backward()
        ref64 = copy.deepcopy(ref).double()  # fp64 referee for routing-ambiguous gradients (oracle/parity.py)
        ref64.zero_grad()
        ref64(**mb_cpu).backward()
        ref64_params = dict(ref64.named_parameters())
        torch.testing.assert_close(loss.cpu(), loss_ref, **TOL)
        groups, lp, gnn_out, _ = nn.compute_localization_logprobs(mb["graph_data"])
        torch.testing.assert_close(gnn_out.output_node_representations.detach().cpu(), det["node_states"].detach(), **TOL)
        torch.testing.assert_close(lp.detach().cpu(), det["localization_logprobs"].detach(), **TOL)
        assert torch.equal(groups.cpu(), det["localization_groups"])
        from oracle import parity

        ref_params = dict(ref.named_parameters())
        for name, p in nn.named_parameters():
            g_ref = ref_params[name].grad
            if g_ref is None:
                assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
                continue
            parity.assert_grad_close(p.grad, g_ref, name, expected_fp64=ref64_params[name].grad)


def test_training_trajectory_matches_oracle(cuda_device):
    """5 optimiser steps: fused flat Adam + clip + warm-up on the GPU vs torch Adam + clip_grad_norm_ on the CPU oracle."""
    from buglab.models.utils import LinearWarmupScheduler, optimizer
    from oracle import model_ref

    model, nn, ref, data, tensors = _setup(32, cuda_device, n_graphs=8)
    opt = optimizer(nn.parameters(), lr=1e-3)
    opt.max_grad_norm = 0.5
    sched = LinearWarmupScheduler(opt, num_warmup_steps=3)
    opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-3)
    sched_ref = torch.optim.lr_scheduler.LambdaLR(opt_ref, lambda s: min(1.0, s / 3.0))
    mbs = list(model.minibatch_iterator(iter(tensors), cuda_device, 4, parallelize=False))
    for step in range(5):
        mb = mbs[step % len(mbs)][0]
        opt.zero_grad()
        loss = nn(**mb)
        loss.backward()
        opt.step()
        sched.step(0, step)
        loss_ref = model_ref.train_step_ref(ref, opt_ref, model_ref.minibatch_to_cpu(mb), 0.5)
        sched_ref.step()
        assert abs(float(loss) - loss_ref) < 5e-4, (step, float(loss), loss_ref)
    from oracle import parity

    ref_sd = ref.state_dict()
    for k, v in nn.state_dict().items():
        # Adam normalises the update, so a flipped max-winner moves a weight by ~lr regardless of its gradient size
        frac_bad, rel_l2, _ = parity.grad_mismatch(v, ref_sd[k])
        assert rel_l2 < 1e-2 and frac_bad < 0.05, (k, frac_bad, rel_l2)


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
