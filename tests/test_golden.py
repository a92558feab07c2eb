"""Golden-vector tests: fixtures in tests/golden/ were produced by the REAL reference code (tests/golden/make_golden.py).

CPU part (not gpu): the host-side mirror (buglab.* / ptgnn.* of this repo) must reproduce the reference's integer
bookkeeping bit for bit, and the CPU oracle restatement must reproduce the reference's arithmetic.
GPU part: the CUDA path must reproduce the reference's outputs within 1e-4.
"""
import os
from pathlib import Path

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden", "gnn_mlp_h16.npz")
SAMPLES = os.path.join(HERE, "golden", "samples.msgpack.l.gz")
HIDDEN = 16


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN, allow_pickle=True)


def _load_samples():
    from buglab.utils.msgpackutils import load_msgpack_l_gz

    return list(load_msgpack_l_gz(SAMPLES))


@pytest.fixture(scope="module")
def mirror():
    """This repo's model built from the same samples and registry spec as the golden run."""
    from buglab.models.modelregistry import load_model

    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": HIDDEN, "dropout_rate": 0.0,
                              "node_representations": {"dropout_rate": 0.0, "min_freq_threshold": 1}},
                             Path("/tmp/golden_mirror.pkl.gz"))
    model.compute_metadata(iter(_load_samples()))
    tensorized = [t for t, _ in model.tensorize_dataset(iter(_load_samples()), parallelize=False)]
    return model, tensorized


def _state_dict(golden):
    return {k[len("state/"):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("state/")}


def test_metadata_matches_reference(golden, mirror):
    model, _ = mirror
    assert list(model.gnn_model.node_representation_model.vocabulary.id_to_token) == list(golden["meta/vocabulary"])
    assert list(model.gnn_model.edge_types) == list(golden["meta/edge_types"])
    assert list(model._target_rewrite_ops.id_to_token) == list(golden["meta/rewrite_ops"])
    assert model.gnn_model.num_edge_types == 2 * len(golden["meta/edge_types"]) + 1


def test_per_sample_rewrite_tables_bit_exact(golden, mirror):
    _, tensorized = mirror
    assert len(tensorized) == 8
    for i, t in enumerate(tensorized):
        for field in t._fields:
            if field in ("graph_data", "rewrite_logprobs"):
                continue
            v = getattr(t, field)
            np.testing.assert_array_equal(np.array(-1 if v is None else v, dtype=np.int64), golden[f"tensorized/{i}/{field}"],
                                          err_msg=f"sample {i} {field}")
        for name, ids in t.graph_data.reference_nodes.items():
            np.testing.assert_array_equal(np.asarray(ids, dtype=np.int64), golden[f"tensorized/{i}/ref/{name}"], err_msg=name)
        assert t.graph_data.num_nodes == int(golden[f"tensorized/{i}/num_nodes"])


def _pack(model, tensorized, device):
    mb = model.initialize_minibatch()
    for t in tensorized:
        model.extend_minibatch_with(t, mb)
    return model.finalize_minibatch(mb, device)


def test_minibatch_index_tensors_bit_exact(golden, mirror):
    model, tensorized = mirror
    mb = _pack(model, tensorized, "cpu")
    checked = 0
    for key in golden.files:
        if not key.startswith("mb/") or key.startswith("mb/graph/"):
            continue
        name = key[3:]
        got = mb[name]
        assert got.dtype == torch.from_numpy(golden[key]).dtype, (name, got.dtype)
        np.testing.assert_array_equal(got.numpy(), golden[key], err_msg=name)
        checked += 1
    assert checked >= 13
    g = mb["graph_data"]
    for name in g["reference_node_ids"]:
        np.testing.assert_array_equal(g["reference_node_ids"][name].numpy(), golden[f"mb/graph/ref_ids/{name}"])
        np.testing.assert_array_equal(g["reference_node_graph_idx"][name].numpy(), golden[f"mb/graph/ref_graph/{name}"])
    np.testing.assert_array_equal(g["node_to_graph_idx"].numpy(), golden["mb/graph/node_to_graph_idx"])
    for k, (s, t) in enumerate(g["adjacency_lists"]):
        np.testing.assert_array_equal(s.numpy(), golden[f"mb/graph/adj/{k}/src"])
        np.testing.assert_array_equal(t.numpy(), golden[f"mb/graph/adj/{k}/tgt"])


def _oracle(golden, mirror, dtype=torch.float32):
    from oracle import model_ref

    model, _ = mirror
    ref = model_ref.GnnBugLabModule(HIDDEN, model.gnn_model.num_edge_types,
                                    len(model.gnn_model.node_representation_model.vocabulary), len(model._target_rewrite_ops))
    ref.load_state_dict(_state_dict(golden), strict=True)
    return ref.to(dtype)


def test_oracle_restatement_matches_reference(golden, mirror):
    """Pins oracle/model_ref.py (the in-repo part of the oracle) to the reference's own outputs."""
    from oracle import model_ref

    model, tensorized = mirror
    ref = _oracle(golden, mirror)
    mb = model_ref.minibatch_to_cpu(_pack(model, tensorized, "cpu"))
    loss, det = ref(**mb, return_details=True)
    loss.backward()
    tol = dict(atol=1e-5, rtol=1e-5)
    torch.testing.assert_close(loss.detach(), torch.from_numpy(golden["out/loss"]), **tol)
    torch.testing.assert_close(det["node_states"].detach(), torch.from_numpy(golden["out/node_states"]), **tol)
    torch.testing.assert_close(det["localization_logprobs"].detach(), torch.from_numpy(golden["out/localization_logprobs"]), **tol)
    assert torch.equal(det["localization_groups"], torch.from_numpy(golden["out/localization_groups"]))
    for name in ("text", "varmisuse", "argswap"):
        torch.testing.assert_close(det[f"{name}_logprobs"].detach(), torch.from_numpy(golden[f"out/{name}_logprobs"]), **tol)
    for n, p in ref.named_parameters():
        if "grad/" + n in golden.files:
            torch.testing.assert_close(p.grad, torch.from_numpy(golden["grad/" + n]), atol=1e-5, rtol=1e-4, msg=lambda m, n=n: f"{n}: {m}")


def test_schedules_match_reference(golden):
    from buglab.models.modelregistry import buggy_sample_weight_schedule
    from buglab.models.utils import LinearWarmupScheduler

    fn = buggy_sample_weight_schedule("warmdown(4, 0.25)")
    np.testing.assert_allclose([fn(e) for e in range(8)], golden["sched/warmdown"], rtol=0, atol=0)
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-4)
    warm = LinearWarmupScheduler(opt, num_warmup_steps=5)
    lrs = []
    for step in range(8):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        warm.step(0, step)
    np.testing.assert_allclose(lrs, golden["sched/warmup_lrs"], rtol=1e-12)


@pytest.mark.gpu
def test_gpu_path_matches_reference_outputs(golden, mirror, cuda_device):
    from oracle import parity

    model, tensorized = mirror
    nn = model.build_neural_module()
    nn.load_state_dict(_state_dict(golden), strict=True)
    nn.to(cuda_device).train()
    mb = _pack(model, tensorized, cuda_device)
    loss = nn(**mb)
    loss.backward()
    parity.assert_forward_close(loss, torch.from_numpy(golden["out/loss"]), "loss")
    with torch.no_grad():
        groups, logprobs, gnn_output, _ = nn.compute_localization_logprobs(mb["graph_data"])
        parity.assert_forward_close(gnn_output.output_node_representations, torch.from_numpy(golden["out/node_states"]), "node states")
        parity.assert_forward_close(logprobs, torch.from_numpy(golden["out/localization_logprobs"]), "localization logprobs")
        assert torch.equal(groups.cpu(), torch.from_numpy(golden["out/localization_groups"]))
        swap, text, misuse, sel = nn._compute_repair_logprobs(
            gnn_output, mb["target_rewrites"], mb["rewrite_to_location_group"], mb["candidate_symbol_to_location_group"],
            mb["swapped_pair_to_call_location_group"])
        parity.assert_forward_close(text, torch.from_numpy(golden["out/text_logprobs"]), "text")
        parity.assert_forward_close(misuse, torch.from_numpy(golden["out/varmisuse_logprobs"]), "varmisuse")
        parity.assert_forward_close(swap, torch.from_numpy(golden["out/argswap_logprobs"]), "argswap")
        assert torch.equal(torch.cat([sel[1], sel[2], sel[0]]).cpu(), torch.from_numpy(golden["out/selected_fix_masks"]))
    for n, p in nn.named_parameters():
        if "grad/" + n in golden.files:  # the reference's own routing cannot be forced: norm-wise criterion
            parity.assert_grad_close_normwise(p.grad, torch.from_numpy(golden["grad/" + n]), n)
    metrics = nn.report_metrics()
    assert abs(metrics["Loss"] - float(golden["out/metric_loss"])) < 1e-4
    assert abs(metrics["Localization Accuracy"] - float(golden["out/metric_localization_accuracy"])) < 1e-9
    # predict(): per-sample unpacking, all-location rewrites
    preds = list(model.predict(iter(_load_samples()), nn, cuda_device, parallelize=False))
    assert len(preds) == 8
    for i, (_p, loc, rewrites) in enumerate(preds):
        keys = sorted(loc.keys())
        np.testing.assert_array_equal(np.array(keys), golden[f"predict/{i}/location_nodes"])
        np.testing.assert_allclose([loc[k] for k in keys], golden[f"predict/{i}/location_logprobs"], atol=1e-4, rtol=1e-4)
        np.testing.assert_allclose(rewrites, golden[f"predict/{i}/rewrite_logprobs"], atol=1e-4, rtol=1e-4)


# ---- selector ("bug generator") mode: all-location rewrites + compute_generator_loss (reference utils.py:101-179) ----
SELECTOR_SAMPLES = os.path.join(HERE, "golden", "selector_samples.msgpack.l.gz")
LOSS_TYPES = ("classify-max-loss", "norm-kl", "norm-rmse", "expectation")


def _selector_minibatch(model, device):
    from buglab.utils.msgpackutils import load_msgpack_l_gz

    with model._tensorize_all_location_rewrites():
        tensorized = [t for t, _ in model.tensorize_dataset(iter(load_msgpack_l_gz(SELECTOR_SAMPLES)), parallelize=False)]
        return _pack(model, tensorized, device)


def test_selector_minibatch_bit_exact(golden, mirror):
    model, _ = mirror
    mb = _selector_minibatch(model, "cpu")
    checked = 0
    for key in golden.files:
        if key.startswith("selector_mb/"):
            got = mb[key[len("selector_mb/"):]]
            if got.dtype.is_floating_point:
                np.testing.assert_array_equal(got.numpy(), golden[key])  # rewrite_logprobs incl. -inf, exact
            else:
                np.testing.assert_array_equal(got.numpy(), golden[key], err_msg=key)
            checked += 1
    assert checked >= 14 and "rewrite_logprobs" in mb


@pytest.mark.parametrize("loss_type", LOSS_TYPES)
def test_oracle_generator_loss_matches_reference(golden, mirror, loss_type):
    from oracle import model_ref

    model, _ = mirror
    ref = _oracle(golden, mirror)
    ref.generator_loss_type = loss_type
    mb = model_ref.minibatch_to_cpu(_selector_minibatch(model, "cpu"))
    loss = ref(**mb)
    loss.backward()
    torch.testing.assert_close(loss.detach(), torch.from_numpy(golden[f"selector/{loss_type}/loss"]), atol=1e-5, rtol=1e-5)
    grad = dict(ref.named_parameters())["_GnnBugLabModule__localization_module._l1.weight"].grad
    torch.testing.assert_close(grad, torch.from_numpy(golden[f"selector/{loss_type}/grad_l1"]), atol=1e-5, rtol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("loss_type", LOSS_TYPES)
def test_gpu_generator_loss_matches_reference(golden, mirror, cuda_device, loss_type):
    from oracle import parity

    model, _ = mirror
    nn = model.build_neural_module()
    nn.load_state_dict(_state_dict(golden), strict=True)
    nn.to(cuda_device).train()
    nn._GnnBugLabModule__generator_loss_type = loss_type
    mb = _selector_minibatch(model, cuda_device)
    loss = nn(**mb)
    loss.backward()
    parity.assert_forward_close(loss, torch.from_numpy(golden[f"selector/{loss_type}/loss"]), f"{loss_type} loss")
    grad = dict(nn.named_parameters())["_GnnBugLabModule__localization_module._l1.weight"].grad
    parity.assert_grad_close_normwise(grad, torch.from_numpy(golden[f"selector/{loss_type}/grad_l1"]), f"{loss_type} d l1.weight")
    assert abs(nn.report_metrics()["Loss"] - float(golden[f"selector/{loss_type}/loss"])) < 1e-4
