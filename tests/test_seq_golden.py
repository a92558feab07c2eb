"""Sequence models, host side: graph -> token-sequence projection, per-sample tensors and minibatch
packing of this repo's ``buglab.models.seqmodel.SeqBugLabModel`` against the REAL reference's
(tests/golden/seq_model.npz, written by tests/golden/make_seq_model_golden.py).  Integer bookkeeping: bit-exact."""
import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
SPEC = {"modelName": "seq-great", "hidden_state_size": 32, "dropout_rate": 0.0, "num_layers": 2, "num_heads": 4,
        "intermediate_dimension_size": 64, "max_seq_size": 200}


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, np.integer):
        return int(x)
    if isinstance(x, np.floating):
        return float(x)
    return x


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN_DIR, "seq_model.npz"), allow_pickle=True)


@pytest.fixture(scope="module")
def samples():
    from buglab.utils.msgpackutils import load_msgpack_l_gz

    return lambda: list(load_msgpack_l_gz(os.path.join(GOLDEN_DIR, "seq_samples.msgpack.l.gz")))


@pytest.fixture(scope="module")
def mirror(samples):
    import logging
    from pathlib import Path

    from buglab.models.modelregistry import load_model

    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)  # two fixture samples are rejected on purpose
    model, nn, needs_metadata = load_model(dict(SPEC), Path("/tmp/_seq_mirror.pkl.gz"))
    assert nn is None and needs_metadata
    model.compute_metadata(iter(samples()))
    return model


def test_metadata_matches_reference(golden, mirror):
    assert list(mirror.token_embedder.vocabulary.id_to_token) == list(golden["meta/vocabulary"])
    assert set(mirror.edge_types) == set(golden["meta/edge_types_in_reference_order"])
    assert list(mirror.edge_types) == sorted(mirror.edge_types)      # deterministic numbering (the reference's is hash-seeded)


def test_projection_and_sample_tensors_bit_exact(golden, mirror, samples):
    tensorized = [mirror.tensorize(dp) for dp in samples()]
    assert [t is None for t in tensorized] == golden["meta/dropped"].tolist() == [False] * 6 + [True, True]
    for i, t in enumerate(tensorized):
        if t is None:
            continue
        expected = json.loads(str(golden[f"sample/{i}"]))
        mine = t._asdict()
        mine["target_subtokens_ids"] = [np.asarray(x).tolist() for x in mine["target_subtokens_ids"]]
        mine = json.loads(json.dumps(jsonable(mine)))
        assert mine.keys() == expected.keys()
        for key in expected:
            assert mine[key] == expected[key], (i, key)
        # dict ORDER matters too: relations are packed in this order, node mappings are handed to predict()
        assert list(mine["intra_token_edges"]) == list(expected["intra_token_edges"])
        assert list(mine["node_mappings"]) == list(expected["node_mappings"])


def test_minibatch_tensors_bit_exact(golden, mirror, samples):
    mb = mirror.initialize_minibatch()
    for t in (mirror.tensorize(dp) for dp in samples()):
        if t is not None:
            assert mirror.extend_minibatch_with(t, mb) is True
    mb = mirror.finalize_minibatch(mb, "cpu")
    tensor_keys = sorted(k[3:] for k in golden.files if k.startswith("mb/") and k != "mb/edge_type_names")
    assert sorted(k for k, v in mb.items() if isinstance(v, torch.Tensor)) == tensor_keys
    for key in tensor_keys:
        expected = golden[f"mb/{key}"]
        got = mb[key]
        assert got.dtype == torch.from_numpy(expected).dtype and tuple(got.shape) == expected.shape, key
        if key == "edge_types":  # numbered differently (sorted here, set order there): compare by relation NAME
            assert [mirror.edge_types[int(i)] for i in got] == list(golden["mb/edge_type_names"])
        else:
            assert np.array_equal(got.numpy(), expected), key
    assert mb["input_sequence_ids"].shape[0] == 6 and mb["edges"].shape[1] == 3


def test_projection_rejects_malformed_graphs(mirror, samples):
    from buglab.models.seqmodel import TokenProjectionError, project_graph_to_tokens

    good = samples()[1]
    labels, position, relations, refs = project_graph_to_tokens(good["graph"])
    chain = [a for a, _ in good["graph"]["edges"]["NextToken"]] + [good["graph"]["edges"]["NextToken"][-1][1]]
    assert labels == [good["graph"]["nodes"][t] for t in chain]
    assert all(position[t] == i for i, t in enumerate(chain))
    assert set(relations) == {"NextMayUse", "LastMayWrite", "ComputedFrom", "ControlFlowNext"}
    assert refs == [position[n] for n in good["graph"]["reference_nodes"]]
    # symbols sit at their first occurrence
    for token, symbol in good["graph"]["edges"]["OccurrenceOf"]:
        assert position[symbol] <= position[token]

    cyclic = samples()[1]
    cyclic["graph"]["edges"]["NextToken"].append((chain[-1], chain[3]))
    with pytest.raises(TokenProjectionError):
        project_graph_to_tokens(cyclic["graph"])
    assert mirror.tensorize(cyclic) is None

    headless = samples()[1]
    headless["graph"]["edges"]["NextToken"].append((chain[-1], chain[0]))
    with pytest.raises(TokenProjectionError):
        project_graph_to_tokens(headless["graph"])

    no_operator = samples()[1]
    g = no_operator["graph"]
    binop = g["nodes"].index("BinaryOperation")
    for i, e in enumerate(g["edges"]["Child"]):
        if e[0] == binop and e[2] == "operator":
            g["nodes"][e[1]] = "??"
    assert mirror.tensorize(no_operator) is None


def test_unknown_layer_type_is_rejected():
    from buglab.models.seqmodel import SeqBugLabModel

    with pytest.raises(ValueError):
        SeqBugLabModel(32, 100, 0.0, layer_type="lstm")


# ------------------------------------------------------------------------------------------------ the module (oracle)
def _reference_minibatch(golden, mirror, layer_type):
    mb = {k[3:]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith("mb/") and k != "mb/edge_type_names"}
    # relation ids in the numbering the reference module was built with
    order = list(golden[f"{layer_type}/edge_types_in_reference_order"])
    mb["edge_types"] = torch.tensor([order.index(n) for n in golden[f"{layer_type}/edge_type_names"]], dtype=torch.int64)
    return mb, order


@pytest.mark.parametrize("layer_type", ["great", "rat"])
def test_oracle_module_matches_reference(golden, mirror, layer_type):
    """Loss, token representations, every log-probability and every parameter gradient of the real SeqBugLabModule."""
    from oracle.seq_model_ref import SeqBugLabModule

    mb, order = _reference_minibatch(golden, mirror, layer_type)
    prefix = f"{layer_type}/param/"
    state = {k[len(prefix):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith(prefix)}
    rows = state["_SeqBugLabModule__positional_encoding"].shape[1]
    module = SeqBugLabModule(vocabulary_size=state["_SeqBugLabModule__token_embedder._SubtokenUnitEmbedder__embeddings.weight"].shape[0],
                             embedding_dim=SPEC["hidden_state_size"], num_edge_types=len(order), num_layers=SPEC["num_layers"],
                             num_heads=SPEC["num_heads"], intermediate_dimension=SPEC["intermediate_dimension_size"],
                             rewrite_vocabulary_size=state["_text_repair_module._TextRepairModule__text_rewrite_embeddings.weight"].shape[0],
                             layer_type=layer_type, positional_rows=rows)
    missing, unexpected = module.load_state_dict(state, strict=False)
    assert not missing and not unexpected, (missing, unexpected)
    module.train()  # dropout is 0; train mode as in the fixture
    loss, details = module(**mb, return_details=True)
    assert abs(float(loss.detach()) - float(golden[f"{layer_type}/loss"])) < 2e-5
    lengths = mb["token_sequence_lengths"]
    valid = torch.arange(mb["input_sequence_ids"].shape[1])[None, :] < lengths[:, None]
    rep_ref = torch.from_numpy(golden[f"{layer_type}/output_representation"])
    assert float((details["output_representation"].detach() - rep_ref)[valid].abs().max()) < 1e-5
    for mine, theirs in (("localization_logprobs", "localization_logprobs"), ("text_logprobs", "text_logprobs"),
                         ("varmisuse_logprobs", "varmisuse_logprobs"), ("argswap_logprobs", "argswap_logprobs")):
        ref = torch.from_numpy(golden[f"{layer_type}/{theirs}"])
        assert details[mine].shape == ref.shape and ref.numel() > 0, mine
        assert float((details[mine].detach() - ref).abs().max()) < 2e-5, mine
    assert torch.equal(details["localization_groups"], torch.from_numpy(golden[f"{layer_type}/localization_groups"]))
    loss.backward()
    checked = 0
    for name, p in module.named_parameters():
        key = f"{layer_type}/grad/{name}"
        if key not in golden.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name      # norm2 (unused in post-norm mode)
            continue
        ref = torch.from_numpy(golden[key])
        tol = 2e-5 * (float(ref.abs().max()) + 1e-12) + 2e-6
        assert float((p.grad - ref).abs().max()) <= tol, (name, float((p.grad - ref).abs().max()), tol)
        checked += 1
    assert checked >= 35


def test_oracle_module_on_the_mirrors_own_minibatch(golden, mirror, samples):
    """The mirror's packed minibatch (its own relation numbering) drives the oracle to the reference's loss once the
    relation-indexed bias tables are permuted accordingly — numbering is the only difference between the two hosts."""
    from oracle.seq_model_ref import SeqBugLabModule

    packed = mirror.initialize_minibatch()
    for t in (mirror.tensorize(dp) for dp in samples()):
        if t is not None:
            mirror.extend_minibatch_with(t, packed)
    mb = mirror.finalize_minibatch(packed, "cpu")
    order = list(golden["great/edge_types_in_reference_order"])
    perm = torch.tensor([order.index(kind) for kind in mirror.edge_types])   # mirror id -> reference id
    prefix = "great/param/"
    state = {k[len(prefix):]: torch.from_numpy(golden[k]) for k in golden.files if k.startswith(prefix)}
    for k in list(state):
        if "edge_attention_biases" in k or "edge_value_biases" in k:
            state[k] = state[k][perm]
    module = SeqBugLabModule(vocabulary_size=len(mirror.token_embedder.vocabulary), embedding_dim=32, num_edge_types=len(order),
                             num_layers=2, num_heads=4, intermediate_dimension=64,
                             rewrite_vocabulary_size=len(mirror._target_rewrite_ops), layer_type="great",
                             positional_rows=state["_SeqBugLabModule__positional_encoding"].shape[1])
    module.load_state_dict(state)
    tensors = {k: v for k, v in mb.items() if isinstance(v, torch.Tensor)}
    loss = module(**tensors)
    assert abs(float(loss.detach()) - float(golden["great/loss"])) < 2e-5


# ------------------------------------------------------------------------------------------------ the module (mirror)
@pytest.fixture()
def cpu_kernels(monkeypatch, request):
    """Runs the mirror's module on the CPU: every CUDA entry point it reaches is replaced by the oracle's restatement of
    that one op (segment ops, LayerNorm, subtoken max-pool) or, for the new attention kernels, by the host emulation of
    the kernel source.  What is under test is the WIRING — parameter names, shapes, op order, index handling."""
    from buglab_b200 import ops
    from oracle import mp_ref, scatter_ref

    request.getfixturevalue("host_backend")  # attention -> host emulation of csrc/seq_attention_core.h
    monkeypatch.setattr(ops, "layer_norm", lambda x, g, b, eps=1e-5: torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps))
    monkeypatch.setattr(ops, "dense_linear", lambda x, w: torch.nn.functional.linear(x, w))
    monkeypatch.setattr(ops, "subtoken_maxpool", lambda emb, ids, lens, p_drop=0.0, training=False:
                        mp_ref.subtoken_maxpool_ref(emb, ids.long(), lens.long()))
    monkeypatch.setattr(ops, "segment_log_softmax", lambda src, index, eps=1e-12, num_segments=None:
                        scatter_ref.scatter_log_softmax(src, index.long()))
    monkeypatch.setattr(ops, "segment_minmax", lambda src, index, dim=-1, dim_size=None, is_min=False:
                        (scatter_ref.scatter_min if is_min else scatter_ref.scatter_max)(src, index.long(), dim, dim_size))
    monkeypatch.setattr(ops, "segment_sum", lambda src, index, dim=-1, dim_size=None:
                        scatter_ref.scatter_sum(src, index.long(), dim, dim_size))


@pytest.mark.parametrize("layer_type", ["great", "rat", "transformer", "gru"])
def test_mirror_module_reproduces_reference_on_cpu_kernels(golden, samples, cpu_kernels, layer_type):
    import logging
    from pathlib import Path

    from buglab.models.modelregistry import load_model

    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)
    model, _, _ = load_model(dict(SPEC, modelName=f"seq-{layer_type}"), Path("/tmp/_seq_mirror2.pkl.gz"))
    model.compute_metadata(iter(samples()))
    nn = model.build_neural_module()
    nn._argswap_module._input_dim = SPEC["hidden_state_size"]
    order = list(golden[f"{layer_type}/edge_types_in_reference_order"])
    perm = torch.tensor([order.index(kind) for kind in model.edge_types])      # mirror id -> reference id
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
    missing, unexpected = nn.load_state_dict(state, strict=False)
    assert not missing and not unexpected, (missing, unexpected)      # same parameter names as the reference module
    nn.train()

    packed = model.initialize_minibatch()
    for t in (model.tensorize(dp) for dp in samples()):
        if t is not None:
            model.extend_minibatch_with(t, packed)
    mb = model.finalize_minibatch(packed, "cpu")
    loss = nn(**mb)
    assert abs(float(loss.detach()) - float(golden[f"{layer_type}/loss"])) < 3e-5
    loss.backward()
    checked = 0
    for name, p in nn.named_parameters():
        key = f"{layer_type}/grad/{name}"
        if key not in golden.files:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        ref = torch.from_numpy(golden[key])
        grad = p.grad
        if "positional_encoding" in name:
            grad = grad[:, :rows]
        if "edge_attention_biases" in name or "edge_value_biases" in name:
            grad = grad[inverse]
        tol = 3e-5 * (float(ref.abs().max()) + 1e-12) + 3e-6
        assert float((grad - ref).abs().max()) <= tol, (name, float((grad - ref).abs().max()), tol)
        checked += 1
    assert checked >= 35
    metrics = nn.report_metrics()
    assert abs(metrics["Loss"] - float(golden[f"{layer_type}/loss"])) < 3e-5

    # inference through predict(): per graph node and per rewrite log-probabilities, in the reference's order
    expected = json.loads(str(golden[f"{layer_type}/predictions"]))
    got = list(model.predict(iter(samples()), nn, "cpu", parallelize=False))
    assert [dp["graph"]["path"] for dp, _, _ in got] == [e["path"] for e in expected] and len(got) == 6
    for (dp, locations, rewrites), e in zip(got, expected):
        assert [str(k) for k in locations] == list(e["locations"])           # same nodes, same order, NO_BUG (-1) last
        assert max(abs(float(locations[int(k)]) - v) for k, v in e["locations"].items()) < 3e-5
        assert len(rewrites) == len(e["rewrites"]) == len(dp["candidate_rewrites"])
        assert max(abs(float(a) - b) for a, b in zip(rewrites, e["rewrites"])) < 3e-5


def test_shard_dataset_serves_sequence_models_through_the_host_path(mirror, samples, tmp_path):
    """The native shard tensoriser is for the graph models; a sequence model asking a ShardDataset for tensors gets the
    reference-shaped chain (explicitly, with a log line), and the trainer-facing output shape is the same."""
    from buglab.utils.msgpackutils import save_msgpack_l_gz
    from buglab_b200.shards import ShardDataset
    from dpu_utils.utils import RichPath

    save_msgpack_l_gz(samples(), str(tmp_path / "a.msgpack.l.gz"))
    dataset = ShardDataset(RichPath.create(str(tmp_path)), num_threads=1)
    got = list(dataset.tensorized(mirror))
    expected = [t for t in (mirror.tensorize(dp) for dp in samples()) if t is not None]
    assert len(got) == len(expected) == 6 and all(raw is None for _, raw in got)
    for (t, _), e in zip(got, expected):
        assert t.candidate_location_idxs.tolist() == e.candidate_location_idxs.tolist()
        assert t.intra_token_edges == e.intra_token_edges and t.node_mappings == e.node_mappings
    assert dataset.tensorizer is None      # no native tensoriser was built for this model family


def test_registry_defaults_build_and_checkpoints_round_trip(tmp_path):
    """Every sequence model name builds with the registry's default spec (dropout 0.1 included) and survives the reference's
    checkpoint format (gzip(torch.save((model, nn))), modelregistry.py:147-156)."""
    import copy
    import logging
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab.models.seqmodel import SeqBugLabModel
    from buglab_b200.synthetic import SyntheticProgramGenerator

    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)
    for name in ("seq-great", "seq-rat", "seq-transformer", "seq-gru"):
        model, _, _ = load_model({"modelName": name, "hidden_state_size": 64, "num_layers": 1}, Path(str(tmp_path / "m.pkl.gz")))
        model.compute_metadata(SyntheticProgramGenerator(seed=0).samples(6))
        assert sum(p.numel() for p in model.build_neural_module().parameters()) > 0
    model, _, _ = load_model({"modelName": "seq-rat", "hidden_state_size": 32, "num_heads": 4, "num_layers": 2,
                              "buggy_samples_weight_spec": "warmdown(3, 0.5)"}, Path(str(tmp_path / "m.pkl.gz")))
    model.compute_metadata(SyntheticProgramGenerator(seed=0).samples(6))
    nn = model.build_neural_module()
    model.save(Path(str(tmp_path / "m.pkl.gz")), nn)
    restored, restored_nn = SeqBugLabModel.restore_model(Path(str(tmp_path / "m.pkl.gz")), torch.device("cpu"))
    assert restored.edge_types == model.edge_types
    assert all(torch.equal(a, b) for a, b in zip(nn.state_dict().values(), restored_nn.state_dict().values()))
    sample = SyntheticProgramGenerator(seed=3).sample()
    assert model.tensorize(copy.deepcopy(sample)).node_mappings == restored.tensorize(copy.deepcopy(sample)).node_mappings


def test_train_and_evaluate_entry_points_for_seq_great_on_cpu_kernels(cpu_kernels, monkeypatch, tmp_path):
    """``python -m buglab.models.train seq-great …`` then ``evaluate`` over synthetic program shards, with every CUDA entry
    point replaced as above and the fused optimiser by ``torch.optim.Adam``: the wiring of the trainer, the shard loader's
    host-language path for non-graph models, the sequence model's minibatching, metrics and ``predict`` (the GPU twin is
    tests/test_seq_attention_gpu.py::test_train_and_evaluate_entry_points_run_for_seq_great)."""
    import logging

    import buglab.models.train as train_mod
    import buglab.models.utils as utils_mod
    from buglab.models import evaluate
    from buglab_b200.synthetic import write_shards

    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)
    adam = lambda params, lr=0.0001: torch.optim.Adam(params, lr=lr)  # noqa: E731
    monkeypatch.setattr(utils_mod, "optimizer", adam)
    monkeypatch.setattr(train_mod, "optimizer", adam, raising=False)
    write_shards(str(tmp_path / "train"), 2, 8, seed=1, programs=True, statements=6)
    write_shards(str(tmp_path / "valid"), 1, 4, seed=2, programs=True, statements=6)
    model_path = tmp_path / "model.pkl.gz"
    train_mod.main(["seq-great", str(tmp_path / "train"), str(tmp_path / "valid"), str(model_path), "--max-num-epochs=1",
                    "--minibatch-size=4", "--quiet", "--sequential", "--model-spec",
                    '{"hidden_state_size": 32, "num_heads": 2, "num_layers": 1, "max_seq_size": 200, "intermediate_dimension_size": 64}'])
    assert model_path.exists()
    # evaluate.run insists on a CUDA device (no CPU path in the product): its two halves are driven here directly
    from pathlib import Path

    from buglab.models.gnn import GnnBugLabModel
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from dpu_utils.utils import RichPath

    model, nn = GnnBugLabModel.restore_model(Path(model_path), torch.device("cpu"))
    data = load_all_msgpack_l_gz(RichPath.create(str(tmp_path / "valid")), shuffle=False)
    metrics = evaluate.evaluate_predictions(model.predict(data, nn, torch.device("cpu"), parallelize=False), False, False)
    assert metrics["num_samples"] > 0 and 0.0 <= metrics["localization_accuracy"] <= 1.0
