"""Host-side logic on CPU: CLI parsing, vocabulary, identifier splitting, shard format, synthetic schema, packing
offsets against a by-hand computation, evaluation metrics, trainer loop and checkpoint round trip."""
import copy
import gzip
import os
from pathlib import Path

import msgpack
import numpy as np
import pytest
import torch


def test_docopt_matches_reference_usage():
    from docopt import DocoptExit, docopt

    from buglab.models import train

    args = docopt(train.__doc__, ["gnn-mlp", "tr", "va", "m.pkl.gz", "--sequential", "--minibatch-size=5", "--max-num-epochs", "3"])
    assert args["MODEL_NAME"] == "gnn-mlp" and args["MODEL_FILENAME"] == "m.pkl.gz"
    assert args["--sequential"] is True and args["--quiet"] is False and args["--amp"] is False
    assert args["--minibatch-size"] == "5" and args["--max-num-epochs"] == "3" and args["--validate-after"] == "1000000"
    assert args["--restore-path"] is None
    with pytest.raises(DocoptExit):
        docopt(train.__doc__, ["gnn-mlp", "only-two"])


def test_vocabulary_and_splitting():
    from dpu_utils.codeutils import split_identifier_into_parts as split
    from dpu_utils.mlutils import Vocabulary

    assert split("fooBar_baz2") == ["foo", "bar", "baz", "2"]
    assert split("HTTPResponseCode") == ["http", "response", "code"]
    assert split("__") == ["__"] and split("x") == ["x"]
    from collections import Counter

    v = Vocabulary.create_vocabulary(Counter({"a": 5, "b": 5, "c": 1, "d": 9}), max_size=4, count_threshold=2, add_pad=True)
    assert v.id_to_token == ["%PAD%", "%UNK%", "d", "a"]  # ties broken by token, capped at max_size
    assert v.get_id_or_unk("zzz") == 1 and v.get_id_or_unk("d") == 2
    ops = Vocabulary.create_vocabulary(frozenset({"+", "-", "and"}), max_size=3, count_threshold=0, add_unk=False)
    assert ops.id_to_token == ["+", "-", "and"]
    with pytest.raises(KeyError):
        ops.get_id_or_unk("nope")


def test_msgpack_shard_wire_format(tmp_path):
    from dpu_utils.utils import RichPath

    from buglab.utils.msgpackutils import load_all_msgpack_l_gz, load_msgpack_l_gz, save_msgpack_l_gz

    data = [{"a": 1, "b": [1, 2, "x"]}, None, {"c": {"d": (1, 2)}}]
    path = tmp_path / "s.msgpack.l.gz"
    save_msgpack_l_gz(data, path)
    # byte-level: gzip stream of concatenated msgpack objects (reference msgpackutils.py:17-21)
    with gzip.open(path, "rb") as f:
        raw = f.read()
    assert raw == b"".join(msgpack.Packer(use_bin_type=True).pack(e) for e in data)
    assert [dict(x) if x is not None else None for x in load_msgpack_l_gz(path)][0] == {"a": 1, "b": [1, 2, "x"]}
    assert len(list(load_all_msgpack_l_gz(RichPath.create(str(tmp_path))))) == 2  # None elements are dropped
    # rank sharding covers every element exactly once
    for i in range(3):
        save_msgpack_l_gz([{"i": i, "j": j} for j in range(5)], tmp_path / f"t{i}.msgpack.l.gz")
    seen = []
    for r in range(2):
        seen += [(e["i"], e["j"]) for e in load_all_msgpack_l_gz(RichPath.create(str(tmp_path)), rank=r, world_size=2) if "i" in e]
    assert sorted(seen) == [(i, j) for i in range(3) for j in range(5)]


def test_synthetic_sample_schema():
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    gen = SyntheticBugLabGenerator(seed=3, mean_nodes=200, min_nodes=50)
    s = gen.sample()
    g = s["graph"]
    n = len(g["nodes"])
    assert set(g) >= {"nodes", "edges", "path", "text", "reference_nodes", "code_range"}
    assert len(s["candidate_rewrites"]) == len(s["candidate_rewrite_metadata"]) == len(g["reference_nodes"]) >= 20
    nt = g["edges"]["NextToken"]
    heads = set(a for a, _ in nt) - set(b for _, b in nt)
    assert len(heads) == 1  # one chain (reference tests/test_extraction.py:49-52)
    for kind, edges in g["edges"].items():
        assert all(0 <= e[0] < n and 0 <= e[1] < n for e in edges), kind
    assert all(len(e) == 3 for e in g["edges"]["Child"])
    assert s["target_fix_action_idx"] is None or 0 <= s["target_fix_action_idx"] < len(g["reference_nodes"])
    # deterministic
    assert SyntheticBugLabGenerator(seed=3, mean_nodes=200, min_nodes=50).sample() == s


def _tiny_model(samples, hidden=8):
    from buglab.models.modelregistry import load_model

    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": 0.0,
                              "node_representations": {"min_freq_threshold": 1}}, Path("/tmp/tiny.pkl.gz"))
    model.compute_metadata(iter(copy.deepcopy(samples)))
    return model


def test_two_graph_minibatch_offsets_by_hand():
    """node ids of the second graph are shifted by the first graph's node count, candidate / group /
    rewrite indices by the running counters (gnn.py:474-536)."""
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    gen = SyntheticBugLabGenerator(seed=5, mean_nodes=80, min_nodes=40, max_nodes=120)
    samples = [gen.sample(), gen.sample()]
    samples[0]["target_fix_action_idx"], samples[1]["target_fix_action_idx"] = 0, 3
    model = _tiny_model(samples)
    t = [x for x, _ in model.tensorize_dataset(iter(copy.deepcopy(samples)), parallelize=False)]
    mb = model.initialize_minibatch()
    for x in t:
        model.extend_minibatch_with(x, mb)
    out = model.finalize_minibatch(mb, "cpu")
    g = out["graph_data"]
    n0, n1 = t[0].graph_data.num_nodes, t[1].graph_data.num_nodes
    assert g["node_to_graph_idx"].tolist() == [0] * n0 + [1] * n1
    K = model.gnn_model.num_edge_types
    n_fwd = (K - 1) // 2
    for k in range(K):
        src, tgt = g["adjacency_lists"][k]
        e0 = t[0].graph_data.adjacency_lists[k][0].shape[0]
        np.testing.assert_array_equal(src[:e0].numpy(), t[0].graph_data.adjacency_lists[k][0])
        np.testing.assert_array_equal(src[e0:].numpy(), t[1].graph_data.adjacency_lists[k][0] + n0)
        np.testing.assert_array_equal(tgt[e0:].numpy(), t[1].graph_data.adjacency_lists[k][1] + n0)
    # backward edge kinds are the forward ones with columns swapped; the last kind is the self loop (P1)
    for k in range(n_fwd):
        assert torch.equal(g["adjacency_lists"][k][0], g["adjacency_lists"][n_fwd + k][1])
        assert torch.equal(g["adjacency_lists"][k][1], g["adjacency_lists"][n_fwd + k][0])
    assert torch.equal(g["adjacency_lists"][K - 1][0], torch.arange(n0 + n1, dtype=torch.int32))
    c0 = len(t[0].graph_data.reference_nodes["candidate_nodes"])
    cand = g["reference_node_ids"]["candidate_nodes"]
    np.testing.assert_array_equal(cand[c0:].numpy(), t[1].graph_data.reference_nodes["candidate_nodes"] + n0)
    assert out["correct_candidate_node_idxs"].tolist() == [t[0].target_location_node_idx, t[1].target_location_node_idx + c0]
    assert out["has_bug"].tolist() == [True, True]
    r0 = len(t[0].text_rewrite_original_idx) + len(t[0].candidate_rewrite_original_idx) + len(t[0].pair_rewrite_original_idx)
    assert out["rewrite_to_graph_id"].tolist().count(0) == r0
    groups0 = t[0].num_rewrite_locations_considered
    second_groups = (t[1].target_rewrite_to_location_group + t[1].candidate_symbol_to_varmisused_node + t[1].swapped_pair_to_call)
    all_groups = torch.cat((out["rewrite_to_location_group"], out["candidate_symbol_to_location_group"],
                            out["swapped_pair_to_call_location_group"])).tolist()
    assert sorted(x for x in all_groups if x >= groups0) == sorted(gp + groups0 for gp in second_groups)


def test_stop_extending_and_max_nodes():
    from buglab.models.modelregistry import load_model
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    gen = SyntheticBugLabGenerator(seed=6, mean_nodes=100, min_nodes=60, max_nodes=140)
    samples = [gen.sample() for _ in range(6)]
    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 8, "stop_extending_minibatch_after_num_nodes": 250,
                              "max_nodes_per_graph": 10 ** 6, "node_representations": {"min_freq_threshold": 1}}, Path("/tmp/t.pkl.gz"))
    model.compute_metadata(iter(copy.deepcopy(samples)))
    t = list(model.tensorize_dataset(iter(copy.deepcopy(samples)), parallelize=False))
    sizes = [mb["graph_data"]["num_graphs"] for mb, _ in model.minibatch_iterator(iter(t), "cpu", 100, parallelize=False)]
    assert sum(sizes) == 6 and max(sizes) <= 3  # a batch closes once it holds >= 250 nodes (F8 knob)
    model.gnn_model.max_nodes_per_graph = 10
    assert model.tensorize(copy.deepcopy(samples[0])) is None  # dropped, like the reference (gnn.py:404-405)


def test_evaluate_metrics_logic():
    from buglab.models.evaluate import evaluate_predictions

    point = {"target_fix_action_idx": 1, "graph": {"reference_nodes": [4, 4, 9]},
             "candidate_rewrite_metadata": [("A", None), ("A", None), ("B", None)]}
    clean = {"target_fix_action_idx": None, "graph": {"reference_nodes": [4, 9]}, "candidate_rewrite_metadata": [("A", None), ("B", None)]}
    preds = [
        (point, {4: -0.1, 9: -3.0, -1: -4.0}, [-2.0, -0.2, -0.5]),   # location right, rewrite right
        (point, {4: -3.0, 9: -0.1, -1: -4.0}, [-2.0, -0.2, -0.5]),   # location wrong, repair given location right
        (clean, {4: -3.0, 9: -3.0, -1: -0.1}, [-1.0, -1.0]),          # correctly silent
    ]
    m = evaluate_predictions(preds)
    assert m["num_samples"] == 3 and abs(m["localization_accuracy"] - 2 / 3) < 1e-12
    assert m["repair_accuracy_given_location"] == 1.0 and m["localization_and_repair_accuracy"] == 0.5
    assert m["no_bug_recall"] == 1.0 and m["bug_detection_rate"] == 1.0


from typing import Any, Dict  # noqa: E402

from ptgnn.baseneuralmodel import AbstractNeuralModel, ModelTrainer, ModuleWithMetrics  # noqa: E402


class _ToyNet(ModuleWithMetrics):
    def __init__(self):
        super().__init__()
        self.param = torch.nn.Parameter(torch.tensor(0.0))

    def _reset_module_metrics(self):
        self.seen = 0

    def _module_metrics(self):
        return {"seen": self.seen}

    def forward(self, data):
        self.seen += int(data.shape[0])
        return ((self.param - data) ** 2).mean()


class _ToyModel(AbstractNeuralModel[float, float, _ToyNet]):
    def update_metadata_from(self, datapoint):
        self.count = getattr(self, "count", 0) + 1

    def build_neural_module(self):
        return _ToyNet()

    def tensorize(self, datapoint):
        return None if datapoint < 0 else datapoint

    def initialize_minibatch(self) -> Dict[str, Any]:
        return {"data": []}

    def extend_minibatch_with(self, t, mb):
        mb["data"].append(t)
        return True

    def finalize_minibatch(self, mb, device):
        return {"data": torch.tensor(mb["data"], device=device)}


def test_trainer_loop_hooks_and_checkpoint(tmp_path):
    """ModelTrainer on CPU with a 1-parameter model (the reference's own MockNeuralModel pattern, tests/test_modelsync.py:15-45)."""
    data = [3.0] * 16 + [-1.0]  # the negative sample is dropped by tensorize
    model = _ToyModel()
    path = tmp_path / "toy.pkl.gz"
    trainer = ModelTrainer(model, path, max_num_epochs=30, minibatch_size=4,
                           optimizer_creator=lambda p: torch.optim.SGD(p, lr=0.3), clip_gradient_norm=100.0)
    epochs = []
    trainer.register_train_epoch_end_hook(lambda m, nn, e, metrics: epochs.append((e, metrics["seen"])))
    trainer.train(data, data, show_progress_bar=False, parallelize=False, patience=3, device="cpu")
    assert model.count == 17 and epochs[0] == (0, 16)
    assert abs(float(trainer.neural_module.param.detach()) - 3.0) < 0.05
    assert path.exists()
    model2, nn2 = _ToyModel.restore_model(path, "cpu")
    assert isinstance(nn2, _ToyNet) and model2.count == 17


def test_modules_are_picklable_with_stable_paths(tmp_path):
    """Checkpoints are pickles of (model, nn) (reference modelsync/server.py:34): no ctypes handle may live on a module."""
    import io

    from buglab_b200.synthetic import SyntheticBugLabGenerator

    gen = SyntheticBugLabGenerator(seed=7, mean_nodes=60, min_nodes=40, max_nodes=90)
    samples = [gen.sample() for _ in range(3)]
    model = _tiny_model(samples)
    nn = model.build_neural_module()
    buf = io.BytesIO()
    torch.save((model, nn), buf)
    buf.seek(0)
    model2, nn2 = torch.load(buf, weights_only=False)
    assert type(model2).__module__ == "buglab.models.gnn" and type(nn2).__module__ == "buglab.models.gnn"
    assert nn2.state_dict().keys() == nn.state_dict().keys()
    assert model2.gnn_model.edge_types == model.gnn_model.edge_types
    model.save(tmp_path / "m.pkl.gz", nn)
    model3, nn3 = type(model).restore_model(tmp_path / "m.pkl.gz", "cpu")
    for (k, a), (_, b) in zip(nn.state_dict().items(), nn3.state_dict().items()):
        assert torch.equal(a, b), k


def test_prefetcher_stops_its_producer_when_the_consumer_leaves_early():
    """An abandoned epoch (exception, or a data-parallel rank that ran out of data first) must not leave a producer
    thread blocked on a full queue."""
    import threading
    import time

    import torch
    from ptgnn.baseneuralmodel.trainer import _Prefetcher

    produced = []

    def endless():
        i = 0
        while True:
            produced.append(i)
            yield i
            i += 1

    before = threading.active_count()
    pf = _Prefetcher(endless, torch.device("cpu"), depth=2)
    it = iter(pf)
    assert [next(it), next(it), next(it)] == [0, 1, 2]
    it.close()                      # what happens when the training loop's generator chain is dropped
    assert not pf._thread.is_alive() and threading.active_count() == before
    n = len(produced)
    time.sleep(0.2)
    assert len(produced) == n       # nothing is produced after the close

    # normal exhaustion and error propagation still work
    assert list(_Prefetcher(lambda: iter(range(5)), torch.device("cpu"))) == [0, 1, 2, 3, 4]

    def failing():
        yield 1
        raise ValueError("boom")

    import pytest
    with pytest.raises(ValueError):
        list(_Prefetcher(failing, torch.device("cpu")))


def test_tensorize_dataset_background_producer_matches_sequential():
    """``tensorize_dataset(parallelize=True)``: one producer thread, same pairs in the same order as the sequential loop,
    dropped samples dropped, ``return_input_data`` honoured, a failing ``tensorize`` surfaces in the consumer, and a
    consumer that stops early leaves no thread behind."""
    import threading

    import pytest

    from buglab_b200.synthetic import SyntheticBugLabGenerator

    gen = SyntheticBugLabGenerator(seed=17, mean_nodes=70, min_nodes=30, max_nodes=110)
    samples = [gen.sample() for _ in range(40)]
    model = _tiny_model(samples)
    model.gnn_model.max_nodes_per_graph = sorted(len(s["graph"]["nodes"]) for s in samples)[-4]   # a few samples are dropped
    sequential = list(model.tensorize_dataset(iter(copy.deepcopy(samples)), return_input_data=True, parallelize=False))
    assert 30 <= len(sequential) < 40
    before = threading.active_count()
    background = list(model.tensorize_dataset(iter(copy.deepcopy(samples)), return_input_data=True, parallelize=True))
    assert len(background) == len(sequential)
    for (t_a, dp_a), (t_b, dp_b) in zip(sequential, background):
        assert dp_a == dp_b and dp_a is not None
        assert t_a.graph_data.num_nodes == t_b.graph_data.num_nodes
        for (s_a, g_a), (s_b, g_b) in zip(t_a.graph_data.adjacency_lists, t_b.graph_data.adjacency_lists):
            assert np.array_equal(s_a, s_b) and np.array_equal(g_a, g_b)
    assert all(dp is None for _, dp in model.tensorize_dataset(iter(copy.deepcopy(samples[:5])), parallelize=True))

    it = model.tensorize_dataset(iter(copy.deepcopy(samples)), parallelize=True)
    next(it), next(it)
    it.close()                                                   # early stop: the producer thread is stopped and joined
    assert threading.active_count() <= before

    broken = copy.deepcopy(samples[:6])
    del broken[3]["graph"]["reference_nodes"]
    with pytest.raises(KeyError):
        list(model.tensorize_dataset(iter(broken), parallelize=True))
    assert threading.active_count() <= before


def test_buffered_shuffle_is_a_permutation_and_streams():
    import random

    from ptgnn.baseneuralmodel.trainer import _buffered_shuffle

    pulled = []

    def source():
        for i in range(1000):
            pulled.append(i)
            yield i

    out = []
    for x in _buffered_shuffle(source(), 64, random.Random(0)):
        if not out:
            assert len(pulled) == 64          # the first sample leaves as soon as one buffer is full, not at the end
        out.append(x)
    assert sorted(out) == list(range(1000)) and out != list(range(1000))
    assert max(abs(pos - x) for pos, x in enumerate(out)) < 600      # local mixing: nothing travels arbitrarily far ...
    assert sum(1 for pos, x in enumerate(out) if abs(pos - x) > 16) > 500   # ... but most samples do move
    assert list(_buffered_shuffle(iter([]), 8)) == [] and sorted(_buffered_shuffle(iter([3, 1, 2]), 8)) == [1, 2, 3]


def test_record_stream_traversal_reaches_every_plan_tensor():
    """the device plan is a NamedTuple whose first fields are ints; every tensor field of it (and of
    the adjacency carrying it) must be visited so the caching allocator learns about the cross-stream use."""
    import torch
    from buglab_b200.ops import EdgePlan
    from ptgnn.baseneuralmodel.trainer import _reachable_cuda_tensors
    from ptgnn.neuralmodels.gnn.messagepassing.abstractmessagepassing import PlannedAdjacency

    t = lambda: torch.zeros(3, dtype=torch.int32)  # noqa: E731
    fields = {}
    for name, kind in EdgePlan.__annotations__.items():
        fields[name] = t() if kind is torch.Tensor else ((0, 1) if "host" in name else 5)
    plan = EdgePlan(**fields)
    adjacency = PlannedAdjacency([(t(), t()), (t(), t())])
    adjacency.plan = plan
    minibatch = {"graph_data": {"adjacency_lists": adjacency, "node_data": {"token_idxs": t(), "lengths": t()},
                                "reference_node_ids": {"a": t()}, "num_graphs": 4, "counts": [1, 2, 3]},
                 "correct_idxs": t(), "names": ["x", "y"], "nested": [(t(), None, 3)]}
    found = _reachable_cuda_tensors(minibatch, [], any_device=True)
    ids = {id(x) for x in found}
    n_plan_tensors = sum(1 for v in plan if isinstance(v, torch.Tensor))
    assert n_plan_tensors >= 14
    for v in plan:
        if isinstance(v, torch.Tensor):
            assert id(v) in ids
    assert len(found) == n_plan_tensors + 4 + 3 + 1 + 1


def test_identifier_splitting_follows_the_unicode_state_machine():
    """upstream dpu_utils classifies characters with str.isupper/isdigit/isalnum.  The ASCII regex fast path must
    equal that state machine on ASCII input, and non-ASCII letters must not be treated as separators."""
    import random

    from dpu_utils.codeutils.identifiersplitting import (split_camelcase, split_camelcase_unicode,
                                                         split_identifier_into_parts)

    rng = random.Random(5)
    alphabet = "abXYZq019-+.( $A"
    for _ in range(20000):
        s = "".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 12)))
        assert split_camelcase(s) == split_camelcase_unicode(s), repr(s)
    assert split_identifier_into_parts("naïve") == ["naïve"]
    assert split_identifier_into_parts("Größe") == ["größe"]
    assert split_identifier_into_parts("fooΣigma") == ["foo", "σigma"]
    assert split_identifier_into_parts("x٣y") == ["x", "٣", "y"]          # ARABIC-INDIC DIGIT THREE is a digit
    assert split_identifier_into_parts("a→b") == ["a", "→", "b"]
    assert split_identifier_into_parts("__") == ["__"]
