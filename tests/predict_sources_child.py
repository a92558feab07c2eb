"""Child process of tests/test_shards_cpu.py::test_predict_from_native_shards_equals_the_host_chain: ``model.predict`` on the
host cores (oracle/cpu_backend.py underneath, process-wide patch) over the same shard files through three data sources —
the reference-shaped loader, ShardDataset with unpacked datapoints, ShardDataset with lazy datapoints — must give identical
predictions, datapoints and evaluation metrics.  Prints one JSON line."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def main(directory: str) -> None:
    from pathlib import Path

    import torch

    from oracle import cpu_backend

    cpu_backend.install()
    from buglab.models.evaluate import evaluate_predictions, evaluation_data
    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.shards import LazyDatapoint, ShardDataset
    from buglab_b200.synthetic import SyntheticBugLabGenerator, write_shards
    from dpu_utils.utils import RichPath

    torch.manual_seed(0)
    write_shards(os.path.join(directory, "test"), 3, 40, seed=3, mean_nodes=150, min_nodes=40)   # > CHUNK samples per file
    rich = RichPath.create(os.path.join(directory, "test"))
    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 16}, Path(directory) / "m.pkl.gz")
    model.compute_metadata(SyntheticBugLabGenerator(seed=12345, mean_nodes=150, min_nodes=40).samples(48))
    nn = model.build_neural_module()

    def predictions(source):
        return list(model.predict(source, nn, "cpu", parallelize=False))

    host = predictions(load_all_msgpack_l_gz(rich))
    full = predictions(ShardDataset(rich, num_threads=3))
    lazy = predictions(ShardDataset(rich, num_threads=3, lazy_input_data=True))
    assert len(host) == len(full) == len(lazy) == 120
    still_packed = 0
    for (dp_h, loc_h, rw_h), (dp_f, loc_f, rw_f), (dp_l, loc_l, rw_l) in zip(host, full, lazy):
        assert loc_h == loc_f == loc_l and list(loc_h) == list(loc_f) == list(loc_l)       # same keys, order, values
        assert rw_h == rw_f == rw_l
        assert type(dp_f) is dict and dp_f == dp_h
        assert isinstance(dp_l, LazyDatapoint)
        still_packed += dp_l._full is None                 # predict itself did not need to unpack the graph
        for key in ("target_fix_action_idx", "candidate_rewrites", "candidate_rewrite_metadata"):
            assert dp_l[key] == dp_h[key], key
        assert dp_l["graph"]["reference_nodes"] == dp_h["graph"]["reference_nodes"]
    metrics = [evaluate_predictions(p) for p in (host, full, lazy)]
    assert metrics[0] == metrics[1] == metrics[2]
    still_packed_after_metrics = sum(dp._full is None for dp, _, _ in lazy)
    # everything else is there on demand, and equal to the host loader's objects
    for (dp_h, _, _), (dp_l, _, _) in zip(host[:10], lazy[:10]):
        assert dp_l["graph"]["nodes"] == dp_h["graph"]["nodes"] and dp_l["package_name"] == dp_h["package_name"]
        assert dict(dp_l) == dp_h and dp_l == dp_h and len(dp_l) == len(dp_h) and list(dp_l) == list(dp_h)
        assert dp_l["graph"]["edges"] == dp_h["graph"]["edges"]
    # the evaluate entry point's own data source visits the files in shuffled order: the same samples with the same scores
    # (to rounding: the 50-graph minibatches are composed differently, which moves the last bits of the CPU arithmetic)
    shuffled = predictions(evaluation_data(rich, None, sequential=True))
    identity = lambda dp: (dp["package_name"], tuple(dp["graph"]["reference_nodes"]), dp["target_fix_action_idx"])
    by_identity = {identity(dp): (loc, rw) for dp, loc, rw in host}
    assert len(by_identity) == len(host) == len(shuffled)
    for dp, loc, rw in shuffled:
        loc_h, rw_h = by_identity[identity(dp)]
        assert list(loc) == list(loc_h) and all(abs(loc[k] - loc_h[k]) < 1e-5 for k in loc)
        assert len(rw) == len(rw_h) and all(abs(a - b) < 1e-5 for a, b in zip(rw, rw_h))
    print(json.dumps({"samples": len(host), "lazy_still_packed_after_predict": int(still_packed),
                      "lazy_still_packed_after_metrics": int(still_packed_after_metrics),
                      "localization_accuracy": metrics[0]["localization_accuracy"]}))


if __name__ == "__main__":
    main(sys.argv[1])
