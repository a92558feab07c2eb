#!/usr/bin/env python
"""Host data path of ``predict`` / ``python -m buglab.models.evaluate``: shard files -> (tensorised sample, raw datapoint)
pairs -> minibatches of <= 50 graphs (gnn.py:606-645), on the host cores only (no GPU, no model compute).

  host        the reference-shaped chain: gzip + msgpack.Unpacker -> dicts -> GnnBugLabModel.tensorize (1 thread, and the
              tensorize_dataset background producer thread)
  native      ShardDataset.tensorized(model, return_input_data=True): native decode, datapoints unpacked (and their graphs
              extended as the host chain's tensorisation does) next to every sample
  native-lazy the same with LazyDatapoint views (what evaluate.py uses): nothing is unpacked unless a consumer asks

    python scripts/bench_eval_loader.py [--shards 4] [--graphs-per-shard 100]
"""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shards", type=int, default=4)
    ap.add_argument("--graphs-per-shard", type=int, default=100)
    ap.add_argument("--repeats", type=int, default=3)
    args = ap.parse_args()
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.shards import ShardDataset
    from buglab_b200.synthetic import SyntheticBugLabGenerator, write_shards
    from dpu_utils.utils import RichPath

    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 128}, Path("/tmp/_bench_eval_loader.pkl.gz"))
    model.compute_metadata(SyntheticBugLabGenerator(seed=12345).samples(64))
    directory = tempfile.mkdtemp(prefix="buglab_eval_loader_")
    write_shards(directory, args.shards, args.graphs_per_shard, seed=1)
    rich = RichPath.create(directory)
    total = args.shards * args.graphs_per_shard

    def consume(samples) -> float:
        t0 = time.perf_counter()
        n = 0
        for _mb, datapoints in model.minibatch_iterator(samples, "cpu", max_minibatch_size=50, parallelize=False):
            for dp in datapoints:  # what _iter_per_sample_results / evaluate_predictions read
                n += 1
                dp["graph"]["reference_nodes"], len(dp["candidate_rewrites"]), dp["target_fix_action_idx"]
        assert n == total, (n, total)
        return total / (time.perf_counter() - t0)

    sources = {
        "host_1thread": lambda: model.tensorize_dataset(load_all_msgpack_l_gz(rich), return_input_data=True, parallelize=False),
        "host_background_thread": lambda: model.tensorize_dataset(load_all_msgpack_l_gz(rich), return_input_data=True, parallelize=True),
        "native_1thread": lambda: ShardDataset(rich, num_threads=1).tensorized(model, return_input_data=True),
        "native_4thread": lambda: ShardDataset(rich, num_threads=4).tensorized(model, return_input_data=True),
        "native_lazy_1thread": lambda: ShardDataset(rich, num_threads=1).tensorized(model, True, True),
        "native_lazy_4thread": lambda: ShardDataset(rich, num_threads=4).tensorized(model, True, True),
    }
    results = {}
    with model._tensorize_all_location_rewrites():
        for name, make in sources.items():
            results[name] = round(max(consume(make()) for _ in range(args.repeats)), 1)
    print(json.dumps({"metric": "graphs per second through predict's host data path (tensorise + datapoints + 50-graph minibatches)",
                      "unit": "graphs/s", "cores": os.cpu_count(), "graphs": total, "shards": args.shards,
                      "mean_nodes_per_graph": 2200, "best_of": args.repeats, "results": results}))


if __name__ == "__main__":
    main()
