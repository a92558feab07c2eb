#!/usr/bin/env python
"""The metadata pass (vocabulary + edge types, ``model.compute_metadata``) over shard files: raw datapoints through the
host-language chain vs ``ShardDataset.update_model_metadata`` (native decoder, one counter per worker thread).  Host cores
only; also asserts that both passes build the same vocabulary.  Prints one JSON line.

    python scripts/bench_metadata.py [DIRECTORY with *.msgpack.l.gz; default: 4 synthetic 500-graph shards]
"""
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
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.shards import ShardDataset
    from buglab_b200.synthetic import write_shards
    from dpu_utils.utils import RichPath

    if len(sys.argv) > 1:
        directory = sys.argv[1]
    else:
        directory = tempfile.mkdtemp(prefix="buglab_metadata_")
        write_shards(directory, 4, 500, seed=11)
    rich = RichPath.create(directory)
    total = sum(1 for _ in load_all_msgpack_l_gz(rich))

    def fresh_model():
        model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 16}, Path("/tmp/_bench_metadata.pkl.gz"))
        return model

    def timed(model, source) -> float:
        t0 = time.perf_counter()
        model.compute_metadata(source)
        return round(total / (time.perf_counter() - t0), 1)

    results = {}
    host = fresh_model()
    results["host_graphs_per_s"] = timed(host, load_all_msgpack_l_gz(rich))
    vocabulary = host.gnn_model.node_representation_model.vocabulary.token_to_id
    for threads in (1, 2, 4, 8):
        native = fresh_model()
        results[f"native_{threads}thread_graphs_per_s"] = timed(native, ShardDataset(rich, num_threads=threads))
        assert native.gnn_model.node_representation_model.vocabulary.token_to_id == vocabulary
        assert native.gnn_model.edge_types == host.gnn_model.edge_types
    print(json.dumps({"metric": "metadata pass (vocabulary + edge types), graphs/s", "graphs": total,
                      "cores": os.cpu_count(), "results": results}))


if __name__ == "__main__":
    main()
