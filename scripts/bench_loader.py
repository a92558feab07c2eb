#!/usr/bin/env python
"""Host data path throughput: ``.msgpack.l.gz`` shards -> tensorised samples -> packed minibatches.

Times, on the host cores only (no GPU involved):
  host      the reference-shaped Python chain (gzip + msgpack.Unpacker -> dicts -> GnnBugLabModel.tensorize), 1 thread and
            through tensorize_dataset's background producer thread;
  native    libbuglab_shards.so (include/buglab_shards.h) through ShardDataset.tensorized at 1..N threads;
  + pack    the same with minibatch packing (extend_minibatch_with / finalize_minibatch to CPU tensors) on the consumer.
Prints one JSON line; graphs are the c2-shaped synthetic samples bench.py trains on.

    python scripts/bench_loader.py [--shards 8] [--graphs-per-shard 32] [--threads 1,2,4,8]
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
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--graphs-per-shard", type=int, default=32)
    ap.add_argument("--threads", default="1,2,4,8")
    ap.add_argument("--minibatch", type=int, default=256)
    args = ap.parse_args()

    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.shards import ShardDataset
    from buglab_b200.synthetic import SyntheticBugLabGenerator, write_shards
    from dpu_utils.utils import RichPath

    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 256}, Path("/tmp/_bench_loader.pkl.gz"))
    model.compute_metadata(SyntheticBugLabGenerator(seed=12345).samples(64))
    directory = tempfile.mkdtemp(prefix="buglab_loader_")
    paths = write_shards(directory, args.shards, args.graphs_per_shard, seed=1)
    rich = RichPath.create(directory)
    total = args.shards * args.graphs_per_shard
    gz_bytes = sum(os.path.getsize(p) for p in paths)

    def consume(iterator, pack: bool) -> float:
        t0 = time.perf_counter()
        n = 0
        mb = model.initialize_minibatch() if pack else None
        in_mb = 0
        for t, _ in iterator:
            n += 1
            if pack:
                model.extend_minibatch_with(t, mb)
                in_mb += 1
                if in_mb == args.minibatch:
                    model.finalize_minibatch(mb, "cpu")
                    mb, in_mb = model.initialize_minibatch(), 0
        if pack and in_mb:
            model.finalize_minibatch(mb, "cpu")
        dt = time.perf_counter() - t0
        assert n == total, (n, total)
        return total / dt

    results = {}
    for pack in (False, True):
        tag = "+pack" if pack else ""
        results[f"host_1thread{tag}"] = consume(
            model.tensorize_dataset(load_all_msgpack_l_gz(rich), parallelize=False), pack)
        results[f"host_background_thread{tag}"] = consume(
            model.tensorize_dataset(load_all_msgpack_l_gz(rich), parallelize=True), pack)
        for threads in [int(x) for x in args.threads.split(",")]:
            ds = ShardDataset(rich, num_threads=threads)
            consume(ds.tensorized(model), False) if threads == 1 and not pack else None  # warm the library / page cache
            results[f"native_{threads}thread{tag}"] = consume(ds.tensorized(model), pack)
    line = {
        "metric": "code graphs tensorised per second (host data path)", "unit": "graphs/s",
        "cores": os.cpu_count(), "graphs": total, "shards": args.shards, "gz_megabytes": round(gz_bytes / 1e6, 2),
        "mean_nodes_per_graph": 2200, "minibatch": args.minibatch,
        "results": {k: round(v, 1) for k, v in results.items()},
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
