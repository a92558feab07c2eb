#!/usr/bin/env python
"""End to end FROM DISK: ``.msgpack.l.gz`` shards -> decode -> tensorise -> pack -> H2D -> plan -> train step, on one GPU.

bench.py's ``e2e`` starts from tensorised samples (the decode stage is measured on the host cores by
scripts/bench_loader.py).  This script closes the loop on a GPU box: the c2 workload (bench.py's model and step) fed by
(a) the native shard path (buglab_b200.shards.ShardDataset, default of buglab.models.train) and (b) the reference-shaped
Python chain, both through the trainer's producer thread.  One JSON line: graphs/s for both loaders next to the
device-resident step rate.

    python scripts/bench_e2e_shards.py [--steps 6] [--host-steps 1]
    python scripts/bench_e2e_shards.py --dry-run        # no GPU: loader -> packing to CPU tensors only
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
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--host-steps", type=int, default=1, help="minibatches fed by the Python chain (it is ~10x slower)")
    ap.add_argument("--graphs", type=int, default=256)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--threads", type=int, default=None)
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()

    from pathlib import Path

    import torch

    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.shards import ShardDataset
    from buglab_b200.synthetic import SyntheticBugLabGenerator, write_shards
    from dpu_utils.utils import RichPath
    from ptgnn.baseneuralmodel.trainer import _Prefetcher

    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": args.hidden, "dropout_rate": 0.1,
                              "stop_extending_minibatch_after_num_nodes": 10 ** 9, "max_nodes_per_graph": 10 ** 9},
                             Path("/tmp/buglab_b200_e2e_shards.pkl.gz"))
    model.compute_metadata(SyntheticBugLabGenerator(seed=12345).samples(64))
    work = tempfile.mkdtemp(prefix="buglab_e2e_shards_")
    per_shard = 32
    num_shards = (args.steps + 1) * args.graphs // per_shard
    t0 = time.perf_counter()
    write_shards(os.path.join(work, "train"), num_shards, per_shard, seed=1)
    rich = RichPath.create(os.path.join(work, "train"))
    out = {"graphs_per_step": args.graphs, "hidden": args.hidden, "shards": num_shards,
           "shard_write_seconds": round(time.perf_counter() - t0, 1), "host_cores": os.cpu_count()}

    device = torch.device("cpu") if args.dry_run else torch.device("cuda", 0)

    def minibatches(source, limit_steps):
        def make():
            it = model.minibatch_iterator(source(), device=device, max_minibatch_size=args.graphs,
                                          yield_partial_minibatches=False)
            for i, item in enumerate(it):
                if i >= limit_steps:
                    return
                yield item
        return make

    native = lambda: ShardDataset(rich, num_threads=args.threads).tensorized(model)  # noqa: E731
    host = lambda: model.tensorize_dataset(load_all_msgpack_l_gz(rich), parallelize=True)  # noqa: E731

    if args.dry_run:
        for name, source, steps in (("native", native, args.steps), ("host", host, args.host_steps)):
            t0 = time.perf_counter()
            n = sum(len(raw) for _, raw in _Prefetcher(minibatches(source, steps), device))
            out[f"{name}_loader_graphs_per_s"] = round(n / (time.perf_counter() - t0), 1)
        print(json.dumps(out))
        return

    from buglab.models.utils import LinearWarmupScheduler, optimizer

    torch.manual_seed(0)
    nn = model.build_neural_module().to(device)
    opt = optimizer(nn.parameters())
    opt.max_grad_norm = 0.5
    sched = LinearWarmupScheduler(opt)
    nn.train()

    def train_step(mb):
        opt.zero_grad()
        loss = nn(**mb)
        loss.backward()
        opt.step()
        sched.step(0, 0)
        return loss

    def timed(source, steps, warmup):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        graphs = 0
        for i, (mb, raw) in enumerate(_Prefetcher(minibatches(source, steps + warmup), device)):
            if i == warmup:
                torch.cuda.synchronize(device)
                start.record()
            loss = float(train_step(mb).detach())
            if i >= warmup:
                graphs += len(raw)
        end.record()
        torch.cuda.synchronize(device)
        return graphs / (start.elapsed_time(end) / 1e3), loss

    # device-resident reference point: one packed minibatch trained repeatedly (plan rebuilt each step, as in bench.py)
    first = next(iter(minibatches(native, 1)()))[0]
    for _ in range(3):
        first["graph_data"]["adjacency_lists"].plan = None
        train_step(first)
    torch.cuda.synchronize(device)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for _ in range(args.steps):
        first["graph_data"]["adjacency_lists"].plan = None
        train_step(first)
    end.record()
    torch.cuda.synchronize(device)
    out["resident_graphs_per_s"] = round(args.graphs * args.steps / (start.elapsed_time(end) / 1e3), 1)
    out["from_shards_native_graphs_per_s"], _ = timed(native, args.steps - 1, 1)
    out["from_shards_host_graphs_per_s"], _ = timed(host, args.host_steps, 0)
    out = {k: (round(v, 1) if isinstance(v, float) else v) for k, v in out.items()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
