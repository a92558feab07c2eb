#!/usr/bin/env python
"""Where a SMALL-batch train step spends its time (BASELINE configs[0] regime: hidden 128, ~15 graphs = 30 000 nodes per
minibatch, the reference's own batch size): device time (CUDA events) vs wall time per step on resident minibatches, the
number of C-ABI calls / kernels per step, and a cProfile of the host side (top functions by own time).

    python scripts/profile_small_batch.py [--hidden 128] [--graphs 15] [--steps 30]
"""
import argparse, cProfile, io, json, os, pstats, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--hidden", type=int, default=128)
    ap.add_argument("--graphs", type=int, default=15)
    ap.add_argument("--steps", type=int, default=30)
    args = ap.parse_args()
    import torch
    import bench
    from buglab.models.utils import LinearWarmupScheduler, optimizer
    from buglab_b200 import _lib

    device = torch.device("cuda:0")
    torch.manual_seed(0)
    model, host_batches = bench.make_workload(7, args.graphs, args.hidden, 2000, 4, 0.2)
    nn = model.build_neural_module().to(device)
    opt = optimizer(nn.parameters()); opt.max_grad_norm = 0.5
    sched = LinearWarmupScheduler(opt)
    nn.train()
    resident = [bench.pack(model, tb, device) for tb in host_batches]

    def step(i):
        mb = resident[i % len(resident)]
        mb["graph_data"]["adjacency_lists"].plan = None
        opt.zero_grad()
        loss = nn(**mb)
        loss.backward()
        opt.step()
        sched.step(0, 0)
        return loss

    for i in range(5):
        step(i)
    torch.cuda.synchronize()
    _lib.launch_counter["kernels"] = 0; _lib.launch_counter["calls"] = 0
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); s.record()
    for i in range(args.steps):
        step(i)
    e.record(); t_issue = time.perf_counter() - t0
    torch.cuda.synchronize(); t_wall = time.perf_counter() - t0
    out = {"hidden": args.hidden, "graphs_per_step": args.graphs, "nodes_per_step": int(resident[0]["graph_data"]["node_to_graph_idx"].shape[0]),
           "device_ms_per_step": s.elapsed_time(e) / args.steps, "host_issue_ms_per_step": 1e3 * t_issue / args.steps,
           "wall_ms_per_step": 1e3 * t_wall / args.steps, "c_abi_calls_per_step": _lib.launch_counter["calls"] / args.steps,
           "own_kernels_per_step": _lib.launch_counter["kernels"] / args.steps}
    prof = cProfile.Profile()
    prof.enable()
    for i in range(10):
        step(i)
    torch.cuda.synchronize()
    prof.disable()
    buf = io.StringIO()
    pstats.Stats(prof, stream=buf).sort_stats("tottime").print_stats(28)
    print(buf.getvalue()[:6000])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
