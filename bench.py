#!/usr/bin/env python
"""bench.py — code-graphs/sec of the gnn-mlp train step (fwd + bwd + grad all-reduce + clip + Adam) on B200.

Contract: ``python bench.py --gpus N --steps K --warmup W [--impl reference]`` prints ONE JSON line on rank 0.
Workload = BASELINE.json configs[1]: gnn-mlp, hidden 256, 8 message-passing layers, 256 synthetic code graphs per
step and GPU (~2k nodes each, 8 forward edge kinds -> 17 kinds per layer), fp32, dropout 0.2.
  value      graphs/s with the packed minibatch already resident in HBM (plan build + step inside the timed region)
  e2e        graphs/s through the public API from HOST (tensorised numpy) samples: pack -> pinned -> H2D -> plan ->
             step -> D2H of the loss, every step
  roofline        the DOMINANT kernel of the step (the TMA-fed tcgen05 projection, tensor-bound), timed alone
  roofline_edge   the fused typed-edge message+aggregate kernel (north_star's HBM-bound kernel), timed alone
  roofline_layer  SURVEY §8(d)'s object: one H->H message+aggregate layer forward (split + 2 projections + edge kernel),
                  §8(d) algorithmic bytes / CUDA-event time, at this workload's shape; roofline_c3 = the same at
                  BASELINE configs[2] (1 M nodes / 10 M edges / 14 edge kinds)
  e2e_shards      graphs/s from .msgpack.l.gz files on disk (native decode -> tensorise -> pack -> H2D -> step)
  config1         BASELINE configs[0] (hidden 128, one ~500-graph shard, the reference's 30 000-node minibatch budget)
                  through the kept buglab.models.train entry point: this GPU vs the host cores
  cpu_baseline / --impl reference   the kept train entry point on the host cores with the CPU oracle underneath
                  (oracle/cpu_backend.py), on this arm's model config, minibatches cut by the reference's own
                  30 000-node budget (its real batch size) — the bounded sample
"""
import argparse
import copy
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

HIDDEN = 256
GRAPHS_PER_STEP = 256
MEAN_NODES = 2000
DROPOUT = 0.2
NUM_DISTINCT_BATCHES = 2
CPU_GRAPHS_PER_STEP = 2


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--graphs", type=int, default=GRAPHS_PER_STEP)
    ap.add_argument("--hidden", type=int, default=HIDDEN)
    ap.add_argument("--mean-nodes", type=int, default=MEAN_NODES)
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--skip-extras", action="store_true", help="only value / e2e / clocks (no rooflines, shards, config 1)")
    ap.add_argument("--cpu-threads", type=int, default=0, help="host threads of the CPU arm (0: min(32, affinity))")
    ap.add_argument("--profile", action="store_true",
                    help="profiling run for ncu: one minibatch, resident steps only (numbers printed under a profiler are not bench values)")
    return ap.parse_args()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return json.load(f), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index, self.proc, self.path = gpu_index, None, None

    def __enter__(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
        return self

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()

    def summary(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if not self.path or not os.path.exists(self.path):
            return out
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out["sm_mhz"] = statistics.median(sm)
            out["sm_max_mhz"] = max(mx)
        out["reasons"] = sorted(reasons)
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def make_workload(seed: int, graphs: int, hidden: int, mean_nodes: int, batches: int, dropout: float):
    """Synthetic samples -> metadata -> tensorised (host numpy) samples grouped into `batches` minibatches."""
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    # metadata from a fixed seed so every rank builds the same vocabulary / edge-type layout
    meta_gen = SyntheticBugLabGenerator(seed=12345, mean_nodes=mean_nodes)
    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": dropout,
                              "stop_extending_minibatch_after_num_nodes": 10 ** 9, "max_nodes_per_graph": 10 ** 9},
                             Path("/tmp/buglab_b200_bench.pkl.gz"))
    model.compute_metadata(meta_gen.samples(64))
    gen = SyntheticBugLabGenerator(seed=seed, mean_nodes=mean_nodes)
    host_batches = []
    for _ in range(batches):
        samples = [gen.sample() for _ in range(graphs)]
        host_batches.append([t for t, _ in model.tensorize_dataset(iter(samples), parallelize=False)])
    return model, host_batches


def workload_config(args, model, nodes_per_step, edges_per_step, world):
    """The ``config`` object of the JSON line.  Both arms print the SAME object (the reference arm times a bounded sample
    of this workload and says which in ``cpu_baseline.sample``), so it is built in one place."""
    return {
        "workload": f"gnn-mlp hidden={args.hidden}, 8 MP layers, {args.graphs} graphs/step/GPU (~{args.mean_nodes} nodes each, "
                    f"{model.gnn_model.num_edge_types} edge kinds per layer), train step fwd+bwd+allreduce+clip+Adam, dropout {DROPOUT}",
        "nodes_per_step": int(nodes_per_step), "edges_per_step": int(edges_per_step), "parallelism": f"dp{world}",
        "l2_policy": "inputs larger than L2 (per-layer tables are GBs; 126 MB L2), distinct minibatches cycled",
    }


def pack(model, tensorized, device):
    mb = model.initialize_minibatch()
    for t in tensorized:
        model.extend_minibatch_with(t, mb)
    return model.finalize_minibatch(mb, device)


def run_ours(args):
    import torch

    from buglab.models.utils import LinearWarmupScheduler, optimizer
    from buglab_b200 import _lib, distributed, ops

    local_rank = distributed.init_from_env("nccl")
    rank, world = distributed.rank(), distributed.world_size()
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    torch.manual_seed(0)
    model, host_batches = make_workload(1000 + rank, args.graphs, args.hidden, args.mean_nodes,
                                        1 if args.profile else NUM_DISTINCT_BATCHES, DROPOUT)
    nn = model.build_neural_module().to(device)
    distributed.broadcast_module(nn)
    opt = optimizer(nn.parameters())
    opt.max_grad_norm = 0.5
    sched = LinearWarmupScheduler(opt)
    nn.train()

    reducer = opt.gradient_reducer()  # bucketed all-reduce on a side stream, overlapped with backward (no-op at N=1)

    def train_step(mb):
        opt.zero_grad()
        loss = nn(**mb)
        reducer.begin()
        loss.backward()
        opt.step(grad_scale=reducer.finish())
        sched.step(0, 0)
        return loss

    def barrier():
        if distributed.is_distributed():
            torch.distributed.barrier()
        torch.cuda.synchronize(device)

    # ---------------- device-resident throughput ("value") ----------------
    resident = [pack(model, tb, device) for tb in host_batches]
    nodes = [int(mb["graph_data"]["node_to_graph_idx"].shape[0]) for mb in resident]
    edges = [int(mb["graph_data"]["adjacency_lists"].plan.num_edges) for mb in resident]
    h2d_bytes = [int(mb["graph_data"]["h2d_bytes"]) for mb in resident]
    # device -> host per step: the loss (4 B) and the plan's one small copy (ops.build_edge_plan_from_flat): the two pair counts
    # with node-blocked pair tables, counts + both per-type pointer tables with type-major ones
    K = model.gnn_model.num_edge_types
    plan_d2h_bytes = 8 if resident[0]["graph_data"]["adjacency_lists"].plan.block_nodes > 0 else 4 * (2 + 2 * (K + 1))

    def resident_step(i):
        mb = resident[i % len(resident)]
        mb["graph_data"]["adjacency_lists"].plan = None  # the plan is per-minibatch work: rebuilt inside the step
        return train_step(mb)

    for i in range(args.warmup):
        resident_step(i)
    barrier()
    _lib.launch_counter["kernels"] = 0
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        start.record()
        for i in range(args.steps):
            loss = resident_step(i)
        end.record()
        barrier()
    launches = _lib.launch_counter["kernels"]
    exposed_allreduce_ms = reducer.exposed_ms()
    ms_resident = distributed.all_ranks_max(start.elapsed_time(end), device)
    clock_summary = clocks.summary()
    final_loss = float(loss.detach())

    if args.profile:
        if rank == 0:
            print(json.dumps({"profile_run": True, "ms_per_step": ms_resident / args.steps, "gpu_launches": launches}), flush=True)
        return

    # ---------------- end to end from host samples ("e2e") ----------------
    # The public training path (ptgnn ModelTrainer): a producer thread packs minibatch i+1 (numpy -> pinned -> H2D ->
    # device plan, on a side stream) while minibatch i trains; the loss is read back on the host every step.
    from ptgnn.baseneuralmodel.trainer import _Prefetcher

    def host_batches_iter(n):
        for i in range(n):
            yield pack(model, host_batches[i % len(host_batches)], device)

    # ONE loader pipeline (depth 2) feeds the warm-up and the timed steps, as it does in a training epoch: it is already
    # running when the timer starts and keeps running to its end, so exactly `steps` minibatches are packed and copied
    # inside the timed region (the two consumed first were produced during the warm-up; the two produced last are never
    # consumed).  The time to fill the pipeline from cold is reported separately (`pipeline_fill_ms`): a one-off per epoch
    # that measured anywhere between 30 and 650 ms on different boxes and would otherwise decide a 6-step measurement.
    warm_e2e = max(2, min(3, args.warmup))
    prefetcher = _Prefetcher(lambda: host_batches_iter(warm_e2e + args.steps + 2), device)
    batches = iter(prefetcher)
    t_fill = time.perf_counter()
    first = next(batches)
    pipeline_fill_ms = 1e3 * (time.perf_counter() - t_fill)
    train_step(first)
    del first
    for _ in range(warm_e2e - 1):
        train_step(next(batches))
    barrier()
    start.record()
    # The loss of EVERY step is read back on the host (D2H inside the timed region), one step late: step i's kernels are
    # queued before the read of step i-1's loss blocks, so the host's launch work overlaps the device instead of serialising
    # with it (a blocking read right after each step costs ~20 ms of idle device per step at this size).
    pending, losses_host = None, []
    for _ in range(args.steps):
        loss_dev = train_step(next(batches)).detach()
        if pending is not None:
            losses_host.append(float(pending))
        pending = loss_dev
    losses_host.append(float(pending))
    end.record()
    barrier()
    batches.close()  # stops the producer and drops the two surplus minibatches
    assert len(losses_host) == args.steps
    ms_e2e = distributed.all_ranks_max(start.elapsed_time(end), device)

    total_graphs = args.graphs * world * args.steps
    result = {
        "metric": "code-graphs/sec (train step, device-timed)",
        "value": total_graphs / (ms_resident / 1e3),
        "unit": "graphs/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_resident / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, model, nodes[0], edges[0], world),
        "e2e": {"value": total_graphs / (ms_e2e / 1e3), "unit": "graphs/s", "h2d_bytes_per_step": h2d_bytes[0],
                "d2h_bytes_per_step": 4 + plan_d2h_bytes, "ms_per_step": ms_e2e / args.steps,
                "from": "host tensorised samples (numpy) -> minibatch packing -> pinned staging -> H2D -> device plan (producer thread, side stream, overlapped with the previous step as in ModelTrainer) -> step -> loss D2H (every step, read one step late); steady state: the depth-2 loader pipeline runs through warm-up and timed steps, `steps` minibatches are packed and copied inside the timed region",
                "pipeline_fill_ms": pipeline_fill_ms},
        "gpu_launches": launches,
        "clocks": clock_summary,
        "final_loss": final_loss,
        "allreduce": {"buckets": reducer.num_buckets, "bytes": int(opt.flat_grad.numel()) * 4,
                      "exposed_ms_last_step": exposed_allreduce_ms,
                      "how": "per-bucket ncclAllReduce on a side stream as backward completes each bucket; exposed = compute-stream "
                             "wait after the last backward kernel"},
    }

    if rank == 0:
        if not args.skip_extras:
            del resident
            torch.cuda.empty_cache()
            result.update(rooflines(model, host_batches[0], args.hidden, device))
            if world == 1:
                result["e2e_shards"] = e2e_from_shards(args, device, result["value"])
                result["config1"] = config1_gpu(device)
                result["config4_seq"] = config4_seq()
        if world == 1 and not args.skip_cpu_baseline:
            result["cpu_baseline"] = cpu_arm_subprocess(args.hidden, steps=2, warmup=1, threads=args.cpu_threads)
            if "config1" in result:
                result["config1"]["cpu"] = cpu_arm_subprocess(CONFIG1_HIDDEN, steps=10, warmup=2, threads=args.cpu_threads)
        print(json.dumps(result), flush=True)
    if distributed.is_distributed():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


# ---------------------------------------------------------------------------------------------------------------
# rooflines
# ---------------------------------------------------------------------------------------------------------------
def ncu_traffic(kernel: str, key: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ``ncu --set full`` captures
    (profiles/ncu_traffic.json, keyed by kernel and shape); None when that shape was never captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as f:
            return json.load(f).get(kernel, {}).get(key)
    except (OSError, ValueError):
        return None


def _time_on_stream(fn, device, reps=10, warm=3):
    import torch

    stream = torch.cuda.current_stream(device)
    for _ in range(warm):
        fn()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(device)
    s.record(stream)
    for _ in range(reps):
        fn()
    e.record(stream)
    e.synchronize()
    return s.elapsed_time(e) / reps


def layer_rooflines(plan, hidden, device, label, seed=0):
    """One H->H typed-edge message+aggregate layer forward at the shape of ``plan``: the whole layer (SURVEY §8(d) bytes
    over the CUDA-event time of split + both projections + edge kernel), the edge kernel alone and one projection alone."""
    import torch

    from buglab_b200 import _lib, ops

    peaks, peak_kind = measured_peaks()
    N, E, K, D, M = plan.num_nodes, plan.num_edges, plan.num_edge_types, hidden, hidden
    lib = _lib.load()
    g = torch.Generator(device=device).manual_seed(seed)
    h = torch.randn(N, D, device=device, generator=g)
    weight = torch.randn(K, M, 2 * D, device=device, generator=g) / (2 * D) ** 0.5
    bias = torch.randn(K, M, device=device, generator=g) * 0.1
    shape_key = f"{N}x{E}x{M}"
    algo_bytes = E * (2 * D * 4 + 12) + N * M * 4            # SURVEY.md §8(d), reference formulation
    out = {}

    # (1) the whole layer forward
    def layer():
        with torch.no_grad():
            return ops.typed_edge_message_max(h, weight, bias, plan)

    ms_layer = _time_on_stream(layer, device, reps=6, warm=2)
    flops_layer = 2.0 * (plan.num_s_pairs + plan.num_t_pairs) * D * M * 3   # split-fp16: three MMAs per product
    out["layer"] = {"bound": "hbm", "what": f"{label}: H->H message+aggregate layer forward (node split + U and V projections + "
                    "fused edge kernel)", "achieved": algo_bytes / ms_layer / 1e6, "peak": peaks["hbm_gbs"], "peak_kind": peak_kind,
                    "unit": "GB/s", "frac": algo_bytes / ms_layer / 1e6 / peaks["hbm_gbs"], "traffic": None,
                    "algorithmic_bytes": algo_bytes, "ms": ms_layer, "nodes": N, "edges": E, "edge_kinds": K,
                    "s_pairs": plan.num_s_pairs, "t_pairs": plan.num_t_pairs,
                    "tensor_view": {"tflop_issued": flops_layer / 1e12, "achieved_tflops": flops_layer / ms_layer / 1e9,
                                    "peak_tflops": peaks["bf16_tflops"],
                                    "frac": flops_layer / ms_layer / 1e9 / peaks["bf16_tflops"],
                                    "note": "fp32-exact projections need 3 fp16 MMAs per product: at the tensor peak alone they take "
                                            f"{flops_layer / peaks['bf16_tflops'] / 1e9:.2f} ms, the HBM floor of the §8(d) bytes is "
                                            f"{algo_bytes / peaks['hbm_gbs'] / 1e6:.2f} ms"}}

    # (2) the fused edge kernel alone
    u = torch.randn(plan.num_s_pairs, M, device=device, generator=g)
    v = torch.randn(plan.num_t_pairs, M, device=device, generator=g)
    agg = torch.empty(N, M, device=device); xwin = torch.empty_like(agg)
    ewin = torch.empty(N, M, device=device, dtype=torch.int32)
    stream = torch.cuda.current_stream(device)

    def edge():
        _lib.check(lib.bl_edge_segmax_fwd(u.data_ptr(), v.data_ptr(), plan.row_ptr.data_ptr(), plan.urow.data_ptr(),
                                          plan.vrow.data_ptr(), N, M, agg.data_ptr(), xwin.data_ptr(), ewin.data_ptr(),
                                          stream.cuda_stream), "bl_edge_segmax_fwd")

    ms_edge = _time_on_stream(edge, device)
    out["edge"] = {"bound": "hbm", "kernel": "edge_segmax_fwd_warp (bl_edge_segmax_fwd), H->H layer",
                   "achieved": algo_bytes / ms_edge / 1e6, "peak": peaks["hbm_gbs"], "peak_kind": peak_kind, "unit": "GB/s",
                   "frac": algo_bytes / ms_edge / 1e6 / peaks["hbm_gbs"], "traffic": ncu_traffic("edge_segmax_fwd_warp", shape_key),
                   "algorithmic_bytes": algo_bytes, "ms_per_launch": ms_edge, "nodes": N, "edges": E,
                   "working_set_bytes": int((u.numel() + v.numel()) * 4)}
    del u, v, agg, xwin, ewin

    # (3) the dominant kernel: one projection (U table) on the TMA-fed tcgen05 kernel
    if ops.USE_TMA and plan.s_tiles is not None and lib.bl_tma_gemm_supported(M, D):
        h_split = ops.rows_split(h)
        parts = ops.weight_parts(weight, M, D, 0, False)
        P = plan.num_s_pairs

        def proj():
            return ops.tma_project(h_split, plan.s_node, parts, None, None, plan.s_tiles, P, plan.s_slabs)

        ms_proj = _time_on_stream(proj, device)
        flops = 2.0 * P * D * M * 3
        out["projection"] = {"bound": "tensor", "kernel": "tg::proj_kernel<256, 2, gather> (bl_tma_project), U table of an H->H layer",
                             "achieved": flops / ms_proj / 1e9, "peak": peaks["bf16_tflops"], "peak_kind": peak_kind + " (cuBLAS bf16 burst)",
                             "unit": "TFLOP/s", "frac": flops / ms_proj / 1e9 / peaks["bf16_tflops"],
                             "traffic": ncu_traffic("proj_kernel", f"{P}x{D}x{M}"), "flops_per_launch": flops,
                             "flops_note": "2*P*D*M*3: hi.hi + hi.lo + lo.hi MMAs actually issued (fp32-equivalent work is a third)",
                             "ms_per_launch": ms_proj, "pair_rows": P,
                             "hbm_view": {"algorithmic_bytes": P * (D * 4 + 4) + P * M * 4,
                                          "achieved_GBs": (P * (D * 4 + 4) + P * M * 4) / ms_proj / 1e6}}
    return out


def rooflines(model, tensorized_batch, hidden, device):
    import torch

    from buglab_b200 import ops
    from buglab_b200.synthetic import packed_edge_batch

    mb = pack(model, tensorized_batch, device)
    graph = mb["graph_data"]
    block_nodes = ops.plan_block_nodes_for([(hidden, hidden), (2 * hidden, 2 * hidden)])  # the layout the model's plan uses
    plan = ops.build_edge_plan(graph["adjacency_lists"], int(graph["node_to_graph_idx"].shape[0]), block_nodes)
    del mb, graph
    here = layer_rooflines(plan, hidden, device, "this workload (configs[1])")
    del plan
    torch.cuda.empty_cache()
    result = {"roofline": here.get("projection", here["edge"]), "roofline_edge": here["edge"], "roofline_layer": here["layer"]}
    # BASELINE configs[2]: 1 M nodes / 10 M edges / 14 edge kinds, directly synthesised packed batch
    src, tgt, etype = packed_edge_batch(1_000_000, 10_000_000, 14, 0)
    plan = ops.build_edge_plan_from_flat(*(torch.from_numpy(a).to(device) for a in (src, tgt, etype)), 1_000_000, 14,
                                         ops.plan_block_nodes_for([(256, 256)]))
    c3 = layer_rooflines(plan, 256, device, "configs[2] (1M nodes / 10M edges / 14 kinds)")
    result["roofline_c3"] = {"layer": c3["layer"], "edge": c3["edge"], "projection": c3.get("projection")}
    del plan
    torch.cuda.empty_cache()
    return result


# ---------------------------------------------------------------------------------------------------------------
# end to end from shard files; BASELINE configs[0] through the train entry point
# ---------------------------------------------------------------------------------------------------------------
def e2e_from_shards(args, device, resident_value):
    """The c2 train step fed from ``.msgpack.l.gz`` files: native decode + tensorise (include/buglab_shards.h) -> packing ->
    pinned staging -> H2D -> device plan (producer thread) -> step -> loss D2H, through ModelTrainer's own prefetcher."""
    import torch
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab.models.utils import LinearWarmupScheduler, optimizer
    from buglab_b200.shards import ShardDataset
    from buglab_b200.synthetic import SyntheticBugLabGenerator, write_shards
    from dpu_utils.utils import RichPath
    from ptgnn.baseneuralmodel.trainer import _Prefetcher

    steps, warm = max(2, min(args.steps, 4)), 1
    work = tempfile.mkdtemp(prefix="buglab_bench_shards_")
    per_shard = 32
    t0 = time.perf_counter()
    write_shards(os.path.join(work, "train"), (steps + warm) * args.graphs // per_shard, per_shard, seed=1, mean_nodes=args.mean_nodes)
    write_s = time.perf_counter() - t0
    rich = RichPath.create(os.path.join(work, "train"))
    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": args.hidden, "dropout_rate": DROPOUT,
                              "stop_extending_minibatch_after_num_nodes": 10 ** 9, "max_nodes_per_graph": 10 ** 9},
                             Path("/tmp/buglab_b200_bench_shards.pkl.gz"))
    model.compute_metadata(SyntheticBugLabGenerator(seed=12345, mean_nodes=args.mean_nodes).samples(64))
    torch.manual_seed(0)
    nn = model.build_neural_module().to(device)
    opt = optimizer(nn.parameters())
    opt.max_grad_norm = 0.5
    sched = LinearWarmupScheduler(opt)
    nn.train()

    def make():
        it = model.minibatch_iterator(ShardDataset(rich).tensorized(model), device=device, max_minibatch_size=args.graphs,
                                      yield_partial_minibatches=False)
        for i, item in enumerate(it):
            if i >= steps + warm:
                return
            yield item

    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    graphs, pending = 0, None
    for i, (mb, raw) in enumerate(_Prefetcher(make, device)):
        if i == warm:
            torch.cuda.synchronize(device)
            start.record()
        opt.zero_grad()
        loss = nn(**mb)
        loss.backward()
        opt.step()
        sched.step(0, 0)
        if pending is not None:
            float(pending)  # every step's loss is read on the host, one step late (see run_ours)
        pending = loss.detach()
        if i >= warm:
            graphs += len(raw)
    float(pending)
    end.record()
    torch.cuda.synchronize(device)
    value = graphs / (start.elapsed_time(end) / 1e3)
    del nn, opt
    torch.cuda.empty_cache()
    return {"value": value, "unit": "graphs/s", "steps": steps, "fraction_of_resident_value": value / resident_value,
            "from": "gzip+msgpack shard files -> native decode/tensorise (libbuglab_shards.so, worker threads) -> packing -> pinned "
                    "staging -> H2D -> device plan -> step -> loss D2H", "shard_write_seconds": round(write_s, 1),
            "host_cores": os.cpu_count()}


CONFIG1_HIDDEN = 128
CONFIG1_GRAPHS = 500


def _config1_arguments(work, hidden, epochs, epoch_samples, sequential=True):
    """docopt-style arguments of ``python -m buglab.models.train gnn-mlp TRAIN VALID MODEL`` for BASELINE configs[0]:
    one shard of ~500 graphs (~2 000 nodes each), hidden 128 (or this arm's width for the reference arm), reference
    defaults otherwise (``--minibatch-size 300``: the 30 000-node budget of modelregistry.py:53-54 ends the minibatches)."""
    from buglab_b200.synthetic import write_shards

    if not os.path.exists(os.path.join(work, "train")):
        write_shards(os.path.join(work, "train"), 1, CONFIG1_GRAPHS, seed=11, mean_nodes=MEAN_NODES)
        write_shards(os.path.join(work, "valid"), 1, 16, seed=12, mean_nodes=MEAN_NODES)
    return {"MODEL_NAME": "gnn-mlp", "TRAIN_DATA_PATH": os.path.join(work, "train"), "VALID_DATA_PATH": os.path.join(work, "valid"),
            "MODEL_FILENAME": os.path.join(work, "model.pkl.gz"), "--aml": False, "--azure-info": None, "--amp": False,
            "--sequential": bool(sequential), "--quiet": True, "--debug": False, "--host-loader": False, "--restore-path": None,
            "--max-num-epochs": str(epochs), "--minibatch-size": "300", "--validate-after": str(epoch_samples),
            "--limit-num-elements": None, "--max-files-per-fold": None,
            "--model-spec": json.dumps({"hidden_state_size": hidden, "dropout_rate": DROPOUT})}


def _train_entry_rate(arguments):
    """Runs the kept train entry point and returns (graphs/s, steps, graphs) of its LAST training epoch (the epochs
    before it are the warm-up), as ModelTrainer itself measures them."""
    from buglab.models import train

    trainer = train.run(arguments)
    stats = trainer.last_epoch_stats
    return stats["samples_per_second"], stats["train_steps"], stats["train_samples"], stats["train_seconds"]


def config1_gpu(device):
    import torch

    work = tempfile.mkdtemp(prefix="buglab_bench_config1_")
    # ~15 graphs (30 000 nodes) per step.  An epoch is the whole 500-graph shard (~34 steps): the entry point re-opens and
    # inflates the shard at every epoch start, which a 10-step epoch would charge to 150 graphs.  Epoch 1 warms up, epoch 2
    # is the timed one.
    rate, steps, graphs, seconds = _train_entry_rate(_config1_arguments(work, CONFIG1_HIDDEN, epochs=2, epoch_samples=CONFIG1_GRAPHS,
                                                                        sequential=False))
    torch.cuda.synchronize(device)
    torch.cuda.empty_cache()
    return {"workload": "BASELINE configs[0]: gnn-mlp hidden=128, one shard of 500 synthetic graphs (~2 000 nodes each), "
                        "python -m buglab.models.train defaults (minibatches cut at 30 000 nodes; data loading in the background "
                        "on the GPU, --sequential on the CPU arm)",
            "gpu": {"value": rate, "unit": "graphs/s", "steps": steps, "graphs": graphs, "ms_per_step": 1e3 * seconds / max(steps, 1),
                    "includes": "shard decode + tensorise + packing + H2D + plan + step, as the entry point runs them"}}


# ---------------------------------------------------------------------------------------------------------------
# CPU arm
# ---------------------------------------------------------------------------------------------------------------
def config4_seq():
    """BASELINE.json configs[3] (seq-great, hidden 512, 8 heads, 5 layers, 64 sequences of <= 512 tokens): the train step
    of the sequence model, measured by scripts/bench_seq.py in a child process (its own CUDA context, after this one has
    released its cache).  Context only: the headline metric stays the gnn-mlp step."""
    import subprocess

    cmd = [sys.executable, os.path.join(ROOT, "scripts", "bench_seq.py"), "--steps", "8", "--warmup", "3"]
    try:
        proc = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
        lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
        if proc.returncode != 0 or not lines:
            return {"unavailable": (proc.stderr or proc.stdout)[-300:]}
        out = json.loads(lines[-1])
        out["command"] = "python scripts/bench_seq.py --steps 8 --warmup 3"
        return out
    except Exception as exc:  # a context leg must never take the headline line down with it
        return {"unavailable": repr(exc)[:300]}


def cpu_thread_count(requested: int = 0) -> int:
    """Fixed thread count of the CPU arm: min(32, usable CPUs).  (os.cpu_count() over-reports inside CPU-limited
    containers: round 1's box reported 128 and ran the oracle 30x slower with 128 threads than with 32.)"""
    if requested > 0:
        return requested
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    return max(1, min(32, avail))


def cpu_arm_subprocess(hidden, steps, warmup, threads):
    """Runs the CPU arm in a fresh process (its oracle/cpu_backend.py patches are process-wide) and returns its
    ``cpu_baseline`` object."""
    cmd = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--hidden", str(hidden), "--steps", str(steps),
           "--warmup", str(warmup), "--cpu-threads", str(threads)]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1", CUDA_VISIBLE_DEVICES="")
    for k in ("LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    proc = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    for line in reversed(proc.stdout.strip().splitlines()):
        if line.startswith("{"):
            return json.loads(line)["cpu_baseline"]
    return {"error": (proc.stderr or proc.stdout)[-400:]}


def cpu_train_entry(hidden, steps, warmup, threads):
    """``buglab.models.train`` on the host cores with the CPU oracle underneath (oracle/cpu_backend.py): warm-up epoch(s)
    then one timed epoch of ``steps`` minibatches, each cut by the reference's 30 000-node budget (~15 graphs)."""
    import torch

    from oracle import cpu_backend

    cores = cpu_thread_count(threads)
    torch.set_num_threads(cores)
    cpu_backend.install()
    work = tempfile.mkdtemp(prefix="buglab_bench_cpu_")
    graphs_per_step = 15  # 30 000 nodes / ~2 000 nodes per graph; the entry point counts epochs in samples
    epochs = 2 if warmup > 0 else 1
    arguments = _config1_arguments(work, hidden, epochs=epochs, epoch_samples=graphs_per_step * steps)
    if warmup > 0 and warmup != steps:
        # a shorter warm-up epoch: run it as its own call (epoch lengths are per call), then the timed call
        _train_entry_rate(_config1_arguments(work, hidden, epochs=1, epoch_samples=graphs_per_step * warmup))
        arguments = _config1_arguments(work, hidden, epochs=1, epoch_samples=graphs_per_step * steps)
    rate, nsteps, graphs, seconds = _train_entry_rate(arguments)
    return {"value": rate, "unit": "graphs/s", "cores": cores, "kind": "port",
            "sample": f"{nsteps} minibatches ({graphs} graphs, <= 30 000 nodes each: the reference's own minibatch budget) of the "
                      f"gnn-mlp hidden={hidden} train step through python -m buglab.models.train --sequential on {cores} host "
                      f"threads, after {warmup} warm-up minibatches; arithmetic = oracle/ (PyTorch CPU restatement of ptgnn / "
                      "torch_scatter, the same substitution that generates tests/golden)",
            "seconds": seconds, "ms_per_step": 1e3 * seconds / max(nsteps, 1)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # the CPU arm must not see a GPU: the kept entry point trains on cuda:0 whenever one is visible (reference
    # modelregistry.py:155 / train.py), while the oracle underneath is a host implementation.  torch is not imported yet.
    os.environ["CUDA_VISIBLE_DEVICES"] = ""
    for key in ("WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):  # rank 0 runs alone; no process group
        os.environ.pop(key, None)
    os.environ["RANK"] = "0"
    # the config object is this arm's workload exactly as the GPU arm prints it (rank 0's first minibatch: same generator,
    # same seed; node and edge counts taken from the tensorised samples, edges incl. backward and self edges)
    model, host_batches = make_workload(1000, args.graphs, args.hidden, args.mean_nodes, 1, DROPOUT)
    nodes = sum(t[0].num_nodes for t in host_batches[0])
    edges = sum(len(src) for t in host_batches[0] for src, _ in t[0].adjacency_lists)
    config = workload_config(args, model, nodes, edges, max(1, args.gpus))
    del model, host_batches
    warmup = min(args.warmup, 2)  # whole minibatches of ~15 s each on the host cores: bounded (the line prints what was run)
    base = cpu_train_entry(args.hidden, steps=args.steps, warmup=warmup, threads=args.cpu_threads)
    print(json.dumps({
        "impl": "reference",
        "metric": "code-graphs/sec (train step, device-timed)", "value": base["value"], "unit": "graphs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config,
        "reference_arm": {"runs_on": "host CPU cores (GPUs hidden), buglab.models.train --sequential over oracle/cpu_backend.py",
                          "bounded_sample": f"each step = one minibatch of the same generator cut by the reference's own 30 000-node "
                                            f"budget (~15 graphs of ~{args.mean_nodes} nodes) instead of {args.graphs} graphs; "
                                            "graphs/s is per-graph throughput, comparable across the two step sizes",
                          "warmup_steps_run": warmup},
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }), flush=True)


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
