"""bench.py's CPU-runnable parts: the reference arm (`--impl reference`: the kept train entry point over the CPU oracle on the
host cores) prints one JSON line with the contract's keys and the SAME `config` object the GPU arm prints; ranks other than 0
exit without work; helper functions that need no device."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BENCH = os.path.join(ROOT, "bench.py")


def _run(extra, env_extra=None, timeout=600):
    env = dict(os.environ)
    for key in ("RANK", "WORLD_SIZE", "LOCAL_RANK", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(key, None)
    env.update(env_extra or {})
    return subprocess.run([sys.executable, BENCH] + extra, env=env, capture_output=True, text=True, timeout=timeout)


def test_reference_arm_prints_the_contract_line_with_this_arms_config():
    args = ["--impl", "reference", "--gpus", "1", "--steps", "1", "--warmup", "0", "--hidden", "16", "--graphs", "8",
            "--mean-nodes", "300", "--cpu-threads", "2"]
    proc = _run(args)
    assert proc.returncode == 0, proc.stderr[-2000:]
    lines = [l for l in proc.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, proc.stdout[-2000:]
    line = json.loads(lines[0])
    assert line["impl"] == "reference"
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "gpu_launches"):
        assert key in line, key
    assert line["steps"] == 1 and line["warmup"] == 0 and line["n_gpus"] == 1
    assert line["value"] > 0 and line["unit"] == "graphs/s" and line["higher_is_better"] is True
    assert line["gpu_launches"] == 0
    assert line["e2e"] == {"value": line["value"], "unit": "graphs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    base = line["cpu_baseline"]
    assert base["kind"] == "port" and base["cores"] == 2 and base["value"] == line["value"] and "30 000" in base["sample"]

    # the config object is the GPU arm's: built by the same function from the same generator and seed
    sys.path.insert(0, ROOT)
    import bench

    class A:
        hidden, graphs, mean_nodes = 16, 8, 300

    model, host_batches = bench.make_workload(1000, 8, 16, 300, 1, bench.DROPOUT)
    nodes = sum(t[0].num_nodes for t in host_batches[0])
    edges = sum(len(src) for t in host_batches[0] for src, _ in t[0].adjacency_lists)
    assert line["config"] == bench.workload_config(A, model, nodes, edges, 1)
    assert line["config"]["parallelism"] == "dp1" and "8 graphs/step/GPU" in line["config"]["workload"]


def test_reference_arm_other_ranks_exit_without_work():
    proc = _run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                {"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1", "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": "29999"},
                timeout=120)
    assert proc.returncode == 0
    assert proc.stdout.strip() == ""


def test_cpu_thread_count_is_fixed_and_bounded():
    sys.path.insert(0, ROOT)
    import bench

    assert bench.cpu_thread_count(7) == 7
    assert 1 <= bench.cpu_thread_count(0) <= 32


def test_ncu_traffic_reads_only_committed_captures():
    sys.path.insert(0, ROOT)
    import bench

    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
        table = json.load(f)
    kernel = next(k for k, v in table.items() if isinstance(v, dict) and v)
    key = next(iter(table[kernel]))
    assert bench.ncu_traffic(kernel, key) == table[kernel][key]
    assert bench.ncu_traffic(kernel, "1x1x1") is None          # a shape that was never captured has no traffic figure
    assert bench.ncu_traffic("no_such_kernel", key) is None
