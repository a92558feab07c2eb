"""Data-parallel host logic with world_size 2 over gloo on CPU: rank sharding, flat-gradient all-reduce, identical
parameters on every rank after a step, and equality with single-process training on the union batch."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")


def _worker(rank: int, world: int, port: int, out_dir: str):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from buglab_b200 import distributed
    from ptgnn.baseneuralmodel.trainer import _allreduce_dense_gradients

    distributed.init_from_env("gloo")
    assert distributed.is_distributed() and distributed.world_size() == world and distributed.rank() == rank
    # independent units are sharded round-robin, every unit exactly once
    mine = list(distributed.shard_for_rank(range(10)))
    assert mine == list(range(rank, 10, world))

    torch.manual_seed(1234 + rank)  # different initial weights per rank ...
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    distributed.broadcast_module(net)  # ... until rank 0's are broadcast
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(8, 4, generator=g), torch.randn(8, 1, generator=g)
    x, y = x_all[rank::world], y_all[rank::world]  # equal-sized shards -> mean of means == global mean
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.mse_loss(net(x), y).backward()
        # flat-buffer path (what FlatAdam's bucket does on the GPU): sum then scale by 1/world
        flat = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        scale = distributed.allreduce_flat_gradient(flat)
        assert scale == 1.0 / world
        # compat path used for optimisers without a flat buffer
        _allreduce_dense_gradients(list(net.parameters()), world)
        flat2 = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
        torch.testing.assert_close(flat * scale, flat2)
        torch.nn.utils.clip_grad_norm_(net.parameters(), 0.5)  # after the all-reduce -> identical on all ranks
        opt.step()
    assert distributed.all_ranks_max(float(rank), "cpu") == world - 1
    assert distributed.all_ranks_sum(1.0, "cpu") == world
    torch.save({k: v.clone() for k, v in net.state_dict().items()}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_data_parallel_equals_single_process(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    for k in a:
        assert torch.equal(a[k], b[k]), f"ranks diverged on {k}"
    # single process on the union batch
    torch.manual_seed(1234)
    net = torch.nn.Sequential(torch.nn.Linear(4, 8), torch.nn.Tanh(), torch.nn.Linear(8, 1))
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(8, 4, generator=g), torch.randn(8, 1, generator=g)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    for _ in range(3):
        opt.zero_grad()
        torch.nn.functional.mse_loss(net(x_all), y_all).backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 0.5)
        opt.step()
    for k, v in net.state_dict().items():
        torch.testing.assert_close(a[k], v, atol=1e-6, rtol=1e-5)


def _uneven_worker(rank: int, world: int, port: int, out_dir: str):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from buglab_b200 import distributed
    from ptgnn.baseneuralmodel.trainer import _allreduce_dense_gradients, _RankSync

    distributed.init_from_env("gloo")
    # rank 0 has two more minibatches than rank 1, and the ranks' minibatches hold different numbers of graphs
    torch.manual_seed(0)
    data = torch.randn(40, 3)                      # the union data set; y = mean over graphs of (w . x)^2
    sizes = [[4, 2, 3, 4, 4], [1, 2, 3]][rank]
    starts = [[0, 5, 9, 20, 30], [4, 7, 12]][rank]
    my_batches = [(data[s: s + n], [None] * n) for s, n in zip(starts, sizes)]
    w = torch.nn.Parameter(torch.tensor([0.5, -1.0, 2.0]))
    # host_side=True: the count exchange over the separate gloo group the trainer uses next to NCCL (here next to gloo)
    sync = _RankSync("cpu", host_side=True)
    assert _RankSync._host_group is None
    seen = []
    for x, raw in sync.batches(iter(my_batches)):
        w.grad = None
        ((x @ w) ** 2).mean().backward()           # local mean over the local graphs, like the model's losses
        _allreduce_dense_gradients([w], world, sync.weight)   # the per-step collective that would dead-lock on uneven epochs
        seen.append((len(raw), sync.weight, w.grad.clone()))
    assert _RankSync._host_group is not None
    # the default (device-side exchange on the main group, what a gloo run uses) gives the same schedule
    again = [(len(raw), _s.weight) for _s in [_RankSync("cpu")] for x, raw in _s.batches(iter(my_batches))]
    assert again == [(n, wgt) for n, wgt, _ in seen]
    torch.save(seen, os.path.join(out_dir, f"uneven{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_uneven_shards_end_the_epoch_together_and_weight_by_graphs(tmp_path):
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_uneven_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    a, b = torch.load(tmp_path / "uneven0.pt"), torch.load(tmp_path / "uneven1.pt")
    assert [x[0] for x in a] == [4, 2, 3] and [x[0] for x in b] == [1, 2, 3]      # both stop after 3 steps, nobody hangs
    assert [x[1] for x in a] == [1.6, 1.0, 1.0] and [x[1] for x in b] == [0.4, 1.0, 1.0]
    torch.manual_seed(0)
    data = torch.randn(40, 3)
    w = torch.tensor([0.5, -1.0, 2.0], requires_grad=True)
    union = torch.cat([data[0:4], data[4:5]])                                       # step 0: 4 graphs on rank 0, 1 on rank 1
    expected, = torch.autograd.grad(((union @ w) ** 2).mean(), w)
    for rank_result in (a, b):
        assert torch.allclose(rank_result[0][2], expected, atol=1e-6)               # == single-device step on the union batch
    assert torch.equal(a[1][2], b[1][2]) and torch.equal(a[2][2], b[2][2])


def _reducer_worker(rank: int, world: int, port: int, out_dir: str):
    for p in (ROOT, PKG):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from buglab_b200 import distributed

    distributed.init_from_env("gloo")
    torch.manual_seed(5)
    # three "layers" + a head that only rank 1's minibatch uses: bucket completion order differs between the ranks
    layers = torch.nn.ModuleList([torch.nn.Linear(6, 6) for _ in range(3)])
    rare_head = torch.nn.Linear(6, 1)
    head = torch.nn.Linear(6, 1)
    params = [p for m in (layers, rare_head, head) for p in m.parameters()]
    offsets, total = [], 0
    for p in params:
        offsets.append(total)
        total += (p.numel() + 3) // 4 * 4          # 16-byte aligned views, like FlatAdam
    flat = torch.zeros(total)
    for p, off in zip(params, offsets):
        p.grad = flat[off: off + p.numel()].view_as(p)
    reducer = distributed.OverlappedGradientReducer(flat, params, offsets, bucket_bytes=160)
    assert reducer.num_buckets >= 3
    x = torch.randn(5, 6, generator=torch.Generator().manual_seed(100 + rank))
    weight = 0.5 if rank == 0 else 1.5            # uneven-minibatch weights (trainer._RankSync)
    results = []
    for step in range(2):
        flat.zero_()
        h = x
        for layer in layers:
            h = torch.tanh(layer(h))
        loss = head(h).sum() + (rare_head(h).sum() if rank == 1 else 0.0)
        reducer.begin(weight)
        loss.backward()
        assert all(p.grad.data_ptr() == flat.data_ptr() + 4 * off for p, off in zip(params, offsets))  # accumulated in place
        scale = reducer.finish()
        assert scale == 1.0 / world
        results.append(flat.clone())
    # the same step without the reducer: local gradient, scaled, summed in one piece
    flat.zero_()
    h = x
    for layer in layers:
        h = torch.tanh(layer(h))
    (head(h).sum() + (rare_head(h).sum() if rank == 1 else 0.0)).backward()
    expected = flat.clone() * weight
    dist.all_reduce(expected)
    torch.testing.assert_close(results[0], expected)
    torch.testing.assert_close(results[1], expected)
    reducer.close()
    torch.save(results[1], os.path.join(out_dir, f"reduced{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_overlapped_bucket_reducer_matches_one_piece_allreduce(tmp_path):
    """The bucketed reducer (hooks fire in backward order, buckets launch in a fixed order even when a rank leaves a
    module unused) gives exactly the weighted sum a single all-reduce of the flat buffer gives, on both ranks."""
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_reducer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert torch.equal(torch.load(tmp_path / "reduced0.pt"), torch.load(tmp_path / "reduced1.pt"))
