#!/usr/bin/env python
"""BASELINE config 4: seq-great (relational transformer) hidden=512, seq_len<=512, batch=64 on one B200.

Times (CUDA events, after warm-up) the full train step of the sequence model (fwd + bwd + clip + Adam) over synthetic
programs and, separately, the attention kernels alone (bl_seq_attention_fwd / _bwd) with their algorithmic FLOP count
(4 * B * H * L^2 * d for the forward: QK^T and PV) — the first measurement of SURVEY.md §8(f) row 2.  One JSON line.

    python scripts/bench_seq.py [--steps 5] [--hidden 512] [--batch 64] [--layers 5]
    python scripts/bench_seq.py --dry-run      # no GPU: data generation, tensorisation and packing only
"""
import argparse
import json
import logging
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--hidden", type=int, default=512)
    ap.add_argument("--heads", type=int, default=8)
    ap.add_argument("--layers", type=int, default=5)
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--max-seq", type=int, default=512)
    ap.add_argument("--layer-type", default="great")
    ap.add_argument("--dry-run", action="store_true")
    args = ap.parse_args()

    from pathlib import Path

    import torch

    from buglab.models.modelregistry import load_model
    from buglab_b200.synthetic import SyntheticProgramGenerator

    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)
    model, _, _ = load_model({"modelName": f"seq-{args.layer_type}", "hidden_state_size": args.hidden, "dropout_rate": 0.0,
                              "num_layers": args.layers, "num_heads": args.heads, "max_seq_size": args.max_seq,
                              "intermediate_dimension_size": 2 * args.hidden}, Path("/tmp/_bench_seq.pkl.gz"))
    gen = SyntheticProgramGenerator(seed=0, statements=42)   # ~450 tokens per program
    t0 = time.perf_counter()
    pool = []
    while len(pool) < 2 * args.batch:
        s = gen.sample()
        pool.append(s)
    model.compute_metadata(iter(pool))
    tensorized = [t for t in (model.tensorize(dp) for dp in pool) if t is not None]
    batches = []
    for start in range(0, len(tensorized) - args.batch + 1, args.batch):
        batches.append(tensorized[start: start + args.batch])
    out = {"workload": f"seq-{args.layer_type} hidden={args.hidden} heads={args.heads} layers={args.layers} batch={args.batch} "
                       f"max_seq={args.max_seq}", "prepare_seconds": round(time.perf_counter() - t0, 1),
           "kept_samples": len(tensorized), "mean_tokens": round(sum(len(t.target_subtokens_ids) for t in tensorized) / max(1, len(tensorized)), 1)}

    def pack(chunk, device):
        mb = model.initialize_minibatch()
        for t in chunk:
            model.extend_minibatch_with(t, mb)
        return model.finalize_minibatch(mb, device)

    if args.dry_run:
        mb = pack(batches[0], "cpu")
        out["padded_length"] = int(mb["input_sequence_ids"].shape[1])
        out["edges_per_batch"] = int(mb["edges"].shape[0])
        print(json.dumps(out))
        return

    from buglab.models.utils import LinearWarmupScheduler, optimizer
    from buglab_b200 import ops

    device = torch.device("cuda", 0)
    torch.manual_seed(0)
    nn = model.build_neural_module().to(device)
    nn._argswap_module._input_dim = args.hidden
    opt = optimizer(nn.parameters())
    opt.max_grad_norm = 0.5
    sched = LinearWarmupScheduler(opt)
    nn.train()
    resident = [pack(chunk, device) for chunk in batches]

    def step(mb):
        opt.zero_grad()
        loss = nn(**mb)
        loss.backward()
        opt.step()
        sched.step(0, 0)
        return loss

    for i in range(args.warmup):
        step(resident[i % len(resident)])
    torch.cuda.synchronize(device)
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for i in range(args.steps):
        loss = step(resident[i % len(resident)])
    end.record()
    torch.cuda.synchronize(device)
    ms = start.elapsed_time(end) / args.steps
    out.update({"ms_per_step": round(ms, 2), "sequences_per_s": round(args.batch / (ms / 1e3), 1), "loss": float(loss.detach())})

    # attention kernels alone
    mb = resident[0]
    B, L = mb["input_sequence_ids"].shape[:2]
    H, d = args.heads, args.hidden // args.heads
    plan = ops.build_seq_attention_plan(mb["edges"], mb["edge_types"], mb["token_sequence_lengths"], L, len(model.edge_types))
    q, k, v = (torch.randn(B, H, L, d, device=device, requires_grad=True) for _ in range(3))
    bias = torch.randn(plan.num_tables, H, d, device=device, requires_grad=True)
    for _ in range(2):
        ops.seq_edge_attention(q, k, v, bias, None, plan).sum().backward()
    torch.cuda.synchronize(device)
    start.record()
    for _ in range(args.steps):
        o = ops.seq_edge_attention(q, k, v, bias, None, plan)
    end.record()
    torch.cuda.synchronize(device)
    fwd_ms = start.elapsed_time(end) / args.steps
    g = torch.randn_like(o)
    start.record()
    for _ in range(args.steps):
        o = ops.seq_edge_attention(q, k, v, bias, None, plan)
        o.backward(g)
    end.record()
    torch.cuda.synchronize(device)
    both_ms = start.elapsed_time(end) / args.steps
    lengths = mb["token_sequence_lengths"].double()
    flops_fwd = float((4.0 * H * d * lengths * lengths).sum())        # unmasked part of QK^T and PV
    out["attention"] = {"backend": "tensor-core (tcgen05 GEMMs + row kernels)" if ops._seq_tc_ok(q) else "cuda-core (fp32, thread per row)",
                        "padded_length": int(L), "entries": int(plan.row_key.shape[0]), "fwd_ms": round(fwd_ms, 3),
                        "fwd_bwd_ms": round(both_ms, 3), "fwd_tflops": round(flops_fwd / (fwd_ms / 1e3) / 1e12, 2)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
