#!/usr/bin/env python
"""The attention of BASELINE config 4 alone (64 sequences, 8 heads of 64, 512 tokens, ~50 k typed-edge entries): a few
forward + backward passes of ``ops.seq_edge_attention`` on the tensor-core path, for profiling under ncu:

    ncu --set full --clock-control none --import-source on -k regex:"softmax|proj_kernel|wgrad_kernel" --launch-skip 12 -c 12 \\
        -o gpurun_out/seq_attention python scripts/profile_seq_attention.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch

from buglab_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, H, L, D, T = 64, 8, 512, 64, 12
lengths = torch.randint(250, L + 1, (B,), generator=g)
lengths[0] = L
E = 25000
eb = torch.randint(0, B, (E,), generator=g)
es = (torch.rand(E, generator=g) * lengths[eb]).long()
et = (torch.rand(E, generator=g) * lengths[eb]).long()
plan = ops.build_seq_attention_plan(torch.stack((eb, es, et), dim=1).to(dev), torch.randint(0, T, (E,), generator=g).to(dev),
                                    lengths.to(dev), L, T)
q, k, v = (torch.randn(B, H, L, D, device=dev, requires_grad=True) for _ in range(3))
bias = torch.randn(2 * T, H, D, device=dev, requires_grad=True)
assert ops._seq_tc_ok(q)
for _ in range(int(os.environ.get("PASSES", "2"))):
    out = ops.seq_edge_attention(q * D ** -0.5, k, v, bias, None, plan, 0.1, True)
    out.sum().backward()
torch.cuda.synchronize()
print("done")
