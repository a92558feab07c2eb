"""GPU tests of the TMA-fed tcgen05 GEMMs (csrc/gemm_tma.cu) against fp64 references: gathered and contiguous
projections, segment tables with repeated types, the weight gradient, both as CTA pairs (cta_group::2, the default) and as
single CTAs (BUGLAB_B200_TMA_CG=1).  Tolerance 1e-4 abs+rel as everywhere; the observed error is asserted to be
fp32-class (< 2e-5).

The cases of one cluster size run together in one child process (the cluster size is fixed when the library is first
used; a protocol bug traps on a stalled mbarrier and kills the CUDA context, which must not take the pytest process and
the other configuration with it).  Cases after a crashed one are reported as not run."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROJECT_CASES = [
    # counts per segment, seg types (None = identity), k_in, n_out, gather, bias, prescale
    ([5], None, 64, 128, False, False, False),
    ([128, 0, 300, 1], None, 64, 128, True, True, False),
    ([700, 33, 0, 260], None, 128, 256, True, True, False),
    ([1000, 515], None, 256, 256, True, False, False),
    ([400, 77], None, 256, 512, True, True, False),
    ([5000, 3000, 2500], None, 512, 256, True, True, False),
    ([900, 0, 450], None, 256, 128, False, False, True),          # contiguous, pre-scaled tiny gradients
    ([300, 260, 10, 700, 255, 257], [2, 0, 1, 0, 2, 1], 256, 256, True, True, False),   # segments sharing weight matrices
    ([70000, 41000], None, 256, 256, True, True, False),          # several tiles per CTA pair: ring + accumulator reuse
    ([3000, 2000], None, 512, 512, False, False, True),           # the wide layers' backward shape
    ([9000, 4100, 50, 0, 300], None, 256, 256, False, True, True),  # contiguous 256 x 256 with several slabs per segment
    ([512, 512, 300], None, 64, 512, False, False, True),         # attention scores: one 64-deep chunk per tile
    ([1024, 512, 77], None, 512, 64, False, False, True),         # attention P'V / dS K: 64-wide output tile
    ([700, 33, 260], None, 256, 64, True, True, False),           # 64-wide output, gathered rows, bias
    ([2500], None, 512, 1536, False, False, True),                # sequence models: q/k/v head transforms (6 column tiles)
    ([2500], None, 2048, 512, False, False, True),                # feed-forward down projection (32 chunks deep)
]

WGRAD_CASES = [
    # counts per segment, seg types, m_out, n_in
    ([700, 0, 130], None, 256, 256),
    ([9000, 4100, 50], None, 256, 256),        # > one 4096-row slab per type
    ([5000, 300], None, 512, 512),             # wide layers: 2 x 2 output tiles per slab (pairs) / 4 x 2 (single CTAs)
    ([600, 500, 4200, 64], [1, 0, 1, 0], 256, 256),
    ([512, 512, 384], None, 512, 64),          # attention dK / dV: 64-wide products on single CTAs
    ([128, 300], None, 128, 64),
    ([5000], None, 2048, 512),                 # feed-forward weight gradients of the sequence models
    ([5000], None, 512, 2048),
]

_DRIVER = r"""
import json, sys, os
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "neurips21-self-supervised-bug-detection-and-repair_b200"))
import torch
from buglab_b200 import ops, _lib
cases = json.loads(sys.argv[1])
dev = torch.device("cuda:0")
torch.manual_seed(1)

def ptr_of(counts):
    out = [0]
    for c in counts:
        out.append(out[-1] + c)
    return out

def run_case(case):
    if case["kind"] == "project":
        counts, seg_types, k_in, n_out, gather, use_bias, prescale = case["args"]
        sp = ptr_of(counts); P = sp[-1]
        types = seg_types if seg_types is not None else list(range(len(counts)))
        K = max(types) + 1
        n_src = 3000
        g = torch.Generator().manual_seed(P + k_in + n_out)
        src = torch.randn(n_src if gather else P, k_in, generator=g) * (1e-6 if prescale else 1.0)
        idx = torch.randint(0, n_src, (P,), generator=g, dtype=torch.int32) if gather else None
        col0, ld = 8, k_in + 24
        weight = torch.randn(K, n_out, ld, generator=g) / k_in ** 0.5
        bias = torch.randn(K, n_out, generator=g) if use_bias else None
        rows = src[idx.long()] if gather else src
        ref = torch.empty(P, n_out, dtype=torch.float64)
        for s, k in enumerate(types):
            lo, hi = sp[s], sp[s + 1]
            ref[lo:hi] = rows[lo:hi].double() @ weight[k, :, col0:col0 + k_in].double().t()
            if use_bias:
                ref[lo:hi] += bias[k].double()
        src_d, weight_d = src.to(dev), weight.to(dev)
        # both operands pre-scaled by powers of two, as the model path does (mandatory for the tiny-gradient cases, optional
        # otherwise: odd case indices exercise the unscaled entry too)
        scaled = prescale or (case["index"] % 2 == 0)
        amax = ops.absmax(src_d) if scaled else None
        amax_w = ops.absmax(weight_d) if scaled else None
        a_split = ops.rows_split(src_d, None, amax)
        parts = ops.weight_parts(weight_d, n_out, k_in, col0, transposed=False, amax=amax_w)
        seg_ptr = torch.tensor(sp, dtype=torch.int32, device=dev)
        seg_type = torch.tensor(types, dtype=torch.int32, device=dev) if seg_types is not None else None
        tiles = ops.segment_units(seg_ptr, seg_type, ops.tma_tile_rows(), P)
        # slab table given: 256 x 256 products take the weight-stationary kernel (pairs only), everything else the streaming one
        slabs = ops.segment_units(seg_ptr, seg_type, ops.tma_slab_rows(), P) if os.environ.get("TEST_TMA_STATIONARY", "0") == "1" else None
        out = ops.tma_project(a_split, idx.to(dev) if gather else None, parts, bias.to(dev) if use_bias else None, amax, tiles, P,
                              slabs, amax_w)
        torch.cuda.synchronize()
        got = out.cpu().double()
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        ok = bool(torch.allclose(got, ref, atol=1e-4 * max(scale, 1e-30) if prescale else 1e-4, rtol=1e-4))
        return (dict(ok=ok, max_err=err, scale=scale, rel=err / max(scale, 1e-30)))
    else:
        counts, seg_types, m_out, n_in = case["args"]
        sp = ptr_of(counts); P = sp[-1]
        types = seg_types if seg_types is not None else list(range(len(counts)))
        K = max(types) + 1
        n_src = 2500
        g = torch.Generator().manual_seed(P + m_out)
        x = torch.randn(n_src, n_in, generator=g)
        idx = torch.randint(0, n_src, (P,), generator=g, dtype=torch.int32)
        grad = torch.randn(P, m_out, generator=g) * 1e-5
        col0, ld = n_in, 2 * n_in + 8
        ref = torch.zeros(K, m_out, n_in, dtype=torch.float64)
        for s, k in enumerate(types):
            lo, hi = sp[s], sp[s + 1]
            ref[k] += grad[lo:hi].double().t() @ x[idx[lo:hi].long()].double()
        grad_d = grad.to(dev)
        amax = torch.empty(1, device=dev)
        _lib.check(_lib.load().bl_absmax(_lib.f32(grad_d), grad_d.numel(), _lib.f32(amax), _lib.stream_ptr(dev)), "bl_absmax")
        g_split = ops.rows_split(grad_d, None, amax)
        x_d = x.to(dev)
        amax_x = ops.absmax(x_d) if case["index"] % 2 == 0 else None
        x_split = ops.rows_split(x_d, None, amax_x)
        d_weight = torch.full((K, m_out, ld), 7.0, device=dev)
        seg_ptr = torch.tensor(sp, dtype=torch.int32, device=dev)
        seg_type = torch.tensor(types, dtype=torch.int32, device=dev) if seg_types is not None else None
        ops.tma_weight_grad(g_split, x_split, idx.to(dev), amax, ops.segment_units(seg_ptr, seg_type, ops.tma_slab_rows(), P),
                            d_weight, col0, amax_x)
        torch.cuda.synchronize()
        got = d_weight[:, :, col0:col0 + n_in].cpu().double()
        untouched = bool((d_weight[:, :, :col0] == 7.0).all()) and bool((d_weight[:, :, col0 + n_in:] == 7.0).all())
        scale = float(ref.abs().max())
        err = float((got - ref).abs().max())
        return (dict(ok=bool(err <= 1e-4 * scale) and untouched, max_err=err, scale=scale, rel=err / scale, untouched=untouched))

for i, case in enumerate(cases):
    case["index"] = i
    res = run_case(case)
    res["index"] = i
    print("RESULT " + json.dumps(res), flush=True)
"""


_CACHE = {}


def _results(cg):
    """All cases at cluster size ``cg`` (``"2s"`` = pairs with the weight-stationary kernel for the 256 x 256 products), run
    once per session in a child process: {(kind, index): result dict}."""
    if cg in _CACHE:
        return _CACHE[cg]
    cases = [dict(kind="project", args=a) for a in PROJECT_CASES] + [dict(kind="wgrad", args=a) for a in WGRAD_CASES]
    stationary = "1" if str(cg).endswith("s") else "0"
    env = dict(os.environ, BUGLAB_B200_TMA_CG=str(cg)[0], TEST_TMA_STATIONARY=stationary, BUGLAB_B200_TMA_BSTAT=stationary)
    proc = subprocess.run([sys.executable, "-c", _DRIVER.format(root=ROOT), json.dumps(cases)], env=env, capture_output=True,
                          text=True, timeout=900)
    out = {}
    for line in proc.stdout.splitlines():
        if line.startswith("RESULT "):
            res = json.loads(line[len("RESULT "):])
            i = res["index"]
            out[(cases[i]["kind"], i if cases[i]["kind"] == "project" else i - len(PROJECT_CASES))] = res
    out["_tail"] = f"rc={proc.returncode}\n{proc.stdout[-1500:]}\n{proc.stderr[-3000:]}"
    _CACHE[cg] = out
    return out


def _get(cg, kind, index):
    results = _results(cg)
    assert (kind, index) in results, f"case did not run (the child process died earlier):\n{results['_tail']}"
    return results[(kind, index)]


@pytest.mark.parametrize("cg", [2, 1, "2s"])
@pytest.mark.parametrize("case_index", range(len(PROJECT_CASES)))
def test_tma_project_matches_fp64(cuda_device, case_index, cg):
    res = _get(cg, "project", case_index)
    assert res["ok"], res
    assert res["rel"] < 2e-5, res  # fp32-class accuracy, not merely inside the budget


@pytest.mark.parametrize("cg", [2, 1])
@pytest.mark.parametrize("case_index", range(len(WGRAD_CASES)))
def test_tma_weight_grad_matches_fp64(cuda_device, case_index, cg):
    res = _get(cg, "wgrad", case_index)
    assert res["ok"], res
    assert res["rel"] < 2e-5, res
