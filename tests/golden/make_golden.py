#!/usr/bin/env python
"""Generates the golden fixtures of tests/golden/ by running the REAL reference code from /root/reference.

Run in the build container only (the GPU box has no /root/reference):  python tests/golden/make_golden.py

What runs here
  * the reference's own in-repo modules, imported unmodified from /root/reference:
    buglab/models/{gnn,basemodel,modelregistry,gnnlayerdefs,utils}.py, buglab/models/layers/*.py,
    buglab/representations/data.py, buglab/utils/msgpackutils.py;
  * underneath them, because ptgnn / torch_scatter / dpu_utils are unpinned third-party packages that are not
    installable offline: this repo's host-side ptgnn / dpu_utils classes, with every COMPUTE
    class swapped for the pure-PyTorch CPU oracle (oracle/mp_ref.py, oracle/scatter_ref.py).
So the fixtures pin the in-repo reference arithmetic and integer bookkeeping (heads, losses, rewrite tables,
minibatch offsets, prediction unpacking); the ptgnn / torch_scatter layer itself stays "parity unpinned" and is
covered by the known-answer tests in tests/test_oracle_kat.py.

One patch to the reference objects: ``nn._argswap_module._input_dim = hidden`` — the attribute the reference reads
but never assigns (fixermodules.py:120); without it every ArgSwap sample raises AttributeError.
"""
import copy
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
REFERENCE = "/root/reference"
# /root/reference first: `import buglab` must resolve to the REAL reference, everything else to this repo
sys.path[:0] = [REFERENCE, PKG, ROOT]

import numpy as np  # noqa: E402
import torch  # noqa: E402

HIDDEN = 16
SEED = 20210527


def _stub_missing_third_party(name: str, **attrs):
    """Packages the reference imports at package-import time but never touches on this path."""
    try:
        __import__(name)
    except ImportError:
        mod = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(mod, k, v)
        sys.modules[name] = mod


def install_cpu_reference_backend():
    from oracle import mp_ref, scatter_ref

    _stub_missing_third_party("chardet", UniversalDetector=object)  # buglab/utils/fileopen.py:3 (file decoding; unused here)

    ts = types.ModuleType("torch_scatter")
    for name in ("scatter_max", "scatter_min", "scatter_sum", "scatter_mean"):
        fn = getattr(scatter_ref, name)
        setattr(ts, name, (lambda f: lambda src, index, dim=-1, out=None, dim_size=None: f(src, index, dim, dim_size))(fn))
    sys.modules["torch_scatter"] = ts

    import ptgnn.neuralmodels.embeddings.strelementrepresentationmodel as srm
    import ptgnn.neuralmodels.gnn.graphneuralnetwork as gnn_mod
    import ptgnn.neuralmodels.gnn.messagepassing as mp_pkg
    import ptgnn.neuralmodels.gnn.messagepassing.mlpmessagepassing as mlp_mod

    mp_pkg.MlpMessagePassingLayer = mp_ref.MlpMessagePassingLayer
    mlp_mod.MlpMessagePassingLayer = mp_ref.MlpMessagePassingLayer
    srm.SubtokenUnitEmbedder = mp_ref.SubtokenUnitEmbedder
    gnn_mod.plan_for = lambda adjacency_lists, num_nodes, block_nodes=None: None  # the plan is a GPU-side object; the oracle needs none


def main():
    install_cpu_reference_backend()
    import buglab  # the real reference

    assert buglab.__file__.startswith(REFERENCE), buglab.__file__
    from pathlib import Path

    from buglab.models.modelregistry import buggy_sample_weight_schedule, load_model
    from buglab.models.utils import LinearWarmupScheduler
    from buglab.utils.msgpackutils import load_msgpack_l_gz, save_msgpack_l_gz
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    torch.manual_seed(SEED)
    gen = SyntheticBugLabGenerator(seed=SEED, mean_nodes=90, min_nodes=40, max_nodes=200)
    samples = []
    while len(samples) < 8:
        s = gen.sample()
        samples.append(s)
    # make sure all three rewrite families and both bug / no-bug cases occur among the TARGETS
    scouts = [s["candidate_rewrite_metadata"][s["target_fix_action_idx"]][0] for s in samples
              if s["target_fix_action_idx"] is not None]
    for wanted in ("ArgSwapRewriteScout", "VariableMisuseRewriteScout", "BinaryOperatorRewriteScout"):
        if wanted not in scouts:
            for s in samples:
                idxs = [i for i, m in enumerate(s["candidate_rewrite_metadata"]) if m[0] == wanted]
                if idxs and (s["target_fix_action_idx"] is None or
                             s["candidate_rewrite_metadata"][s["target_fix_action_idx"]][0] != wanted):
                    s["target_fix_action_idx"] = idxs[0]
                    scouts.append(wanted)
                    break
    assert any(s["target_fix_action_idx"] is None for s in samples)
    shard = os.path.join(HERE, "samples.msgpack.l.gz")
    save_msgpack_l_gz(samples, shard)
    load = lambda: list(load_msgpack_l_gz(shard))  # noqa: E731  (fresh dicts: as_graph_data mutates the graph)

    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": HIDDEN, "dropout_rate": 0.0,
                              "node_representations": {"dropout_rate": 0.0, "min_freq_threshold": 1}},
                             Path("/tmp/golden.pkl.gz"))
    model.compute_metadata(iter(load()))
    nn = model.build_neural_module()
    nn._argswap_module._input_dim = HIDDEN  # F9
    nn.train()

    out = {}
    node_model = model.gnn_model.node_representation_model
    vocab = node_model.vocabulary
    out["meta/vocabulary"] = np.array(vocab.id_to_token, dtype=object)
    out["meta/edge_types"] = np.array(model.gnn_model.edge_types, dtype=object)
    out["meta/rewrite_ops"] = np.array(model._target_rewrite_ops.id_to_token, dtype=object)
    for k, v in nn.state_dict().items():
        out["state/" + k] = v.numpy()

    tensorized = [t for t, _ in model.tensorize_dataset(iter(load()), parallelize=False)]
    # per-sample rewrite tables (reference basemodel.py:80-238 via gnn.py:361-429)
    for i, t in enumerate(tensorized):
        for field in t._fields:
            if field in ("graph_data", "rewrite_logprobs"):
                continue
            v = getattr(t, field)
            out[f"tensorized/{i}/{field}"] = np.array(-1 if v is None else v, dtype=np.int64)
        for name, ids in t.graph_data.reference_nodes.items():
            out[f"tensorized/{i}/ref/{name}"] = np.asarray(ids, dtype=np.int64)
        out[f"tensorized/{i}/num_nodes"] = np.array(t.graph_data.num_nodes)

    mb, _raw = next(model.minibatch_iterator(((t, None) for t in tensorized), "cpu", max_minibatch_size=100,
                                              parallelize=False))
    for k, v in mb.items():
        if isinstance(v, torch.Tensor):
            out["mb/" + k] = v.numpy()
    g = mb["graph_data"]
    for name in g["reference_node_ids"]:
        out[f"mb/graph/ref_ids/{name}"] = g["reference_node_ids"][name].numpy()
        out[f"mb/graph/ref_graph/{name}"] = g["reference_node_graph_idx"][name].numpy()
    out["mb/graph/node_to_graph_idx"] = g["node_to_graph_idx"].numpy()
    for k, (s, t) in enumerate(g["adjacency_lists"]):
        out[f"mb/graph/adj/{k}/src"] = s.numpy()
        out[f"mb/graph/adj/{k}/tgt"] = t.numpy()

    loss = nn(**mb)
    loss.backward()
    out["out/loss"] = loss.detach().numpy()
    for k, p in nn.named_parameters():
        if p.grad is not None:
            out["grad/" + k] = p.grad.numpy()
    with torch.no_grad():
        groups, logprobs, gnn_output, _ = nn.compute_localization_logprobs(mb["graph_data"])
        out["out/node_states"] = gnn_output.output_node_representations.numpy()
        out["out/localization_groups"] = groups.numpy()
        out["out/localization_logprobs"] = logprobs.numpy()
        swap, text, misuse, sel = nn._compute_repair_logprobs(
            gnn_output, mb["target_rewrites"], mb["rewrite_to_location_group"],
            mb["candidate_symbol_to_location_group"], mb["swapped_pair_to_call_location_group"])
        out["out/argswap_logprobs"], out["out/text_logprobs"], out["out/varmisuse_logprobs"] = (
            swap.numpy(), text.numpy(), misuse.numpy())
        out["out/selected_fix_masks"] = torch.cat([sel[1], sel[2], sel[0]]).numpy()
    metrics = nn.report_metrics()
    out["out/metric_loss"] = np.array(metrics["Loss"])
    out["out/metric_localization_accuracy"] = np.array(metrics["Localization Accuracy"])

    # predict(): all-location rewrites, per-sample unpacking (gnn.py:606-645, basemodel.py:240-346)
    for i, (_point, loc, rewrites) in enumerate(model.predict(iter(load()), nn, "cpu", parallelize=False)):
        keys = sorted(loc.keys())
        out[f"predict/{i}/location_nodes"] = np.array(keys, dtype=np.int64)
        out[f"predict/{i}/location_logprobs"] = np.array([loc[k] for k in keys], dtype=np.float64)
        out[f"predict/{i}/rewrite_logprobs"] = np.array(rewrites, dtype=np.float64)

    # selector ("bug generator") mode: samples carry the detector's log-probs for every rewrite + NO_BUG, rewrites are
    # tensorised at ALL locations (gnn.py:365-366), and the loss is compute_generator_loss (utils.py:101-179)
    rng = np.random.default_rng(SEED)
    selector_samples = load()
    for s in selector_samples:
        lp = np.log(rng.uniform(0.02, 0.98, size=len(s["candidate_rewrites"]) + 1))
        lp[rng.random(lp.shape[0]) < 0.2] = -np.inf   # unobserved slots
        lp[-1] = np.log(0.5)                           # keep NO_BUG observed so no graph is empty
        s["candidate_rewrite_logprobs"] = [float(x) for x in lp]
    selector_shard = os.path.join(HERE, "selector_samples.msgpack.l.gz")
    save_msgpack_l_gz(selector_samples, selector_shard)
    with model._tensorize_all_location_rewrites():
        sel_tensorized = [t for t, _ in model.tensorize_dataset(iter(load_msgpack_l_gz(selector_shard)), parallelize=False)]
        sel_mb, _ = next(model.minibatch_iterator(((t, None) for t in sel_tensorized), "cpu", max_minibatch_size=100,
                                                  parallelize=False))
    for k, v in sel_mb.items():
        if isinstance(v, torch.Tensor):
            out["selector_mb/" + k] = v.numpy()
    for loss_type in ("classify-max-loss", "norm-kl", "norm-rmse", "expectation"):
        nn._GnnBugLabModule__generator_loss_type = loss_type
        nn.zero_grad()
        sel_loss = nn(**sel_mb)
        sel_loss.backward()
        out[f"selector/{loss_type}/loss"] = sel_loss.detach().numpy()
        out[f"selector/{loss_type}/grad_l1"] = nn._GnnBugLabModule__localization_module._l1.weight.grad.numpy().copy()

    # schedules (modelregistry.py:18-41, utils.py:55-66)
    sched_fn = buggy_sample_weight_schedule("warmdown(4, 0.25)")
    out["sched/warmdown"] = np.array([sched_fn(e) for e in range(8)])
    p = torch.nn.Parameter(torch.zeros(1))
    opt = torch.optim.Adam([p], lr=1e-4)
    warm = LinearWarmupScheduler(opt, num_warmup_steps=5)
    lrs = []
    for step in range(8):
        lrs.append(opt.param_groups[0]["lr"])
        opt.step()
        warm.step(0, step)
    out["sched/warmup_lrs"] = np.array(lrs)

    path = os.path.join(HERE, "gnn_mlp_h16.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {shard} ({os.path.getsize(shard)} B) and {path} ({os.path.getsize(path)} B); loss={float(loss):.6f}; "
          f"target scouts={sorted(set(scouts))}; samples={len(samples)}")


if __name__ == "__main__":
    main()
