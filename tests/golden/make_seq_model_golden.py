#!/usr/bin/env python
"""Generates tests/golden/seq_model.npz + seq_samples.msgpack.l.gz by running the REAL reference sequence model
(/root/reference/buglab/models/seqmodel.py: SeqBugLabModel + SeqBugLabModule, layer types great and rat) on the CPU.

Same arrangement as make_golden.py: the reference's own modules are imported unmodified; underneath, the unpinned
third-party pieces (ptgnn's StrElementRepresentationModel / subtoken embedder, torch_scatter) are this repo's host classes
with the compute swapped for the CPU oracle.  The relational transformer layers are the reference's own (torch only).

What the fixture pins: the graph -> token-sequence projection, per-sample tensors, minibatch tensors (integer, bit-exact)
and, for stage 3, the module's loss / log-probabilities / gradients with seeded weights (dropout 0).

    python tests/golden/make_seq_model_golden.py         # build container only (needs /root/reference)
"""
import copy
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as base  # noqa: E402  (sets sys.path: reference first, then this repo)

import numpy as np  # noqa: E402
import torch  # noqa: E402

HIDDEN, HEADS, LAYERS, FF = 32, 4, 2, 64
SEED = 20210922


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.integer,)):
        return int(x)
    if isinstance(x, (np.floating,)):
        return float(x)
    return x


def main():
    base.install_cpu_reference_backend()
    import buglab

    assert buglab.__file__.startswith(base.REFERENCE), buglab.__file__
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_msgpack_l_gz, save_msgpack_l_gz
    from buglab_b200.synthetic import SyntheticProgramGenerator

    gen = SyntheticProgramGenerator(seed=SEED, statements=7)
    samples = [gen.sample() for _ in range(6)]
    samples[0]["target_fix_action_idx"] = None
    for wanted, s in zip(("ArgSwapRewriteScout", "VariableMisuseRewriteScout", "BinaryOperatorRewriteScout"), samples[1:]):
        idxs = [i for i, m in enumerate(s["candidate_rewrite_metadata"]) if m[0] == wanted]
        assert idxs, wanted
        s["target_fix_action_idx"] = idxs[len(idxs) // 2]
    # a graph the projection must reject (two token chains) and one over the length limit are part of the fixture
    broken = copy.deepcopy(samples[2])
    broken["graph"]["edges"]["NextToken"] = broken["graph"]["edges"]["NextToken"][:-3] + broken["graph"]["edges"]["NextToken"][-2:]
    too_long = SyntheticProgramGenerator(seed=SEED + 1, statements=60).sample()
    samples += [broken, too_long]
    shard = os.path.join(HERE, "seq_samples.msgpack.l.gz")
    save_msgpack_l_gz(samples, shard)
    load = lambda: list(load_msgpack_l_gz(shard))  # noqa: E731

    out = {}
    for layer_type in ("great", "rat", "transformer", "gru"):
        torch.manual_seed(SEED)
        model, _, _ = load_model({"modelName": f"seq-{layer_type}", "hidden_state_size": HIDDEN, "dropout_rate": 0.0,
                                  "num_layers": LAYERS, "num_heads": HEADS, "intermediate_dimension_size": FF,
                                  "max_seq_size": 200}, Path("/tmp/seq_golden.pkl.gz"))
        logging_off()
        model.compute_metadata(iter(load()))
        edge_type_to_idx = model._SeqBugLabModel__edge_type_to_idx
        embedder_model = model._SeqBugLabModel__token_embedder
        pre = f"{layer_type}/"
        if layer_type == "great":
            out["meta/vocabulary"] = np.array(embedder_model.vocabulary.id_to_token, dtype=object)
            out["meta/edge_types_in_reference_order"] = np.array(list(edge_type_to_idx), dtype=object)
            tensorized = [model.tensorize(dp) for dp in load()]
            out["meta/dropped"] = np.array([t is None for t in tensorized])
            for i, t in enumerate(tensorized):
                if t is None:
                    continue
                d = t._asdict()
                d["target_subtokens_ids"] = [np.asarray(x).tolist() for x in d["target_subtokens_ids"]]
                out[f"sample/{i}"] = np.array(json.dumps(jsonable(d)))
            kept = [t for t in tensorized if t is not None]
            mb = model.initialize_minibatch()
            for t in kept:
                model.extend_minibatch_with(t, mb)
            mb = model.finalize_minibatch(mb, "cpu")
            for k, v in mb.items():
                if isinstance(v, torch.Tensor):
                    out[f"mb/{k}"] = v.numpy()
            # relation ids are stored by NAME: the reference numbers them in set-iteration order (PYTHONHASHSEED)
            names = list(edge_type_to_idx)
            out["mb/edge_type_names"] = np.array([names[int(i)] for i in mb["edge_types"]], dtype=object)

        # ---- stage 3: the module (seeded weights, dropout 0, train mode so metrics run) ----
        kept = [t for t in (model.tensorize(dp) for dp in load()) if t is not None]
        mbp = model.initialize_minibatch()
        for t in kept:
            model.extend_minibatch_with(t, mbp)
        mb = model.finalize_minibatch(mbp, "cpu")
        nn = model.build_neural_module()
        nn._argswap_module._input_dim = HIDDEN  # F9 (fixermodules.py:120 reads an attribute that is never assigned)
        nn.train()
        names = list(edge_type_to_idx)
        out[pre + "edge_type_names"] = np.array([names[int(i)] for i in mb["edge_types"]], dtype=object)
        out[pre + "edge_types_in_reference_order"] = np.array(names, dtype=object)
        keep_rows = 256  # the positional table has 5000 rows; only the first max_len are ever read (or get a gradient)
        for k, v in nn.state_dict().items():
            out[pre + "param/" + k] = v.detach().numpy()[:, :keep_rows].copy() if "positional_encoding" in k else v.detach().numpy().copy()
        loss = nn(**mb)
        loss.backward()
        out[pre + "loss"] = np.array(float(loss.detach()))
        for k, p in nn.named_parameters():
            if p.grad is not None:
                out[pre + "grad/" + k] = p.grad.numpy()[:, :keep_rows].copy() if "positional_encoding" in k else p.grad.numpy().copy()
        with torch.no_grad():
            rep = nn._compute_output_representation(mb["input_sequence_ids"], mb["input_seq_num_subtokens"],
                                                    mb["token_sequence_lengths"], mb["edges"], mb["edge_types"])
            out[pre + "output_representation"] = rep.numpy().copy()
            loc = rep[mb["candidate_location_idxs"][:, 0], mb["candidate_location_idxs"][:, 1]]
            groups, logprobs = nn._compute_localization_logprobs(loc, mb["candidate_location_idxs"][:, 0], rep.shape[0])
            out[pre + "localization_groups"], out[pre + "localization_logprobs"] = groups.numpy(), logprobs.numpy()
            swap, text, misuse = nn._compute_repair_logprobs(
                rep, mb["target_rewrite_node_ids"], mb["target_rewrites"], mb["rewrite_to_location_group"],
                mb["varmisused_node_ids"], mb["candidate_symbol_node_ids"], mb["candidate_symbol_to_location_group"],
                mb["call_node_ids"], mb["candidate_swapped_node_ids"], mb["swapped_pair_to_call_location_group"])
            out[pre + "argswap_logprobs"], out[pre + "text_logprobs"] = swap.numpy(), text.numpy()
            out[pre + "varmisuse_logprobs"] = misuse.numpy()
        out[pre + "metrics"] = np.array(json.dumps(jsonable(nn.report_metrics())))
        # inference: per graph node log-probabilities and per rewrite log-probabilities (seqmodel.py:977-1031)
        predictions = []
        for dp, location_logprobs, rewrite_logprobs in model.predict(iter(load()), nn, "cpu", parallelize=False):
            predictions.append({"path": dp["graph"]["path"],
                                "locations": {str(int(k)): float(v) for k, v in location_logprobs.items()},
                                "rewrites": [float(x) for x in rewrite_logprobs]})
        out[pre + "predictions"] = np.array(json.dumps(predictions))
    path = os.path.join(HERE, "seq_model.npz")
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1e3:.0f} kB), dropped = {out['meta/dropped'].tolist()}")


def logging_off():
    import logging

    logging.getLogger("buglab.models.seqmodel").setLevel(logging.CRITICAL)


if __name__ == "__main__":
    main()
