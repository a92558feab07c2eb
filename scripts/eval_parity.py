#!/usr/bin/env python
"""north_star's accuracy criterion: "evaluate.py localization/repair accuracy within +-0.1pt of the CPU reference on the
same held-out split".

Trains gnn-mlp for a few epochs on synthetic shards (GPU, through buglab.models.train), then scores ONE held-out split
twice with the SAME checkpoint: (a) buglab.models.evaluate on the B200 path, (b) the CPU oracle (oracle/model_ref.py, the
restatement of the reference's PyTorch CPU path) driven through the same ``predict`` -> ``evaluate_predictions`` code.
Prints one JSON line with both metric sets and their differences in percentage points; exit status 1 if any headline
accuracy differs by more than 0.1 pt.

    python scripts/eval_parity.py                 # GPU box: train + both evaluations
    python scripts/eval_parity.py --cpu-only      # no GPU: random-init checkpoint, oracle half only (pipeline check)

This file may import oracle/ because it is a parity check, not product code.
"""
import argparse
import json
import os
import sys
import tempfile
from pathlib import Path
from types import SimpleNamespace

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

HEADLINE = ("localization_accuracy", "repair_accuracy_given_location", "localization_and_repair_accuracy",
            "bug_detection_rate", "no_bug_recall")


class OraclePredictor:
    """Gives oracle/model_ref.GnnBugLabModule the two methods ``GnnBugLabModel.predict`` calls on a trained module."""

    def __init__(self, oracle_module):
        self._m = oracle_module

    def eval(self):
        self._m.eval()
        return self

    def compute_localization_logprobs(self, graph_data):
        groups, logprobs, states, arange = self._m.compute_localization_logprobs(graph_data)
        output = SimpleNamespace(states=states, refs=graph_data["reference_node_ids"], num_graphs=graph_data["num_graphs"])
        return groups, logprobs, output, arange

    def _compute_repair_logprobs(self, gnn_output, target_rewrites, rewrite_to_location_group,
                                 candidate_symbol_to_location_group, swapped_pair_to_call_location_group):
        return self._m._compute_repair_logprobs(gnn_output.states, gnn_output.refs, target_rewrites.long(),
                                                rewrite_to_location_group.long(), candidate_symbol_to_location_group.long(),
                                                swapped_pair_to_call_location_group.long())


def oracle_predictor(model, nn, hidden: int):
    import torch

    from oracle import model_ref

    ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types,
                                    len(model.gnn_model.node_representation_model.vocabulary),
                                    len(model._target_rewrite_ops))
    ref.load_state_dict({k: v.detach().cpu() for k, v in nn.state_dict().items()})
    ref.eval()
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    predictor = OraclePredictor(ref)
    to_oracle = predictor.compute_localization_logprobs  # the oracle indexes with int64 tensors and plain adjacency lists
    predictor.compute_localization_logprobs = lambda graph_data: to_oracle(
        model_ref.minibatch_to_cpu({"graph_data": graph_data})["graph_data"])
    return predictor


def load_heldout(data_dir: str):
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from dpu_utils.utils import RichPath

    return list(load_all_msgpack_l_gz(RichPath.create(data_dir)))


def compare_predictions(gpu_preds, cpu_preds):
    """Per-sample agreement of the two ``predict`` outputs: predicted location (arg-max over candidate nodes + NO_BUG),
    best rewrite, and the largest log-probability difference (north_star: within 1e-4)."""
    import math

    same_loc = same_rewrite = same_candidate = 0
    worst = 0.0
    for (_, loc_g, rw_g), (_, loc_c, rw_c) in zip(gpu_preds, cpu_preds):
        assert loc_g.keys() == loc_c.keys() and len(rw_g) == len(rw_c)
        same_loc += max(loc_g, key=loc_g.get) == max(loc_c, key=loc_c.get)
        # NO_BUG (key -1) starts with a constant logit of 1 and wins for weak models; the best CANDIDATE node is the
        # informative arg-max then
        cand_g = {k: v for k, v in loc_g.items() if k != -1}
        cand_c = {k: v for k, v in loc_c.items() if k != -1}
        same_candidate += (not cand_g) or max(cand_g, key=cand_g.get) == max(cand_c, key=cand_c.get)
        if rw_g:
            same_rewrite += max(range(len(rw_g)), key=rw_g.__getitem__) == max(range(len(rw_c)), key=rw_c.__getitem__)
        else:
            same_rewrite += 1
        for a, b in list(zip(loc_g.values(), loc_c.values())) + list(zip(rw_g, rw_c)):
            if math.isfinite(a) or math.isfinite(b):
                worst = max(worst, abs(a - b))
    n = max(len(gpu_preds), 1)
    return {"samples": len(gpu_preds), "same_predicted_location": same_loc / n, "same_best_candidate_node": same_candidate / n,
            "same_best_rewrite": same_rewrite / n,
            "max_abs_logprob_diff": worst}


def score(model, nn, heldout, hidden: int, device):
    """(metrics of the B200 path, metrics of the CPU oracle, per-sample comparison) for one checkpoint."""
    import copy

    import torch

    from buglab.models.evaluate import evaluate_predictions

    cpu_preds = list(model.predict(iter(copy.deepcopy(heldout)), oracle_predictor(model, nn, hidden), "cpu", parallelize=False))
    cpu = evaluate_predictions(iter(cpu_preds))
    out = {"cpu_oracle": {k: cpu[k] for k in HEADLINE}, "num_samples": cpu["num_samples"]}
    if device is not None:
        nn_dev = nn.to(device)
        gpu_preds = list(model.predict(iter(copy.deepcopy(heldout)), nn_dev, device, parallelize=False))
        gpu = evaluate_predictions(iter(gpu_preds))
        out["b200"] = {k: gpu[k] for k in HEADLINE}
        out["diff_points"] = {k: round(100.0 * (out["b200"][k] - out["cpu_oracle"][k]), 4) for k in HEADLINE}
        out["per_sample"] = compare_predictions(gpu_preds, cpu_preds)
        out["within_0.1pt"] = max(abs(v) for v in out["diff_points"].values()) <= 0.1
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cpu-only", action="store_true")
    ap.add_argument("--hidden", type=int, default=64)
    ap.add_argument("--train-graphs", type=int, default=512)
    ap.add_argument("--heldout-graphs", type=int, default=400)
    ap.add_argument("--epochs", type=int, default=4)
    ap.add_argument("--mean-nodes", type=int, default=300)
    args = ap.parse_args()

    import torch

    from buglab.models.gnn import GnnBugLabModel
    from buglab.models.modelregistry import load_model
    from buglab_b200.synthetic import SyntheticBugLabGenerator, write_shards

    work = tempfile.mkdtemp(prefix="buglab_eval_parity_")
    gen_kw = dict(mean_nodes=args.mean_nodes, min_nodes=40)
    write_shards(os.path.join(work, "heldout"), 4, args.heldout_graphs // 4, seed=4242, **gen_kw)
    heldout = load_heldout(os.path.join(work, "heldout"))
    model_path = Path(work) / "model.pkl.gz"
    out = {"hidden": args.hidden, "heldout_graphs": len(heldout)}
    device = None if args.cpu_only else torch.device("cuda:0")

    # (1) a random-initialised checkpoint: predictions are spread over all locations and rewrites, so agreement of the
    #     per-sample arg-maxes is a sharp test even though the accuracies themselves are chance level
    torch.manual_seed(0)
    model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": args.hidden}, model_path)
    model.compute_metadata(SyntheticBugLabGenerator(seed=1, **gen_kw).samples(64))
    nn = model.build_neural_module()
    out["random_init"] = score(model, nn, heldout, args.hidden, device)

    # (2) a checkpoint trained through buglab.models.train on the B200 path, scored like buglab.models.evaluate does
    if not args.cpu_only:
        from buglab.models import train

        write_shards(os.path.join(work, "train"), 8, args.train_graphs // 8, seed=1, **gen_kw)
        write_shards(os.path.join(work, "valid"), 2, 32, seed=2, **gen_kw)
        train.main(["gnn-mlp", os.path.join(work, "train"), os.path.join(work, "valid"), str(model_path),
                    f"--max-num-epochs={args.epochs}", "--minibatch-size=64", "--quiet",
                    "--model-spec", json.dumps({"hidden_state_size": args.hidden})])
        model, nn = GnnBugLabModel.restore_model(model_path, torch.device("cpu"))
        out["trained"] = score(model, nn, heldout, args.hidden, device)
        out["trained"]["checkpoint"] = f"{args.epochs} epochs on {args.train_graphs} synthetic graphs (B200 path)"

    arms = [out[k] for k in ("random_init", "trained") if k in out and "diff_points" in out[k]]
    worst = max((abs(v) for a in arms for v in a["diff_points"].values()), default=0.0)
    out["worst_diff_points"] = worst
    out["within_0.1pt"] = worst <= 0.1
    print(json.dumps(out))
    return 0 if worst <= 0.1 else 1


if __name__ == "__main__":
    sys.exit(main())
