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


def oracle_metrics(model, nn, data_dir: str, hidden: int):
    import torch

    from buglab.models.evaluate import evaluate_predictions
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from dpu_utils.utils import RichPath
    from oracle import model_ref

    ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types,
                                    len(model.gnn_model.node_representation_model.vocabulary),
                                    len(model._target_rewrite_ops))
    ref.load_state_dict({k: v.detach().cpu() for k, v in nn.state_dict().items()})
    ref.eval()
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))
    data = load_all_msgpack_l_gz(RichPath.create(data_dir))

    predictor = OraclePredictor(ref)
    to_oracle = predictor.compute_localization_logprobs  # the oracle indexes with int64 tensors and plain adjacency lists
    predictor.compute_localization_logprobs = lambda graph_data: to_oracle(
        model_ref.minibatch_to_cpu({"graph_data": graph_data})["graph_data"])
    return evaluate_predictions(model.predict(data, predictor, "cpu", parallelize=False))


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
    model_path = Path(work) / "model.pkl.gz"
    out = {"hidden": args.hidden, "heldout_graphs": args.heldout_graphs // 4 * 4}

    if args.cpu_only:
        model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": args.hidden}, model_path)
        model.compute_metadata(SyntheticBugLabGenerator(seed=1, **gen_kw).samples(64))
        nn = model.build_neural_module()
        out["checkpoint"] = "random init (pipeline check only)"
    else:
        from buglab.models import evaluate, train

        write_shards(os.path.join(work, "train"), 8, args.train_graphs // 8, seed=1, **gen_kw)
        write_shards(os.path.join(work, "valid"), 2, 32, seed=2, **gen_kw)
        train.main(["gnn-mlp", os.path.join(work, "train"), os.path.join(work, "valid"), str(model_path),
                    f"--max-num-epochs={args.epochs}", "--minibatch-size=64", "--quiet",
                    "--model-spec", json.dumps({"hidden_state_size": args.hidden})])
        out["checkpoint"] = f"trained {args.epochs} epochs on {args.train_graphs} synthetic graphs (B200 path)"
        gpu = evaluate.run({"MODEL_FILENAME": str(model_path), "TEST_DATA_PATH": os.path.join(work, "heldout"),
                            "--limit-num-elements": None, "--sequential": True, "--azure-info": None})
        out["b200"] = {k: gpu[k] for k in HEADLINE}
        model, nn = GnnBugLabModel.restore_model(model_path, torch.device("cpu"))

    cpu = oracle_metrics(model, nn, os.path.join(work, "heldout"), args.hidden)
    out["cpu_oracle"] = {k: cpu[k] for k in HEADLINE}
    out["num_samples"] = cpu["num_samples"]
    worst = 0.0
    if "b200" in out:
        out["diff_points"] = {k: round(100.0 * (out["b200"][k] - out["cpu_oracle"][k]), 4) for k in HEADLINE}
        worst = max(abs(v) for v in out["diff_points"].values())
        out["within_0.1pt"] = worst <= 0.1
    print(json.dumps(out))
    return 0 if worst <= 0.1 else 1


if __name__ == "__main__":
    sys.exit(main())
