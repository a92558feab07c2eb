#!/usr/bin/env python
"""
Usage:
    evaluate.py [options] MODEL_FILENAME TEST_DATA_PATH

Options:
    --aml                      Run this in Azure ML
    --azure-info=<path>        Azure authentication information file (JSON). Used to load data from Azure storage.
    --minibatch-size=<size>    The minibatch size. [default: 300]
    --assume-buggy             Never predict NO_BUG
    --eval-only-no-bug         Evaluate only NO_BUG samples.
    --restore-path=<path>      The path to previous model file for starting from previous checkpoint.
    --limit-num-elements=<num>  Limit the number of elements to evaluate on.
    --sequential               Do not parallelize data loading. Makes debugging easier.
    --quiet                    Do not show progress bar.
    -h --help                  Show this screen.
    --debug                    Enable debug routines. [default: False]
"""
# Entry point with the reference's command line (buglab/models/evaluate.py:2-18).  The headline numbers follow the
# reference's definitions (evaluate.py:60-173): localisation accuracy, repair accuracy given the true location,
# joint accuracy, and bug-detection / false-warning rates.  Returns the metrics dict (the reference only prints).
import math
from collections import defaultdict
from pathlib import Path
from typing import Dict

import numpy as np
import torch
from docopt import docopt
from dpu_utils.utils import RichPath, run_and_debug

from buglab.models.gnn import GnnBugLabModel


def _logsumexp(values) -> float:
    m = max(values)
    return m + math.log(sum(math.exp(v - m) for v in values))


def evaluate_predictions(predictions, assume_buggy: bool = False, eval_only_no_bug: bool = False) -> Dict[str, float]:
    n = n_loc_correct = n_buggy = n_repair_correct = n_repair_given_loc = 0
    n_buggy_warned = n_clean_silent = n_clean = 0
    per_scout_loc = defaultdict(lambda: np.zeros(2, dtype=np.int64))
    per_scout_repair = defaultdict(lambda: np.zeros(2, dtype=np.int64))
    for datapoint, location_logprobs, rewrite_probs in predictions:
        if assume_buggy:
            location_logprobs = dict(location_logprobs)
            del location_logprobs[-1]
            norm = _logsumexp(list(location_logprobs.values()))
            location_logprobs = {k: v - norm for k, v in location_logprobs.items()}
        target = datapoint["target_fix_action_idx"]
        has_bug = target is not None
        if has_bug and eval_only_no_bug:
            continue
        n += 1
        predicted_node = max(location_logprobs, key=lambda k: location_logprobs[k])
        ref_nodes = datapoint["graph"]["reference_nodes"]
        # best rewrite at the predicted location (joint prediction) and at the true location
        best_rewrite, best_lp = None, -math.inf
        for i, (node, lp) in enumerate(zip(ref_nodes, rewrite_probs)):
            if node == predicted_node and lp > best_lp:
                best_rewrite, best_lp = i, lp
        if has_bug:
            n_buggy += 1
            scout = datapoint["candidate_rewrite_metadata"][target][0]
            target_node = ref_nodes[target]
            loc_ok = predicted_node == target_node
            n_loc_correct += loc_ok
            per_scout_loc[scout] += (int(loc_ok), 1)
            n_buggy_warned += predicted_node != -1
            at_target = [(lp, i) for i, (node, lp) in enumerate(zip(ref_nodes, rewrite_probs)) if node == target_node]
            repair_ok_given_loc = max(at_target)[1] == target if at_target else False
            n_repair_given_loc += repair_ok_given_loc
            per_scout_repair[scout] += (int(repair_ok_given_loc), 1)
            n_repair_correct += loc_ok and best_rewrite == target
        else:
            n_clean += 1
            ok = predicted_node == -1
            n_loc_correct += ok
            n_clean_silent += ok
    metrics = {
        "num_samples": n,
        "localization_accuracy": n_loc_correct / n if n else float("nan"),
        "repair_accuracy_given_location": n_repair_given_loc / n_buggy if n_buggy else float("nan"),
        "localization_and_repair_accuracy": n_repair_correct / n_buggy if n_buggy else float("nan"),
        "bug_detection_rate": n_buggy_warned / n_buggy if n_buggy else float("nan"),
        "no_bug_recall": n_clean_silent / n_clean if n_clean else float("nan"),
    }
    for scout, (ok, total) in per_scout_loc.items():
        metrics[f"localization_accuracy/{scout}"] = ok / total
    for scout, (ok, total) in per_scout_repair.items():
        metrics[f"repair_accuracy_given_location/{scout}"] = ok / total
    return metrics


def evaluation_data(data_path: RichPath, limit_num_elements=None, sequential: bool = False):
    """The held-out shards as ``predict``'s data source.  ``ShardDataset`` iterates like ``load_all_msgpack_l_gz`` (sequence
    models, anything that wants raw datapoints) and additionally lets a graph model pull tensorised samples straight from
    the native shard decoder; the datapoints handed back next to the predictions are lazy views that serve the fields
    :func:`evaluate_predictions` reads without unpacking the graph (the Python chain — gzip + msgpack + tensorize — makes
    ~100 graphs/s per process and would leave the GPU idle > 95 % of an evaluation run)."""
    from buglab_b200.shards import ShardDataset

    return ShardDataset(data_path, shuffle=True, limit_num_yielded_elements=limit_num_elements,
                        num_threads=1 if sequential else None, lazy_input_data=True)


def run(arguments) -> Dict[str, float]:
    data_path = RichPath.create(arguments["TEST_DATA_PATH"], arguments.get("--azure-info", None))
    lim = None if arguments["--limit-num-elements"] is None else int(arguments["--limit-num-elements"])
    if not torch.cuda.is_available():
        raise RuntimeError("evaluate.py needs a CUDA device; the B200 build has no CPU path")
    device = torch.device("cuda")
    model, nn = GnnBugLabModel.restore_model(Path(arguments["MODEL_FILENAME"]), device)
    predictions = model.predict(evaluation_data(data_path, lim, sequential=bool(arguments["--sequential"])), nn, device,
                                parallelize=not arguments["--sequential"])
    metrics = evaluate_predictions(predictions, arguments.get("--assume-buggy", False),
                                   arguments.get("--eval-only-no-bug", False))
    for name, value in metrics.items():
        print(f"{name}: {value:.4f}" if isinstance(value, float) else f"{name}: {value}")
    return metrics


def main(argv=None):
    args = docopt(__doc__, argv)
    run_and_debug(lambda: run(args), args.get("--debug", False))


if __name__ == "__main__":
    main()
