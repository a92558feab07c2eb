#!/usr/bin/env python
"""
Usage:
    train.py [options] MODEL_NAME TRAIN_DATA_PATH VALID_DATA_PATH MODEL_FILENAME

Options:
    --aml                         Azure ML run (not available in this build).
    --amp                         Accepted for compatibility; the B200 path computes in fp32.
    --azure-info=<path>           JSON file with Azure storage credentials, for data paths that live there.
    --max-num-epochs=<epochs>     Stop after this many epochs at the latest. [default: 100]
    --max-files-per-fold=<n>      Read at most n shard files from each of the two data directories.
    --minibatch-size=<size>       Graphs per minibatch (a node budget may close a minibatch earlier). [default: 300]
    --validate-after=<n_samples>  Length of a training "epoch" in samples: validate after this many. [default: 1000000]
    --restore-path=<path>         Continue from this checkpoint instead of starting from the model spec.
    --model-spec=<json>           Extra registry keyword arguments as JSON, e.g. '{"hidden_state_size": 256}'.
    --sequential                  Load and tensorise data in the calling thread (easier to debug).
    --host-loader                 Decode shards with the Python msgpack path instead of the native shard decoder.
    --quiet                       No progress output.
    -h --help                     Print this text.
    --debug                       Drop into the debugger on an exception. [default: False]
"""
# Entry point with the reference's command line (buglab/models/train.py:2-19,54-144).  Under ``torchrun`` every
# rank runs this script: shard files are split round-robin across ranks and gradients are all-reduced over NCCL.
import json
import logging
from pathlib import Path
from typing import Callable, Iterator, Optional

from docopt import docopt
from dpu_utils.utils import RichPath, run_and_debug
from ptgnn.baseneuralmodel import ModelTrainer
from ptgnn.baseneuralmodel.utils.amlutils import configure_logging, log_run
from ptgnn.baseneuralmodel.utils.data import LazyDataIterable

from buglab.models.modelregistry import load_model
from buglab.models.utils import LinearWarmupScheduler, optimizer
from buglab.representations.data import BugLabData
from buglab.utils.msgpackutils import load_all_msgpack_l_gz

LOGGER = logging.getLogger(__name__)


def construct_data_loading_callable(data_path: RichPath, shuffle: bool = False, max_files_per_fold: Optional[int] = None,
                                    limit_num_yielded_elements: Optional[int] = None, rank: int = 0, world_size: int = 1
                                    ) -> Callable[[], Iterator[BugLabData]]:
    return lambda: load_all_msgpack_l_gz(data_path, shuffle=shuffle, take_only_first_n_files=max_files_per_fold,
                                         limit_num_yielded_elements=limit_num_yielded_elements, rank=rank,
                                         world_size=world_size)


def _optional_int(value) -> Optional[int]:
    return None if value is None else int(value)


def _data_sources(arguments, rank: int, world: int):
    """(training data, validation data, a loader for the metadata pass).  The native shard path and the host path select
    the same files, shard them over ranks the same way and stop after the same number of elements."""
    credentials = arguments.get("--azure-info", None)
    files_cap = _optional_int(arguments["--max-files-per-fold"])
    # --validate-after counts samples of the whole job (train.py:13 of the reference, single process): under data
    # parallelism every rank yields its share, so that an "epoch" covers the same number of samples at any world size
    epoch_length = max(1, -(-int(arguments["--validate-after"]) // max(world, 1)))
    folds = {name: RichPath.create(arguments[key], credentials)
             for name, key in (("train", "TRAIN_DATA_PATH"), ("valid", "VALID_DATA_PATH"))}
    if arguments.get("--host-loader"):
        def source(fold, **kw):
            return LazyDataIterable(construct_data_loading_callable(folds[fold], max_files_per_fold=files_cap, rank=rank,
                                                                    world_size=world, **kw))
        training = source("train", shuffle=True, limit_num_yielded_elements=epoch_length)
        validation = source("valid")
    else:
        from buglab_b200.shards import ShardDataset  # file -> packed arrays in native code (include/buglab_shards.h)

        threads = 1 if arguments["--sequential"] else None
        common = dict(take_only_first_n_files=files_cap, rank=rank, world_size=world, num_threads=threads)
        training = ShardDataset(folds["train"], shuffle=True, limit_num_yielded_elements=epoch_length, **common)
        validation = ShardDataset(folds["valid"], **common)
    # Every rank computes metadata from the SAME unsharded, unshuffled prefix of the training fold, so vocabularies and the
    # relation layout agree everywhere without a broadcast (the reference, single-process, samples a shuffled prefix).
    if arguments.get("--host-loader"):
        metadata = LazyDataIterable(construct_data_loading_callable(folds["train"], shuffle=False,
                                                                    limit_num_yielded_elements=250_000))
    else:
        # same prefix; a graph model's vocabulary / edge-type pass is counted by the native decoder's worker threads
        # (~6x per thread, and it scales with them), any other model iterates the raw datapoints as above
        metadata = ShardDataset(folds["train"], shuffle=False, limit_num_yielded_elements=250_000, num_threads=threads)
    return training, validation, metadata


def run(arguments):
    from buglab_b200 import distributed

    if arguments["--aml"]:
        raise NotImplementedError("Azure ML runs are outside the scope of this build")
    distributed.init_from_env()
    configure_logging(None)
    training_data, validation_data, metadata_data = _data_sources(arguments, distributed.rank(), distributed.world_size())

    spec = {"modelName": arguments["MODEL_NAME"], **json.loads(arguments.get("--model-spec") or "{}")}
    checkpoint_path = Path(arguments["MODEL_FILENAME"])
    model, restored_module, needs_metadata = load_model(spec, checkpoint_path, arguments.get("--restore-path", None))

    in_background = not arguments["--sequential"]
    verbose = not arguments["--quiet"]
    trainer = ModelTrainer(model, checkpoint_path, optimizer_creator=optimizer, scheduler_creator=LinearWarmupScheduler,
                           clip_gradient_norm=0.5, minibatch_size=int(arguments["--minibatch-size"]),
                           max_num_epochs=int(arguments["--max-num-epochs"]), enable_amp=arguments["--amp"])
    if restored_module is not None:
        trainer.neural_module = restored_module
    for phase, register in (("train", trainer.register_train_epoch_end_hook),
                            ("valid", trainer.register_validation_epoch_end_hook)):
        register(lambda model, nn, epoch, metrics, phase=phase: log_run(None, phase, model, epoch, metrics))
    if needs_metadata:
        trainer.load_metadata_and_create_network(metadata_data, in_background, verbose)
    trainer.train(training_data, validation_data, initialize_metadata=False, parallelize=in_background,
                  show_progress_bar=verbose, patience=10)
    return trainer  # callers that time the entry point (bench.py) read trainer.last_epoch_stats


def main(argv=None):
    args = docopt(__doc__, argv)
    run_and_debug(lambda: run(args), args.get("--debug", False))


if __name__ == "__main__":
    main()
