#!/usr/bin/env python
"""
Usage:
    train.py [options] MODEL_NAME TRAIN_DATA_PATH VALID_DATA_PATH MODEL_FILENAME

Options:
    --aml                         Run this in Azure ML
    --amp                         Use AMP
    --azure-info=<path>           Azure authentication information file (JSON). Used to load data from Azure storage.
    --max-num-epochs=<epochs>     The maximum number of epochs to run training for. [default: 100]
    --max-files-per-fold=<n>      The maximum number of files to include in each fold.
    --minibatch-size=<size>       The minibatch size. [default: 300]
    --validate-after=<n_samples>  Run the validation after seen n_samples. [default: 1000000]
    --restore-path=<path>         The path to previous model file for starting from previous checkpoint.
    --model-spec=<json>           Extra registry keyword arguments as JSON, e.g. '{"hidden_state_size": 256}'.
    --sequential                  Do not parallelize data loading. Makes debugging easier.
    --host-loader                 Decode shards with the Python msgpack path instead of the native shard decoder.
    --quiet                       Do not show progress bar.
    -h --help                     Show this screen.
    --debug                       Enable debug routines. [default: False]
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


def run(arguments):
    from buglab_b200 import distributed

    if arguments["--aml"]:
        raise NotImplementedError("Azure ML runs are outside the scope of this build")
    distributed.init_from_env()
    rank, world = distributed.rank(), distributed.world_size()
    configure_logging(None)
    azure_info_path = arguments.get("--azure-info", None)
    max_files_per_fold = arguments["--max-files-per-fold"]
    max_files_per_fold = None if max_files_per_fold is None else int(max_files_per_fold)

    train_path = RichPath.create(arguments["TRAIN_DATA_PATH"], azure_info_path)
    valid_path = RichPath.create(arguments["VALID_DATA_PATH"], azure_info_path)
    validate_after = int(arguments["--validate-after"])
    if arguments.get("--host-loader"):
        training_data = LazyDataIterable(construct_data_loading_callable(
            train_path, shuffle=True, max_files_per_fold=max_files_per_fold, limit_num_yielded_elements=validate_after,
            rank=rank, world_size=world))
        validation_data = LazyDataIterable(construct_data_loading_callable(
            valid_path, max_files_per_fold=max_files_per_fold, rank=rank, world_size=world))
    else:
        # same file selection / sharding / limits; samples go file -> packed arrays in native code (include/buglab_shards.h)
        from buglab_b200.shards import ShardDataset

        threads = 1 if arguments["--sequential"] else None
        training_data = ShardDataset(train_path, shuffle=True, take_only_first_n_files=max_files_per_fold,
                                     limit_num_yielded_elements=validate_after, rank=rank, world_size=world,
                                     num_threads=threads)
        validation_data = ShardDataset(valid_path, take_only_first_n_files=max_files_per_fold, rank=rank,
                                       world_size=world, num_threads=threads)

    model_path = Path(arguments["MODEL_FILENAME"])
    model_spec = {"modelName": arguments["MODEL_NAME"]}
    if arguments.get("--model-spec"):
        model_spec.update(json.loads(arguments["--model-spec"]))
    model, nn, initialize_metadata = load_model(model_spec, model_path, arguments.get("--restore-path", None))

    trainer = ModelTrainer(
        model, model_path,
        max_num_epochs=int(arguments["--max-num-epochs"]),
        minibatch_size=int(arguments["--minibatch-size"]),
        optimizer_creator=optimizer,
        clip_gradient_norm=0.5,
        scheduler_creator=lambda o: LinearWarmupScheduler(o),
        enable_amp=arguments["--amp"],
    )
    if nn is not None:
        trainer.neural_module = nn
    trainer.register_train_epoch_end_hook(lambda model, nn, epoch, metrics: log_run(None, "train", model, epoch, metrics))
    trainer.register_validation_epoch_end_hook(lambda model, nn, epoch, metrics: log_run(None, "valid", model, epoch, metrics))

    if initialize_metadata:
        # every rank reads the SAME (unsharded, unshuffled-order-independent) metadata sample so that vocabularies
        # and the edge-type layout are identical everywhere
        data_for_metadata = LazyDataIterable(construct_data_loading_callable(
            RichPath.create(arguments["TRAIN_DATA_PATH"], azure_info_path), shuffle=False,
            limit_num_yielded_elements=250_000))
        trainer.load_metadata_and_create_network(data_for_metadata, not arguments["--sequential"], not arguments["--quiet"])

    trainer.train(training_data, validation_data, show_progress_bar=not arguments["--quiet"],
                  initialize_metadata=False, parallelize=not arguments["--sequential"], patience=10)


def main(argv=None):
    args = docopt(__doc__, argv)
    run_and_debug(lambda: run(args), args.get("--debug", False))


if __name__ == "__main__":
    main()
