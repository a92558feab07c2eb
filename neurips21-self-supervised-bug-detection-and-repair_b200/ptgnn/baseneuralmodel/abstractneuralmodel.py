"""``AbstractNeuralModel``: the host half of a model — metadata, tensorisation, minibatch packing, save/restore.

Contract as used by the reference at buglab/models/gnn.py:348-604 (implemented hooks), :611-616
(``minibatch_iterator`` / ``tensorize_dataset``), modelregistry.py:154 (``restore_model``) and
tests/test_modelsync.py:21-45.
"""
import gzip
import os
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Any, Dict, Generic, Iterable, Iterator, List, Optional, Tuple, TypeVar, Union

import torch

TRawDatapoint = TypeVar("TRawDatapoint")
TTensorizedDatapoint = TypeVar("TTensorizedDatapoint")
TNeuralModule = TypeVar("TNeuralModule")


class AbstractNeuralModel(ABC, Generic[TRawDatapoint, TTensorizedDatapoint, TNeuralModule]):
    def __init__(self):
        self.__metadata_initialized = False

    # ---- metadata -------------------------------------------------------------------------------
    @abstractmethod
    def update_metadata_from(self, datapoint: TRawDatapoint) -> None:
        ...

    def finalize_metadata(self) -> None:
        pass

    def _child_models(self) -> Iterator["AbstractNeuralModel"]:
        for value in vars(self).values():
            if isinstance(value, AbstractNeuralModel):
                yield value
            elif isinstance(value, (list, tuple)):
                yield from (v for v in value if isinstance(v, AbstractNeuralModel))
            elif isinstance(value, dict):
                yield from (v for v in value.values() if isinstance(v, AbstractNeuralModel))

    def _finalize_metadata_recursive(self) -> None:
        for child in self._child_models():
            child._finalize_metadata_recursive()
        self.finalize_metadata()
        self.__metadata_initialized = True

    @property
    def metadata_initialized(self) -> bool:
        return getattr(self, "_AbstractNeuralModel__metadata_initialized", False)

    def compute_metadata(self, dataset_iterator: Iterable[TRawDatapoint], parallelize: bool = True,
                         use_multiprocessing: bool = True) -> None:
        """One pass over the data calling ``update_metadata_from``; then ``finalize_metadata`` bottom-up.
        Metadata construction is order-independent here (sorted vocabularies / edge types), so all
        data-parallel ranks that see the same data build identical metadata."""
        assert not self.metadata_initialized, "Metadata has already been initialized."
        # a data source that can run the pass itself (buglab_b200.shards.ShardDataset: native shard decoder, counts per
        # worker thread) is asked to; it declines (False) for models it does not cover
        native = getattr(dataset_iterator, "update_model_metadata", None)
        if native is None or not native(self):
            for datapoint in dataset_iterator:
                self.update_metadata_from(datapoint)
        self._finalize_metadata_recursive()

    # ---- neural module --------------------------------------------------------------------------
    @abstractmethod
    def build_neural_module(self) -> TNeuralModule:
        ...

    # ---- tensorisation --------------------------------------------------------------------------
    @abstractmethod
    def tensorize(self, datapoint: TRawDatapoint) -> Optional[TTensorizedDatapoint]:
        ...

    def tensorize_dataset(self, dataset_iterator: Iterable[TRawDatapoint], return_input_data: bool = False,
                          parallelize: bool = True, use_multiprocessing: bool = True) -> Iterator:
        """Yields ``(tensorised, raw or None)`` pairs; samples whose ``tensorize`` returns ``None`` are dropped.

        ``parallelize``: the samples are produced by ONE background thread, up to 64 ahead of the consumer, in order.
        ``tensorize`` is host-language code that holds the interpreter lock, so a pool of such threads only takes turns —
        measured 20-40 % SLOWER than a single thread on both model families (gnn-mlp host chain 78 vs 97 graphs/s, seq-great
        1 058 vs 1 726 samples/s) — while one producer still overlaps with whatever the consumer waits for outside the lock
        (device synchronisation in ``predict``, copies).  ``use_multiprocessing`` is accepted for signature compatibility."""
        def sequential() -> Iterator:
            for dp in dataset_iterator:
                tensorized = self.tensorize(dp)
                if tensorized is not None:
                    yield tensorized, (dp if return_input_data else None)

        if not parallelize:
            return sequential()
        from .trainer import _Prefetcher  # the producer-thread iterator of the training loop (no CUDA stream on "cpu")

        return iter(_Prefetcher(sequential, torch.device("cpu"), depth=64))

    # ---- minibatching ---------------------------------------------------------------------------
    @abstractmethod
    def initialize_minibatch(self) -> Dict[str, Any]:
        ...

    @abstractmethod
    def extend_minibatch_with(self, tensorized_datapoint: TTensorizedDatapoint,
                              partial_minibatch: Dict[str, Any]) -> bool:
        ...

    @abstractmethod
    def finalize_minibatch(self, accumulated_minibatch_data: Dict[str, Any],
                           device: Union[str, torch.device]) -> Dict[str, Any]:
        ...

    def minibatch_iterator(self, tensorized_data: Iterator, device: Union[str, torch.device],
                           max_minibatch_size: int, yield_partial_minibatches: bool = True,
                           show_progress_bar: bool = False, parallelize: bool = True
                           ) -> Iterator[Tuple[Dict[str, Any], List[Optional[TRawDatapoint]]]]:
        """Packs ``(tensorised, raw)`` pairs (as yielded by ``tensorize_dataset``) into minibatches
        (reference use: buglab/models/gnn.py:611-616)."""
        mb = self.initialize_minibatch()
        raw_points: List = []
        num_in_mb = 0
        for tensorized, raw in tensorized_data:
            if tensorized is None:
                continue
            keep_extending = self.extend_minibatch_with(tensorized, mb)
            raw_points.append(raw)
            num_in_mb += 1
            if not keep_extending or num_in_mb >= max_minibatch_size:
                yield self.finalize_minibatch(mb, device), raw_points
                mb, raw_points, num_in_mb = self.initialize_minibatch(), [], 0
        if yield_partial_minibatches and num_in_mb > 0:
            yield self.finalize_minibatch(mb, device), raw_points

    # ---- persistence ----------------------------------------------------------------------------
    def save(self, path: Path, model: TNeuralModule) -> None:
        """``.pkl.gz`` = gzip(torch.save((model, nn))) — the reference checkpoint format (modelregistry.py:147-156)."""
        os.makedirs(os.path.dirname(os.path.abspath(str(path))), exist_ok=True)
        with gzip.open(str(path), "wb") as f:
            torch.save((self, model), f)

    @classmethod
    def restore_model(cls, path: Path, device=None) -> Tuple["AbstractNeuralModel", TNeuralModule]:
        with gzip.open(str(path), "rb") as f:
            model, nn = torch.load(f, map_location=device, weights_only=False)
        if device is not None:
            nn.to(device)
        return model, nn
