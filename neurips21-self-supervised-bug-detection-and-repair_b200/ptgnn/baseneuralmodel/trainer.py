"""``ModelTrainer`` — the train/validate loop with the constructor, hooks and method names the reference uses
(buglab/models/train.py:98-134, buglab/controllers/trainbugdetector.py:73-153).

Per step (reference order): zero_grad -> loss = nn(**minibatch) -> backward ->
[all-reduce of the flat gradient bucket when torch.distributed is initialised] -> global-norm clip ->
optimizer.step -> scheduler.step(epoch_idx, epoch_step).

B200-first details: the next minibatch is packed (host, pinned) and copied H2D by a producer thread on a side
CUDA stream while the current one computes; with a ``FlatAdam`` optimiser the clip + Adam are two fused
kernels over one flat buffer, which is also the single NCCL bucket; step time is taken with CUDA events.
"""
import logging
import math
import queue
import random
import threading
import time
from abc import ABC, abstractmethod
from pathlib import Path
from typing import Any, Callable, Dict, Iterable, Iterator, List, Optional, Tuple

import torch

from .abstractneuralmodel import AbstractNeuralModel
from .modulewithmetrics import ModuleWithMetrics

LOGGER = logging.getLogger(__name__)

EndOfEpochHook = Callable[[AbstractNeuralModel, ModuleWithMetrics, int, Dict], None]


class AbstractScheduler(ABC):
    @abstractmethod
    def step(self, epoch_idx: int, epoch_step: int) -> None:
        ...


def _distributed():
    from buglab_b200 import distributed

    return distributed


def _reachable_cuda_tensors(obj, out: List[torch.Tensor], _depth: int = 0, any_device: bool = False) -> List[torch.Tensor]:
    """Every CUDA tensor reachable from a finalised minibatch: dict values, list/tuple items, ALL fields of named
    tuples (an ``EdgePlan`` starts with plain ints, so no first-element shortcut) and the ``plan`` attribute that
    ``PlannedAdjacency`` carries."""
    if isinstance(obj, torch.Tensor):
        if obj.is_cuda or any_device:  # any_device: host-side tests of the traversal itself
            out.append(obj)
        return out
    if _depth > 8:
        return out
    if isinstance(obj, dict):
        for v in obj.values():
            _reachable_cuda_tensors(v, out, _depth + 1, any_device)
    elif isinstance(obj, (list, tuple)):
        is_named = hasattr(obj, "_fields")
        # long homogeneous lists of Python scalars (e.g. per-graph counts) are skipped after a look at the first item
        if is_named or not obj or not isinstance(obj[0], (int, float, str)):
            for v in obj:
                if not isinstance(v, (int, float, str, type(None))):
                    _reachable_cuda_tensors(v, out, _depth + 1, any_device)
        plan = getattr(obj, "plan", None)
        if plan is not None:
            _reachable_cuda_tensors(plan, out, _depth + 1, any_device)
    return out


def _record_stream(obj, stream) -> int:
    """Tell the caching allocator that every tensor reachable from ``obj`` is also used on ``stream`` (they were
    allocated on the producer's side stream; without this their blocks could be recycled while still in use).
    Returns the number of tensors recorded."""
    tensors = _reachable_cuda_tensors(obj, [])
    for t in tensors:
        t.record_stream(stream)
    return len(tensors)


class _Prefetcher:
    """Runs a minibatch iterator in a producer thread, on its own CUDA stream, ``depth`` batches ahead.

    A consumer that stops early (an exception in the step, or a data-parallel epoch cut short because another rank ran
    out of data) closes the generator; that stops the producer and releases the queued minibatches instead of leaving a
    thread blocked on a full queue with device memory attached."""

    _END = object()

    def __init__(self, make_iterator: Callable[[], Iterator], device: torch.device, depth: int = 2):
        self._queue: "queue.Queue" = queue.Queue(maxsize=depth)
        self._device = device
        self._error: Optional[BaseException] = None
        self._stop = threading.Event()
        self._stream = torch.cuda.Stream(device) if device.type == "cuda" else None
        self._thread = threading.Thread(target=self._run, args=(make_iterator,), daemon=True)
        self._thread.start()

    def _put(self, entry) -> bool:
        while not self._stop.is_set():
            try:
                self._queue.put(entry, timeout=0.05)
                return True
            except queue.Full:
                continue
        return False

    def _run(self, make_iterator) -> None:
        try:
            if self._stream is not None:
                with torch.cuda.stream(self._stream):
                    for item in make_iterator():
                        event = torch.cuda.Event()
                        event.record(self._stream)
                        if not self._put((item, event)):
                            break
            else:
                for item in make_iterator():
                    if not self._put((item, None)):
                        break
        except BaseException as e:  # surfaced in the consumer
            self._error = e
        finally:
            self._put(self._END)

    def close(self) -> None:
        self._stop.set()
        while True:  # drop whatever is queued so the minibatches (and their device memory) can be freed
            try:
                self._queue.get_nowait()
            except queue.Empty:
                break
        self._thread.join(timeout=5.0)

    def __iter__(self):
        try:
            while True:
                got = self._queue.get()
                if got is self._END:
                    if self._error is not None:
                        raise self._error
                    return
                item, event = got
                if event is not None:
                    consumer = torch.cuda.current_stream(self._device)
                    consumer.wait_event(event)
                    _record_stream(item, consumer)
                yield item
        finally:
            self.close()


def _buffered_shuffle(items: Iterator, buffer_size: int, rng=random) -> Iterator:
    """Streaming shuffle of the tensorised training samples (``shuffle_training_data=True``, ptgnn's default): fill a
    buffer, shuffle it, emit half, refill.  Shard files are already visited in random order (msgpackutils); this mixes
    samples across neighbouring files without holding the data set in memory."""
    buffer: List = []
    for item in items:
        buffer.append(item)
        if len(buffer) >= buffer_size:
            rng.shuffle(buffer)
            half = len(buffer) // 2
            yield from buffer[half:]
            del buffer[half:]
    rng.shuffle(buffer)
    yield from buffer


class _RankSync:
    """Per-step agreement between data-parallel ranks (one tiny all-gather; a no-op without torch.distributed).

    * Ranks read different shards and may run out of minibatches at different times; every step contains a collective,
      so the epoch ends for all ranks as soon as ANY rank is exhausted (the others would otherwise wait in the gradient
      all-reduce forever).
    * The losses are means over the graphs of the LOCAL minibatch (gnn.py:251, localizationmodule.py:117).  When ranks
      hold different numbers of graphs, averaging the ranks' gradients is not the gradient of the mean over the union;
      ``weight`` = B_rank * world / sum(B) is what the local gradient must be scaled by BEFORE the all-reduce so that the
      data-parallel step equals the single-device step on the union minibatch (1.0 when every rank has the same B)."""

    _host_group = None  # process-wide: a gloo group over all ranks for host-side agreement (created once, by every rank)

    def __init__(self, device, host_side: Optional[bool] = None):
        self.weight = 1.0
        self._device = device if torch.device(device).type == "cuda" else "cpu"
        self._host_side = host_side

    @classmethod
    def _host_side_group(cls):
        """With NCCL the count exchange would be a device collective ordered BEHIND the kernels of the step before it, and
        reading its result would stall the host once per step — the launch work of step i+1 could never overlap the device
        work of step i.  The counts are host integers, so they are exchanged over a gloo group (host memory, no stream)."""
        import torch.distributed as tdist

        if cls._host_group is None:
            cls._host_group = tdist.new_group(backend="gloo")
        return cls._host_group

    def batches(self, batches: Iterator) -> Iterator:
        dist = _distributed()
        if not dist.is_distributed():
            yield from batches
            return
        import torch.distributed as tdist

        world = dist.world_size()
        host_side = self._host_side if self._host_side is not None else (tdist.get_backend() == "nccl")
        group = self._host_side_group() if host_side else None
        device = "cpu" if host_side else self._device
        it = iter(batches)
        while True:
            try:
                item = next(it)
                local = len(item[1])  # (minibatch, raw datapoints)
            except StopIteration:
                item, local = None, -1
            mine = torch.tensor([local], dtype=torch.int64, device=device)
            everyone = [torch.empty_like(mine) for _ in range(world)]
            tdist.all_gather(everyone, mine, group=group)
            sizes = [int(t.item()) for t in everyone]
            if min(sizes) < 0:
                return
            total = sum(sizes)
            self.weight = float(local) * world / total if total > 0 else 1.0
            yield item


class ModelTrainer:
    def __init__(
        self,
        model: AbstractNeuralModel,
        save_location: Path,
        *,
        max_num_epochs: int = 100,
        minibatch_size: int = 200,
        optimizer_creator: Optional[Callable[[Iterable[torch.Tensor]], torch.optim.Optimizer]] = None,
        scheduler_creator: Optional[Callable[[torch.optim.Optimizer], AbstractScheduler]] = None,
        clip_gradient_norm: Optional[float] = None,
        target_validation_metric: Optional[str] = None,
        target_validation_metric_higher_is_better: bool = False,
        enable_amp: bool = False,
        catch_cuda_ooms: bool = False,
    ):
        self.__model = model
        self.__neural_network: Optional[ModuleWithMetrics] = None
        assert str(save_location).endswith(".pkl.gz"), "All models are stored as .pkl.gz."
        self.__save_location = Path(save_location)
        self._max_num_epochs = max_num_epochs
        self._minibatch_size = minibatch_size
        self._optimizer_creator = optimizer_creator or (lambda p: torch.optim.Adam(p, lr=1e-4))
        self._scheduler_creator = scheduler_creator
        self._clip_gradient_norm = clip_gradient_norm
        self._target_metric = target_validation_metric
        self._target_metric_higher_is_better = target_validation_metric_higher_is_better
        if enable_amp:
            LOGGER.warning("--amp accepted but ignored: the B200 path computes in fp32 (1e-4 parity target).")
        self._train_epoch_end_hooks: List[EndOfEpochHook] = []
        self._validation_epoch_end_hooks: List[EndOfEpochHook] = []
        self._training_start_hooks: List[Callable] = []
        self._train_metrics_reporters: List[Callable] = []
        self.last_epoch_stats: Dict[str, float] = {}
        self.shuffle_buffer_size = 2048  # tensorised samples held by the streaming shuffle of the training data

    # ---- accessors ------------------------------------------------------------------------------
    @property
    def model(self) -> AbstractNeuralModel:
        return self.__model

    @property
    def neural_module(self) -> ModuleWithMetrics:
        if self.__neural_network is None:
            raise Exception("Neural network has not been built yet (call load_metadata_and_create_network).")
        return self.__neural_network

    @neural_module.setter
    def neural_module(self, nn: ModuleWithMetrics) -> None:
        self.__neural_network = nn

    # ---- hooks ----------------------------------------------------------------------------------
    def register_train_epoch_end_hook(self, hook: EndOfEpochHook) -> None:
        self._train_epoch_end_hooks.append(hook)

    def register_validation_epoch_end_hook(self, hook: EndOfEpochHook) -> None:
        self._validation_epoch_end_hooks.append(hook)

    def register_training_start_hook(self, hook: Callable[[AbstractNeuralModel, ModuleWithMetrics, torch.optim.Optimizer], None]) -> None:
        self._training_start_hooks.append(hook)

    # ---- set-up ---------------------------------------------------------------------------------
    def load_metadata_and_create_network(self, training_data: Iterable, parallelize: bool = True,
                                         show_progress_bar: bool = True) -> None:
        LOGGER.info("Computing model metadata...")
        self.__model.compute_metadata(training_data if hasattr(training_data, "update_model_metadata") else iter(training_data),
                                      parallelize)
        self.__neural_network = self.__model.build_neural_module()
        LOGGER.info("Model has %s trainable parameters.",
                    sum(p.numel() for p in self.__neural_network.parameters() if p.requires_grad))

    def _create_optimizer(self, parameters) -> torch.optim.Optimizer:
        optimizer = self._optimizer_creator(parameters)
        if getattr(optimizer, "fused_clip", False) and getattr(optimizer, "max_grad_norm", None) is None:
            optimizer.max_grad_norm = self._clip_gradient_norm
        return optimizer

    def _minibatches(self, tensors: Iterable, device, parallelize: bool) -> Iterator:
        def make():
            return self.__model.minibatch_iterator(iter(tensors), device=device,
                                                   max_minibatch_size=self._minibatch_size, parallelize=parallelize)

        if parallelize and torch.device(device).type == "cuda":
            return iter(_Prefetcher(make, torch.device(device)))
        return make()

    # ---- one training epoch ---------------------------------------------------------------------
    def _run_training(self, training_tensors: Iterable, epoch: int, device, optimizer: torch.optim.Optimizer,
                      scheduler: Optional[AbstractScheduler], parallelize: bool, show_progress_bar: bool) -> float:
        dist = _distributed()
        nn = self.neural_module
        nn.train()
        nn.reset_metrics()
        use_cuda = torch.device(device).type == "cuda"
        fused = getattr(optimizer, "fused_clip", False)
        params = [p for p in nn.parameters() if p.requires_grad]
        loss_sum = torch.zeros((), device=device, dtype=torch.float64)
        num_steps, num_samples = 0, 0
        start_evt = end_evt = None
        if use_cuda:
            start_evt, end_evt = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            start_evt.record()
        t0 = time.perf_counter()
        sync = _RankSync(device)
        reducer = optimizer.gradient_reducer() if (fused and dist.is_distributed()) else None
        for step_idx, (mb_data, raw_points) in enumerate(sync.batches(self._minibatches(training_tensors, device, parallelize))):
            optimizer.zero_grad()
            loss = nn(**mb_data)
            if reducer is not None:
                # bucketed all-reduce on a side stream while backward runs; sync.weight != 1 only for uneven minibatches
                reducer.begin(sync.weight)
                loss.backward()
                optimizer.step(grad_scale=reducer.finish())
            elif fused:
                loss.backward()
                if sync.weight != 1.0:
                    optimizer.flat_grad.mul_(sync.weight)
                optimizer.step(grad_scale=1.0)
            else:
                loss.backward()
                if dist.is_distributed():
                    _allreduce_dense_gradients(params, dist.world_size(), sync.weight)
                if self._clip_gradient_norm is not None:
                    torch.nn.utils.clip_grad_norm_(params, self._clip_gradient_norm)
                optimizer.step()
            if scheduler is not None:
                scheduler.step(epoch_idx=epoch, epoch_step=step_idx)
            loss_sum += loss.detach().double()
            num_steps += 1
            num_samples += len(raw_points)
        if num_steps == 0:
            raise RuntimeError("No training minibatches were produced.")
        if use_cuda:
            end_evt.record()
            end_evt.synchronize()
            elapsed = start_evt.elapsed_time(end_evt) / 1e3
        else:
            elapsed = time.perf_counter() - t0
        self.last_epoch_stats = dict(train_seconds=elapsed, train_steps=num_steps, train_samples=num_samples,
                                     samples_per_second=num_samples / max(elapsed, 1e-9))
        mean_loss = float(loss_sum) / num_steps
        metrics = dict(nn.report_metrics())
        LOGGER.info("Epoch %i: Train Loss %.4f | %.1f samples/s | %s", epoch + 1, mean_loss,
                    self.last_epoch_stats["samples_per_second"], metrics)
        for hook in self._train_epoch_end_hooks:
            hook(self.__model, nn, epoch, metrics)
        return mean_loss

    # ---- validation (private name kept: controllers monkey-patch it, trainbugdetector.py:128-143) ---
    def _run_validation(self, validation_tensors: Iterable, epoch: int, best_target_metric: float, device,
                        parallelize: bool, show_progress_bar: bool) -> Tuple[float, bool]:
        dist = _distributed()
        nn = self.neural_module
        nn.eval()
        nn.reset_metrics()
        # Each minibatch loss is a mean over its graphs (gnn.py:251); weighting by the graph count makes the reported
        # validation loss — and, across data-parallel ranks holding different numbers of samples, the target metric —
        # the value over the union of the validation set, so the improved / patience decision matches a one-device run.
        loss_sum = torch.zeros((), device=device, dtype=torch.float64)
        num_samples = 0
        with torch.no_grad():
            for mb_data, raw_points in self._minibatches(validation_tensors, device, parallelize):
                n = len(raw_points)
                loss_sum += nn(**mb_data).double() * n
                num_samples += n
        total, samples = float(loss_sum), float(num_samples)
        if dist.is_distributed():
            total, samples = dist.all_ranks_sum(total, device), dist.all_ranks_sum(samples, device)
        if samples == 0:
            raise RuntimeError("No validation minibatches were produced.")
        validation_loss = total / samples
        metrics = dict(nn.report_metrics())
        if self._target_metric is not None:
            if dist.is_distributed():
                # per-sample ratios: sum(metric_r * n_r) / sum(n_r) is the metric over the union of the ranks' shards; a rank
                # whose share of a tiny validation set is empty reports no metrics and contributes 0 * 0
                local = float(metrics[self._target_metric]) if num_samples > 0 else 0.0
                target_metric = dist.all_ranks_sum(local * num_samples, device) / samples
            else:
                target_metric = metrics[self._target_metric]
            improved = target_metric > best_target_metric if self._target_metric_higher_is_better \
                else target_metric < best_target_metric
        else:
            target_metric = validation_loss
            improved = target_metric < best_target_metric
        LOGGER.info("Epoch %i: Valid Loss %.4f %s", epoch + 1, validation_loss, metrics)
        for hook in self._validation_epoch_end_hooks:
            hook(self.__model, nn, epoch, metrics)
        if improved:
            LOGGER.info("Best validation metric so far: %.4f (was %.4f).", target_metric, best_target_metric)
        return target_metric, improved

    def _save_checkpoint(self) -> None:
        if _distributed().rank() == 0:
            self.__model.save(self.__save_location, self.neural_module)

    # ---- the loop -------------------------------------------------------------------------------
    def train(self, training_data: Iterable, validation_data: Iterable, show_progress_bar: bool = True,
              validate_on_start: bool = True, initialize_metadata: bool = True, parallelize: bool = True,
              use_multiprocessing: bool = True, patience: int = 5, store_tensorized_data_in_memory: bool = False,
              shuffle_training_data: bool = True, device: Optional[str] = None) -> None:
        dist = _distributed()
        if initialize_metadata:
            self.load_metadata_and_create_network(training_data, parallelize, show_progress_bar)
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        device = torch.device(device)
        nn = self.neural_module
        nn.to(device)
        dist.broadcast_module(nn)
        model = self.__model

        def tensors_of(data):
            # a data source that can tensorise itself (buglab_b200.shards.ShardDataset: native shard decoder) is asked to;
            # anything else is a plain iterable of raw datapoints, as in the reference
            if hasattr(data, "tensorized"):
                return data.tensorized(model)
            return model.tensorize_dataset(iter(data), return_input_data=False, parallelize=parallelize)

        def training_tensors():
            if shuffle_training_data and not store_tensorized_data_in_memory:
                return _buffered_shuffle(tensors_of(training_data), self.shuffle_buffer_size)
            return tensors_of(training_data)

        def validation_tensors():
            return tensors_of(validation_data)

        class _Re:
            def __init__(self, f):
                self.f = f

            def __iter__(self):
                return self.f()

        train_it, valid_it = _Re(training_tensors), _Re(validation_tensors)
        if store_tensorized_data_in_memory:
            train_list, valid_list = list(train_it), list(valid_it)

            def reshuffled():
                if shuffle_training_data:
                    random.shuffle(train_list)
                return iter(train_list)

            train_it, valid_it = _Re(reshuffled), valid_list

        optimizer = self._create_optimizer([p for p in nn.parameters() if p.requires_grad])
        scheduler = None if self._scheduler_creator is None else self._scheduler_creator(optimizer)
        for hook in self._training_start_hooks:
            hook(model, nn, optimizer)

        best_metric = -math.inf if (self._target_metric is not None and self._target_metric_higher_is_better) else math.inf
        if validate_on_start:
            best_metric, _ = self._run_validation(valid_it, 0, best_metric, device, parallelize, show_progress_bar)
            self._save_checkpoint()
        num_epochs_not_improved = 0
        for epoch in range(self._max_num_epochs):
            self._run_training(train_it, epoch, device, optimizer, scheduler, parallelize, show_progress_bar)
            target_metric, improved = self._run_validation(valid_it, epoch, best_metric, device, parallelize,
                                                           show_progress_bar)
            if improved:
                best_metric = target_metric
                num_epochs_not_improved = 0
                self._save_checkpoint()
            else:
                num_epochs_not_improved += 1
                if num_epochs_not_improved > patience:
                    LOGGER.warning("After %s epochs loss has not improved. Stopping.", num_epochs_not_improved)
                    break


def _allreduce_dense_gradients(params: List[torch.nn.Parameter], world: int, weight: float = 1.0) -> None:
    """Bucketed all-reduce for optimisers that do not own a flat buffer (compat path, CPU/gloo tests): every rank ends up
    with ``sum_r(weight_r * grad_r) / world``."""
    import torch.distributed as tdist

    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    if weight != 1.0:
        flat.mul_(weight)
    tdist.all_reduce(flat, op=tdist.ReduceOp.SUM)
    flat.div_(world)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off: off + n].view_as(g))
        off += n
