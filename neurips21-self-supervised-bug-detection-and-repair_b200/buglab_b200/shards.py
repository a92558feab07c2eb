"""Native ``.msgpack.l.gz`` shard path: file -> packed per-graph tensors without building Python objects.

ctypes binding of ``include/buglab_shards.h`` (host-only ``libbuglab_shards.so``, built by ``csrc/build.sh``) plus the
glue that turns its arrays into the same ``BaseTensorizedBugLabGnn`` tuples ``GnnBugLabModel.tensorize`` produces
(reference call chain: buglab/utils/msgpackutils.py:11-45 -> buglab/representations/data.py:139-167 ->
buglab/models/gnn.py:463-520).  The decode runs outside the GIL, so the loader's worker threads scale with cores.

Samples the native code declines (``BL_SAMPLE_NEEDS_HOST``: exotic Unicode case mapping, malformed fields, ...) are
unpacked with ``msgpack`` and go through ``model.tensorize`` — the reference-shaped path — so results and exceptions are
the same either way.  There is no silent degradation: a missing library raises at construction.
"""
import ctypes
import logging
import os
import threading
from collections import defaultdict
from concurrent.futures import ThreadPoolExecutor
from collections.abc import Mapping
from typing import Any, Dict, Iterable, Iterator, List, Optional, Sequence, Tuple

import msgpack
import numpy as np

LOGGER = logging.getLogger(__name__)

_LIB_NAME = "libbuglab_shards.so"
_lib: Optional[ctypes.CDLL] = None
_lock = threading.Lock()

c_i32, c_i64, c_ptr = ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p

SAMPLE_OK, SAMPLE_NIL, SAMPLE_NEEDS_HOST = 0, 1, 2
SPLIT_TOKEN, SPLIT_SUBTOKEN = 0, 1


class SampleView(ctypes.Structure):
    """``bl_sample_view`` (include/buglab_shards.h), field for field."""
    _fields_ = [
        ("status", c_i32), ("num_file_nodes", c_i32), ("num_nodes", c_i32), ("max_subtokens", c_i32),
        ("node_ids", c_ptr), ("node_lens", c_ptr),
        ("num_edge_types", c_i32), ("num_reference_nodes", c_i32),
        ("edge_offsets", c_ptr), ("edge_src", c_ptr), ("edge_tgt", c_ptr), ("reference_nodes", c_ptr),
        ("num_call_args", c_i32), ("has_target", c_i32),
        ("call_args", c_ptr),
        ("target_fix_action_idx", c_i64),
        ("raw", c_ptr), ("raw_len", c_i64),
        ("rewrites_off", c_i64), ("rewrites_len", c_i64),
        ("metadata_off", c_i64), ("metadata_len", c_i64),
        ("logprobs_off", c_i64), ("logprobs_len", c_i64),
    ]


_SIGNATURES = {
    "bl_shards_version": (c_i32, []),
    "bl_shards_error_string": (ctypes.c_char_p, [c_i32]),
    "bl_shard_open": (c_i32, [ctypes.c_char_p, ctypes.POINTER(c_ptr)]),
    "bl_shard_open_buffer": (c_i32, [c_ptr, c_i64, ctypes.POINTER(c_ptr)]),
    "bl_shard_close": (None, [c_ptr]),
    "bl_shard_num_objects": (c_i64, [c_ptr]),
    "bl_shard_raw_bytes": (c_i64, [c_ptr]),
    "bl_shard_status": (c_i32, [c_ptr]),
    "bl_shard_object": (c_i32, [c_ptr, c_i64, ctypes.POINTER(c_ptr), ctypes.POINTER(c_i64)]),
    "bl_tokenizer_create": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_i32, c_i32, c_i32, c_ptr, c_i32, ctypes.POINTER(c_ptr)]),
    "bl_tokenizer_destroy": (None, [c_ptr]),
    "bl_tokenizer_ids": (c_i32, [c_ptr, c_ptr, c_i64, c_ptr]),
    "bl_sample_create": (c_i32, [ctypes.POINTER(c_ptr)]),
    "bl_sample_destroy": (None, [c_ptr]),
    "bl_sample_decode": (c_i32, [c_ptr, c_i64, c_ptr, c_ptr, c_i32, c_ptr, ctypes.POINTER(SampleView)]),
    "bl_sample_decode_many": (c_i32, [c_ptr, c_ptr, c_i32, c_ptr, c_ptr, c_i32, c_ptr, c_ptr]),
    "bl_pyset_iteration_order": (c_i64, [c_ptr, c_i64, c_ptr]),
    "bl_shard_non_nil": (c_i64, [c_ptr, c_ptr]),
    "bl_metadata_create": (c_i32, [ctypes.POINTER(c_ptr)]),
    "bl_metadata_destroy": (None, [c_ptr]),
    "bl_metadata_add": (c_i32, [c_ptr, c_ptr, c_ptr, c_i32, c_ptr, c_ptr, c_ptr, ctypes.POINTER(c_i32)]),
    "bl_metadata_num_samples": (c_i64, [c_ptr]),
    "bl_metadata_size": (c_i64, [c_ptr, c_i32, ctypes.POINTER(c_i64)]),
    "bl_metadata_export": (c_i32, [c_ptr, c_i32, c_ptr, c_ptr, c_ptr]),
}


def library_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), _LIB_NAME)


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                path = library_path()
                if not os.path.exists(path):
                    raise RuntimeError(f"{_LIB_NAME} is not built ({path}); run csrc/build.sh or __graft_entry__.build()")
                handle = ctypes.CDLL(path)
                for name, (restype, argtypes) in _SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = restype, argtypes
                _check_set_order(handle)
                _lib = handle
    return _lib


def check(code: int, what: str) -> None:
    if code != 0:
        raise RuntimeError(f"{what}: {lib().bl_shards_error_string(code).decode()}")


def pyset_iteration_order(values: Sequence[int], handle: Optional[ctypes.CDLL] = None) -> List[int]:
    handle = handle or lib()
    a = np.ascontiguousarray(values, dtype=np.int64)
    out = np.empty(max(1, a.size), dtype=np.int64)
    k = handle.bl_pyset_iteration_order(a.ctypes.data, a.size, out.ctypes.data)
    if k < 0:
        raise ValueError("unsupported set element")
    return out[:k].tolist()


def _check_set_order(handle: ctypes.CDLL) -> None:
    """data.py:103-108 makes the iteration order of a CPython ``set`` of token ids load-bearing (it numbers the subtoken
    nodes).  The native model of that order is verified against the running interpreter before first use; an interpreter
    with a different set implementation must not silently produce differently numbered graphs."""
    rng = np.random.default_rng(7)
    probes = [list(range(0, 400, 3)) + list(range(1, 400, 3)),
              rng.integers(0, 5000, size=3000).tolist(),
              rng.integers(0, 1 << 40, size=500).tolist(),
              [i // 2 + (i % 2) for i in range(70000)]]
    for values in probes:
        expected = set()
        for v in values:
            expected.add(v)
        if pyset_iteration_order(values, handle) != list(expected):
            raise RuntimeError("libbuglab_shards: CPython set-order model does not match this interpreter; "
                               "use the host-language loader (buglab.utils.msgpackutils.load_msgpack_l_gz)")


def lower_variant_codepoints() -> np.ndarray:
    """Non-ASCII code points that the native splitter (ASCII character classes, ASCII lower-casing) would treat
    differently from the host: those ``str.lower()`` changes and those that are letters or digits for
    ``str.isalnum()`` (the identifier splitter's classes), from THIS interpreter's Unicode tables.  Labels containing
    one are left to the host path, so native and host tokenisation cannot disagree."""
    global _LOWER_VARIANT
    if _LOWER_VARIANT is None:
        cps = [c for c in range(0x80, 0x110000) if not (0xD800 <= c <= 0xDFFF)
               and (chr(c).lower() != chr(c) or chr(c).isalnum())]
        _LOWER_VARIANT = np.array(cps, dtype=np.int32)
    return _LOWER_VARIANT


_LOWER_VARIANT: Optional[np.ndarray] = None


class Shard:
    """One inflated + indexed ``.msgpack.l.gz`` file."""

    def __init__(self, path: Optional[str] = None, gz_bytes: Optional[bytes] = None):
        handle = c_ptr()
        if path is not None:
            check(lib().bl_shard_open(os.fsencode(path), ctypes.byref(handle)), f"open {path}")
        else:
            buf = (ctypes.c_char * len(gz_bytes)).from_buffer_copy(gz_bytes) if gz_bytes else None
            check(lib().bl_shard_open_buffer(ctypes.cast(buf, c_ptr) if buf is not None else None, len(gz_bytes or b""),
                                             ctypes.byref(handle)), "open buffer")
        self._h = handle
        self.path = path

    def __len__(self) -> int:
        return int(lib().bl_shard_num_objects(self._h))

    @property
    def raw_bytes(self) -> int:
        return int(lib().bl_shard_raw_bytes(self._h))

    @property
    def status(self) -> int:
        """0, or the error code that cut the stream short (objects before the break are still served)."""
        return int(lib().bl_shard_status(self._h))

    def non_nil_indices(self) -> List[int]:
        """Indices of the objects that are not msgpack nil (the loader's element limit counts those, msgpackutils.py:38-43)."""
        out = (c_i64 * max(1, len(self)))()
        n = lib().bl_shard_non_nil(self._h, out)
        return list(out[:n])

    def object_bytes(self, index: int) -> bytes:
        data, n = c_ptr(), c_i64()
        check(lib().bl_shard_object(self._h, index, ctypes.byref(data), ctypes.byref(n)), "object")
        return ctypes.string_at(data.value, n.value)

    def close(self) -> None:
        if self._h is not None and self._h.value:
            lib().bl_shard_close(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Tokenizer:
    """The node-label vocabulary handed to the native side (ptgnn StrElementRepresentationModel semantics)."""

    def __init__(self, vocabulary, splitting_kind: str, max_num_subtokens: int):
        # vocabulary None: a splitter only (metadata pass: the vocabulary is what is being computed)
        tokens = list(vocabulary.token_to_id.items()) if vocabulary is not None else []
        encoded = [t.encode("utf-8", "surrogatepass") for t, _ in tokens]
        offsets = np.zeros(len(encoded) + 1, dtype=np.int64)
        if encoded:
            np.cumsum([len(e) for e in encoded], out=offsets[1:])
        blob = np.frombuffer(b"".join(encoded) or b"\0", dtype=np.uint8)
        ids = np.array([i for _, i in tokens], dtype=np.int32)
        unk = vocabulary.token_to_id.get(vocabulary.get_unk(), -1) if vocabulary is not None else -1
        variants = lower_variant_codepoints()
        kind = {"token": SPLIT_TOKEN, "subtoken": SPLIT_SUBTOKEN}[splitting_kind]
        handle = c_ptr()
        check(lib().bl_tokenizer_create(blob.ctypes.data, offsets.ctypes.data, ids.ctypes.data, len(tokens), int(unk), kind,
                                        int(max_num_subtokens), variants.ctypes.data, variants.size, ctypes.byref(handle)),
              "tokenizer")
        self._h = handle
        self.max_subtokens = 1 if kind == SPLIT_TOKEN else int(max_num_subtokens)

    def ids(self, label: str) -> Optional[Tuple[int, ...]]:
        raw = label.encode("utf-8", "surrogatepass")
        out = np.zeros(self.max_subtokens, dtype=np.int32)
        n = lib().bl_tokenizer_ids(self._h, raw, len(raw), out.ctypes.data)
        return None if n < 0 else tuple(out[:n].tolist())

    def __del__(self):
        try:
            if self._h is not None and self._h.value:
                lib().bl_tokenizer_destroy(self._h)
                self._h = None
        except Exception:
            pass


_ITEMSIZE = {np.int32: 4, np.int64: 8}


def _as_array(ptr: Optional[int], count: int, dtype) -> np.ndarray:
    """A writable numpy copy of ``count`` elements at ``ptr`` (memory owned by a ``bl_sample``: valid only until its next
    decode)."""
    if count == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.frombuffer((ctypes.c_char * (count * _ITEMSIZE[dtype])).from_address(ptr), dtype=dtype).copy()


class _SampleBuffer:
    """Per-thread ``bl_sample`` handle (its vectors are reused from sample to sample)."""

    def __init__(self):
        self.handle = c_ptr()
        check(lib().bl_sample_create(ctypes.byref(self.handle)), "sample")
        self.view = SampleView()

    def __del__(self):
        try:
            if self.handle is not None and self.handle.value:
                lib().bl_sample_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def _unpack_like_the_host_chain(sample_bytes: bytes) -> Dict[str, Any]:
    """``msgpack.unpackb`` + the in-place graph extension ``BugLabData.as_graph_data`` performs while tensorising: the
    object a consumer of ``tensorize_dataset(return_input_data=True)`` / ``predict`` receives from the host chain."""
    from buglab.representations.data import add_open_vocab_nodes_and_edges

    datapoint = msgpack.unpackb(sample_bytes, raw=False)
    add_open_vocab_nodes_and_edges(datapoint["graph"])
    return datapoint


class LazyDatapoint(Mapping):
    """A raw datapoint (``BugLabData``) whose msgpack bytes are unpacked only if somebody asks for a field the native
    decoder has not already produced.  Served without unpacking: ``target_fix_action_idx``, ``candidate_rewrites``,
    ``candidate_rewrite_metadata`` and ``graph["reference_nodes"]`` — everything ``predict`` and ``evaluate`` read
    (basemodel.py:240-346, evaluate.py:60-173).  Any other key, iteration, ``len`` or comparison unpacks the whole object
    once and extends its graph the way tensorisation does in the host chain (~5 ms for a 2 000-node graph, which is why it is
    not done up front), so every field equals what the reference-shaped path hands back.  Read-only; ``dict(dp)`` gives a
    plain mutable copy."""

    __slots__ = ("_raw", "_known", "_full")

    def __init__(self, raw: bytes, known: Dict[str, Any], graph_known: Optional[Dict[str, Any]] = None):
        self._raw = raw
        self._known = known
        self._full: Optional[Dict[str, Any]] = None
        if graph_known is not None:
            known["graph"] = LazyDatapoint._sub(self, "graph", graph_known)

    @classmethod
    def _sub(cls, parent: "LazyDatapoint", key: str, known: Dict[str, Any]) -> "LazyDatapoint":
        sub = cls.__new__(cls)
        sub._raw, sub._known, sub._full = (parent, key), known, None
        return sub

    def unpacked(self) -> Dict[str, Any]:
        """The fully unpacked object, cached: exactly what the host chain hands back next to a prediction, i.e. the stored
        object AFTER tensorisation appended the open-vocabulary subtoken nodes and ``HasSubtoken`` edges to its graph
        (the reference mutates its input there, data.py:97-121 via :139-167)."""
        if self._full is None:
            raw = self._raw
            self._full = raw[0].unpacked()[raw[1]] if isinstance(raw, tuple) else _unpack_like_the_host_chain(raw)
        return self._full

    def __getitem__(self, key):
        if self._full is None and key in self._known:
            return self._known[key]
        return self.unpacked()[key]

    def __iter__(self):
        return iter(self.unpacked())

    def __len__(self) -> int:
        return len(self.unpacked())

    def __repr__(self) -> str:
        return f"LazyDatapoint({'unpacked' if self._full is not None else 'packed'}, known={sorted(self._known)})"


class _ChunkBuffers:
    """Per-thread set of ``bl_sample`` handles + views for ``bl_sample_decode_many`` (one chunk of a shard per call)."""

    def __init__(self, capacity: int):
        self.capacity = capacity
        self.samples = [_SampleBuffer() for _ in range(capacity)]
        self.handles = (c_ptr * capacity)(*[b.handle for b in self.samples])
        self.views = (SampleView * capacity)()
        self.indices = (c_i64 * capacity)()


class NativeShardTensorizer:
    """``shard file -> BaseTensorizedBugLabGnn`` for a ``GnnBugLabModel`` whose metadata is finalised."""

    def __init__(self, model):
        lib()
        self._model = model
        gnn_model = model.gnn_model
        node_model = gnn_model.node_representation_model
        self._gnn_model = gnn_model
        self._tokenizer = Tokenizer(node_model.vocabulary, node_model.splitting_kind, node_model.max_num_subtokens)
        self._splitting_kind = node_model.splitting_kind
        names = [n.encode("utf-8") for n in gnn_model.edge_types]
        self._edge_names = (ctypes.c_char_p * max(1, len(names)))(*names)
        self._num_edge_types = len(names)
        self._local = threading.local()
        self._stats_lock = threading.Lock()
        self.num_native = 0   # samples produced by the native path / handed to the host path (statistics)
        self.num_host = 0

    # ---- one sample -----------------------------------------------------------------------------
    def _buffer(self) -> _SampleBuffer:
        buf = getattr(self._local, "buf", None)
        if buf is None:
            buf = self._local.buf = _SampleBuffer()
        return buf

    def tensorize_object(self, shard: Shard, index: int):
        """Tensorised sample ``index`` of ``shard``; ``None`` for nil objects and for samples the model drops."""
        buf = self._buffer()
        v = buf.view
        check(lib().bl_sample_decode(shard._h, index, self._tokenizer._h, self._edge_names, self._num_edge_types,
                                     buf.handle, ctypes.byref(v)), "decode")
        return self._from_view(v)

    NO_DATAPOINTS, FULL_DATAPOINTS, LAZY_DATAPOINTS = 0, 1, 2

    def _from_view(self, v: SampleView, datapoints: int = 0):
        """The ``BaseTensorizedBugLabGnn`` tuple of one decoded sample (``None``: nil object, or dropped by the model).
        With ``datapoints`` != 0 the result is ``(tuple or None, raw datapoint)``: the unpacked dict (``FULL_DATAPOINTS``) or
        a :class:`LazyDatapoint` over a copy of the sample's bytes (``LAZY_DATAPOINTS``) — what ``predict`` hands back."""
        if v.status == SAMPLE_NIL:
            return None if not datapoints else (None, None)
        if v.status == SAMPLE_NEEDS_HOST:
            with self._stats_lock:
                self.num_host += 1
            datapoint = msgpack.unpackb(ctypes.string_at(v.raw, v.raw_len), raw=False)
            tensorized = self._model.tensorize(datapoint)
            return tensorized if not datapoints else (tensorized, datapoint)
        with self._stats_lock:
            self.num_native += 1
        model, gnn_model = self._model, self._gnn_model

        # the three small candidate-rewrite fields are decoded by the host from their byte ranges (a few hundred bytes; the
        # sample's full raw bytes are never copied)
        raw = v.raw
        rewrites = msgpack.unpackb(ctypes.string_at(raw + v.rewrites_off, v.rewrites_len), raw=False)
        metadata = msgpack.unpackb(ctypes.string_at(raw + v.metadata_off, v.metadata_len), raw=False)
        logprobs = None
        has_logprobs = v.logprobs_len > 0
        if has_logprobs:
            assert not model._tensorize_only_at_target_location_rewrites
            logprobs = msgpack.unpackb(ctypes.string_at(raw + v.logprobs_off, v.logprobs_len), raw=False)

        reference_array = _as_array(v.reference_nodes, v.num_reference_nodes, np.int32)
        reference_nodes = reference_array.tolist()
        target_action = int(v.target_fix_action_idx) if v.has_target else None
        # np.unique over the int64 array np.unique would build from the list of Python ints (data.py:141): same values, same
        # dtypes, without the per-element conversion
        candidate_nodes, inverse = np.unique(reference_array.astype(np.int64), return_inverse=True)
        target_node_idx = None if target_action is None else inverse[target_action]

        call_args: Dict[int, List[int]] = defaultdict(list)
        pairs = _as_array(v.call_args, 2 * v.num_call_args, np.int32).tolist()
        for k in range(0, len(pairs), 2):
            call_args[pairs[k]].append(pairs[k + 1])
        rewrite_data = model._compute_rewrite_data_from(reference_nodes, rewrites, metadata, target_action, call_args,
                                                        candidate_nodes)
        refs: Dict[str, Any] = {"candidate_nodes": candidate_nodes}
        model._add_rewrite_reference_nodes(refs, rewrite_data)

        n, T = v.num_nodes, v.max_subtokens
        ids = _as_array(v.node_ids, n * T, np.int32).reshape(n, T)
        lens = _as_array(v.node_lens, n, np.int32)
        offsets = _as_array(v.edge_offsets, self._num_edge_types + 1, np.int64)
        src = _as_array(v.edge_src, int(offsets[-1]) if offsets.size else 0, np.int32)
        tgt = _as_array(v.edge_tgt, int(offsets[-1]) if offsets.size else 0, np.int32)
        forward = [(src[offsets[k]: offsets[k + 1]], tgt[offsets[k]: offsets[k + 1]]) for k in range(self._num_edge_types)]
        tensorized_graph = gnn_model.tensorize_arrays(n, (ids, lens), forward, refs)
        tensorized = None if tensorized_graph is None else \
            model._assemble_tensorized(tensorized_graph, target_node_idx, rewrite_data, logprobs)
        if not datapoints:
            return tensorized
        if tensorized is None:
            return None, None
        sample_bytes = ctypes.string_at(raw, v.raw_len)  # the shard (owner of `raw`) may be closed before the consumer looks
        if datapoints == self.FULL_DATAPOINTS:
            return tensorized, _unpack_like_the_host_chain(sample_bytes)
        known = {"target_fix_action_idx": target_action, "candidate_rewrites": rewrites, "candidate_rewrite_metadata": metadata}
        if has_logprobs:
            known["candidate_rewrite_logprobs"] = logprobs
        return tensorized, LazyDatapoint(sample_bytes, known, {"reference_nodes": reference_nodes})

    # ---- whole shards ---------------------------------------------------------------------------
    _DROPPED = object()  # a sample the model declined (too many nodes): counted by the loader, not yielded
    CHUNK = 32           # samples per work item: files are inflated whole, then decoded chunk by chunk by the workers
    DECODE_WORKERS = 2   # threads decoding chunks (the rest of ``num_threads`` only inflate files ahead)

    @staticmethod
    def _indices(shard: Shard, rank: int, world_size: int) -> List[int]:
        n = len(shard)
        return list(range(n)) if world_size <= 1 else list(range(rank, n, world_size))

    def _decode_chunk(self, shard: Shard, indices: Sequence[int], datapoints: int = 0) -> List:
        """Tensorised samples of one chunk of a shard, in order (nil objects skipped, dropped samples as ``_DROPPED``);
        with ``datapoints`` != 0 the items are ``(tensorised, raw datapoint)`` pairs (see :meth:`_from_view`)."""
        bufs = getattr(self._local, "chunk", None)
        if bufs is None:
            bufs = self._local.chunk = _ChunkBuffers(self.CHUNK)
        out: List = []
        for start in range(0, len(indices), bufs.capacity):
            part = indices[start: start + bufs.capacity]
            n = len(part)
            bufs.indices[:n] = part
            # ONE native call (outside the GIL) for the whole chunk, then the host-side assembly of its samples
            check(lib().bl_sample_decode_many(shard._h, bufs.indices, n, self._tokenizer._h, self._edge_names,
                                              self._num_edge_types, bufs.handles, bufs.views), "decode")
            for k in range(n):
                v = bufs.views[k]
                if v.status == SAMPLE_NIL:
                    continue
                t = self._from_view(v, datapoints)
                if datapoints:
                    out.append(self._DROPPED if t[0] is None else t)
                else:
                    out.append(self._DROPPED if t is None else t)
        return out

    @staticmethod
    def _open(path: str) -> Optional[Shard]:
        try:
            return Shard(path)
        except RuntimeError as e:  # unreadable shard: skipped like the reference does (msgpackutils.py:44-45)
            print(f"Error loading {path}: {e}.")
            return None

    @staticmethod
    def _finish(shard: Shard) -> None:
        if shard.status != 0:
            print(f"Error loading {shard.path}: {lib().bl_shards_error_string(shard.status).decode()}.")
        shard.close()

    def _shard_items(self, path: str, rank: int, world_size: int) -> List:
        with Shard(path) as shard:
            out = self._decode_chunk(shard, self._indices(shard, rank, world_size))
            if shard.status != 0:
                print(f"Error loading {path}: {lib().bl_shards_error_string(shard.status).decode()}.")
        return out

    def tensorize_shard(self, path: str, rank: int = 0, world_size: int = 1) -> Iterator:
        """All tensorised samples of one file, in file order (``rank``/``world_size``: element-level sharding used when
        there are fewer files than ranks, msgpackutils.load_all_msgpack_l_gz)."""
        for t in self._shard_items(path, rank, world_size):
            if t is not self._DROPPED:
                yield t

    def _chunks_sequential(self, paths: Sequence[str], shard_args: Tuple[int, int], datapoints: int = 0,
                           work=None, select=None) -> Iterator[List]:
        """``work(shard, indices)`` per chunk (default: :meth:`_decode_chunk`); ``select(shard, indices)`` may narrow a file's
        object indices before they are chunked, or return None to end the pass (element budgets)."""
        work = work or (lambda shard, indices: self._decode_chunk(shard, indices, datapoints))
        for path in paths:
            shard = self._open(path)
            if shard is None:
                continue
            try:
                indices = self._indices(shard, *shard_args)
                if select is not None:
                    indices = select(shard, indices)
                    if indices is None:
                        return
                for c in range(0, len(indices), self.CHUNK):
                    yield work(shard, indices[c: c + self.CHUNK])
            finally:
                self._finish(shard)

    def _chunks_parallel(self, paths: Sequence[str], shard_args: Tuple[int, int], num_threads: int,
                         datapoints: int = 0, work=None, select=None) -> Iterator[List]:
        """Chunk results in file order.  Two kinds of work items: *open* (inflate + index one file, outside the GIL, on up to
        ``num_threads`` workers, a few files ahead of the one being decoded) and *decode* (``work`` on one chunk of ``CHUNK``
        objects of an open file, on ``DECODE_WORKERS`` workers) — so a single 500-graph shard keeps the workers busy and its
        first samples are available as soon as the file is inflated, instead of after the whole file has been decoded by one
        thread.  A shard is shared read-only by the workers (``include/buglab_shards.h``: immutable after creation) and closed
        once its last chunk has been consumed.  Look-ahead is bounded: ``num_threads`` inflated files, ``2 * DECODE_WORKERS``
        chunks.  ``work`` / ``select``: see :meth:`_chunks_sequential`."""
        from collections import deque

        work = work or (lambda shard, indices: self._decode_chunk(shard, indices, datapoints))
        # the decode items end with host-side assembly under the GIL (~0.25 ms per sample): beyond DECODE_WORKERS threads they
        # only contend with each other and with the consumer; inflating (zlib, outside the GIL, ~2/3 of the work) scales
        decode_workers = min(num_threads, self.DECODE_WORKERS)
        pool = ThreadPoolExecutor(max_workers=decode_workers)
        open_pool = ThreadPoolExecutor(max_workers=num_threads)
        open_ahead = num_threads
        max_inflight = 2 * decode_workers
        opened: List[Shard] = []  # every shard this call opened and has not closed yet (closed in `finally` on early exit)
        opening: "deque" = deque()  # files being inflated ahead of the one being decoded (futures)

        def work_items() -> Iterator[Tuple[str, Any, Any]]:
            it = iter(paths)

            def open_next() -> None:
                p = next(it, None)
                if p is not None:
                    opening.append(open_pool.submit(self._open, p))

            for _ in range(open_ahead):
                open_next()
            while opening:
                shard = opening.popleft().result()
                open_next()
                if shard is None:
                    continue
                opened.append(shard)
                indices = self._indices(shard, *shard_args)
                if select is not None:
                    indices = select(shard, indices)
                    if indices is None:
                        yield "end", shard, None
                        return
                for c in range(0, len(indices), self.CHUNK):
                    yield "chunk", shard, indices[c: c + self.CHUNK]
                yield "end", shard, None

        items = work_items()
        window: "deque" = deque()

        def fill() -> None:
            while len(window) < max_inflight:
                item = next(items, None)
                if item is None:
                    return
                kind, shard, indices = item
                window.append((kind, shard, pool.submit(work, shard, indices) if kind == "chunk" else None))

        try:
            fill()
            while window:
                kind, shard, future = window.popleft()
                if kind == "chunk":
                    result = future.result()
                    fill()
                    yield result
                else:
                    opened.remove(shard)
                    self._finish(shard)
                    fill()
        finally:
            # a consumer that stops early (element limit, closed generator): no worker may still be decoding from a shard
            # when it is closed
            pool.shutdown(wait=True, cancel_futures=True)
            open_pool.shutdown(wait=True, cancel_futures=True)
            for future in opening:  # files inflated ahead whose turn never came
                if not future.cancelled() and future.exception() is None and future.result() is not None:
                    future.result().close()
            for shard in opened:
                shard.close()

    def tensorize_files(self, paths: Iterable[str], num_threads: Optional[int] = None, rank: int = 0, world_size: int = 1,
                        element_sharding: bool = False, limit_num_elements: Optional[int] = None, datapoints: int = 0
                        ) -> Iterator[Tuple[Any, Any]]:
        """``(tensorised, None)`` pairs — the shape ``tensorize_dataset`` yields — over many shard files, in file order.
        Files are inflated and their samples decoded by ``num_threads`` workers (the native calls run outside the GIL),
        chunk by chunk (:meth:`_chunks_parallel`).  ``limit_num_elements`` counts non-nil file elements the way
        load_all_msgpack_l_gz does (it stops after the element that EXCEEDS the limit, as the reference's loop does).
        ``datapoints``: ``FULL_DATAPOINTS`` / ``LAZY_DATAPOINTS`` pair every sample with its raw datapoint
        (``tensorize_dataset(return_input_data=True)``: what ``predict`` needs) instead of ``None``."""
        paths = list(paths)
        if num_threads is None:
            num_threads = max(1, min(4, (os.cpu_count() or 2) - 1))  # the host-side assembly holds the GIL: more workers only contend
        shard_args = (rank, world_size) if element_sharding else (0, 1)
        chunks = self._chunks_sequential(paths, shard_args, datapoints) if num_threads <= 1 else \
            self._chunks_parallel(paths, shard_args, num_threads, datapoints)
        num_seen = 0
        try:
            for items in chunks:
                for t in items:
                    num_seen += 1
                    if t is not self._DROPPED:
                        yield t if datapoints else (t, None)
                    if limit_num_elements is not None and num_seen > limit_num_elements:
                        return
        finally:
            chunks.close()


class NativeMetadataPass(NativeShardTensorizer):
    """The metadata pass of a graph model over shard files — ``model.compute_metadata(data)`` without building a Python
    object per graph (reference: ptgnn ``compute_metadata`` -> GnnBugLabModel.update_metadata_from, buglab/models/gnn.py:350-358
    -> BugLabData.as_graph_data -> StrElementRepresentationModel / GraphNeuralNetworkModel ``update_metadata_from``).
    Files are inflated and counted by the same worker scheme as the tensoriser; every worker thread counts into its own
    native accumulator (``bl_metadata``), the accumulators are merged at the end (counts and name sets are order
    independent).  Samples the native path declines are unpacked and go through ``model.update_metadata_from``."""

    def __init__(self, model):
        lib()
        self._model = model
        node_model = model.gnn_model.node_representation_model
        self._tokenizer = Tokenizer(None, node_model.splitting_kind, node_model.max_num_subtokens or 1)
        self._local = threading.local()
        self._stats_lock = threading.Lock()
        self._accumulators: List[c_ptr] = []
        self.num_native = 0
        self.num_host = 0

    def _accumulator(self) -> c_ptr:
        acc = getattr(self._local, "metadata", None)
        if acc is None:
            acc = c_ptr()
            check(lib().bl_metadata_create(ctypes.byref(acc)), "metadata")
            self._local.metadata = acc
            with self._stats_lock:
                self._accumulators.append(acc)
        return acc

    def _count_chunk(self, shard: Shard, indices: Sequence[int]) -> List[bytes]:
        """Adds one chunk to this thread's accumulator; returns the raw bytes of the samples left to the host."""
        n = len(indices)
        idx = (c_i64 * n)(*indices)
        declined = (c_i32 * max(1, n))()
        num_declined = c_i32()
        check(lib().bl_metadata_add(self._accumulator(), shard._h, idx, n, self._tokenizer._h, self._buffer().handle,
                                    declined, ctypes.byref(num_declined)), "metadata")
        return [shard.object_bytes(indices[declined[k]]) for k in range(num_declined.value)]

    def run(self, paths: Iterable[str], num_threads: Optional[int] = None, rank: int = 0, world_size: int = 1,
            element_sharding: bool = False, limit_num_elements: Optional[int] = None) -> Tuple[Dict[str, int], List[str], int]:
        """``(token counts, edge-type names, number of samples)`` over the files, with the file selection, rank sharding and
        element limit semantics of ``load_all_msgpack_l_gz`` (the pass covers ``limit + 1`` elements when there are more)."""
        paths = list(paths)
        if num_threads is None:
            num_threads = max(1, min(4, (os.cpu_count() or 2) - 1))
        shard_args = (rank, world_size) if element_sharding else (0, 1)
        budget = [None if limit_num_elements is None else limit_num_elements + 1]

        def select(shard: Shard, indices: List[int]) -> Optional[List[int]]:
            if budget[0] is not None and budget[0] <= 0:
                return None
            not_nil = set(shard.non_nil_indices())
            chosen = [i for i in indices if i in not_nil]
            if budget[0] is not None:
                chosen = chosen[: budget[0]]
                budget[0] -= len(chosen)
            return chosen

        chunks = self._chunks_sequential(paths, shard_args, work=self._count_chunk, select=select) if num_threads <= 1 else \
            self._chunks_parallel(paths, shard_args, num_threads, work=self._count_chunk, select=select)
        model = self._model
        completed = False
        try:
            for declined in chunks:
                for sample_bytes in declined:  # the reference-shaped path for what the decoder does not take
                    self.num_host += 1
                    model.update_metadata_from(msgpack.unpackb(sample_bytes, raw=False))
            completed = True
        finally:
            chunks.close()  # waits for the workers: no thread is counting into an accumulator after this line
            if not completed:  # a malformed sample raised in the host pass, as it does there: nothing is merged
                for acc in self._accumulators:
                    lib().bl_metadata_destroy(acc)
                self._accumulators, self._local = [], threading.local()
        counts: Dict[str, int] = {}
        edge_types: Dict[str, int] = {}
        num_samples = 0
        for acc in self._accumulators:
            num_samples += int(lib().bl_metadata_num_samples(acc))
            for which, into in ((0, counts), (1, edge_types)):
                blob_bytes = c_i64()
                n = int(lib().bl_metadata_size(acc, which, ctypes.byref(blob_bytes)))
                blob = np.empty(max(1, blob_bytes.value), dtype=np.uint8)
                offsets = np.empty(n + 1, dtype=np.int64)
                values = np.empty(max(1, n), dtype=np.int64)
                check(lib().bl_metadata_export(acc, which, blob.ctypes.data, offsets.ctypes.data, values.ctypes.data), "export")
                data = blob.tobytes()
                bounds = offsets.tolist()
                for k, count in enumerate(values[:n].tolist()):
                    token = data[bounds[k]: bounds[k + 1]].decode("utf-8", "surrogatepass")
                    into[token] = into.get(token, 0) + count
            lib().bl_metadata_destroy(acc)
        self._accumulators = []
        self._local = threading.local()
        self.num_native = num_samples
        return counts, sorted(edge_types), num_samples + self.num_host


class ShardDataset:
    """A directory of ``.msgpack.l.gz`` shards as the trainer's data source.

    Iterating it yields raw datapoints through the host decoder (metadata pass, ``predict``); ``tensorized(model)`` yields
    ``(tensorised, None)`` pairs straight from the native decoder — ``ModelTrainer.train`` picks that up when present.
    File selection, rank sharding and the element limit are those of ``load_all_msgpack_l_gz``."""

    def __init__(self, data_path, shuffle: bool = False, take_only_first_n_files: Optional[int] = None,
                 limit_num_yielded_elements: Optional[int] = None, rank: int = 0, world_size: int = 1,
                 num_threads: Optional[int] = None, lazy_input_data: bool = False):
        # lazy_input_data: when a consumer asks for the raw datapoints next to the tensorised samples (``predict``), hand out
        # read-only LazyDatapoint views instead of unpacked dicts (for consumers that read only the fields predict/evaluate use)
        self.lazy_input_data = bool(lazy_input_data)
        self._path, self._shuffle, self._first_n = data_path, shuffle, take_only_first_n_files
        self._limit, self._rank, self._world, self._threads = limit_num_yielded_elements, rank, world_size, num_threads
        self._tensorizer: Optional[NativeShardTensorizer] = None
        self._tensorizer_model = None

    def __iter__(self):
        from buglab.utils.msgpackutils import load_all_msgpack_l_gz

        return load_all_msgpack_l_gz(self._path, shuffle=self._shuffle, take_only_first_n_files=self._first_n,
                                     limit_num_yielded_elements=self._limit, rank=self._rank, world_size=self._world)

    def tensorized(self, model, return_input_data: bool = False, lazy_input_data: bool = False) -> Iterator[Tuple[Any, Any]]:
        """``(tensorised, raw datapoint or None)`` pairs, the shape of ``model.tensorize_dataset``.  ``return_input_data``
        pairs every sample with its raw datapoint (``predict``); with ``lazy_input_data`` that datapoint is a read-only
        :class:`LazyDatapoint` that is only unpacked if a field the decoder did not produce is read (``evaluate``)."""
        from buglab.utils.msgpackutils import select_shard_files

        if not hasattr(model, "gnn_model"):
            # the native tensoriser produces the GRAPH models' per-sample tensors; other model families (the sequence
            # models project graphs onto token sequences in host code) get the reference-shaped chain, said out loud
            LOGGER.info("%s is not a graph model: shards are decoded and tensorised by the host-language path.",
                        type(model).__name__)
            return model.tensorize_dataset(iter(self), return_input_data=return_input_data,
                                           parallelize=(self._threads or 2) > 1)
        if self._tensorizer is None or self._tensorizer_model is not model:
            self._tensorizer, self._tensorizer_model = NativeShardTensorizer(model), model
        files, shard_elements = select_shard_files(self._path, self._shuffle, self._first_n, self._rank, self._world)
        paths = [f.to_local_path().path for f in files]
        mode = NativeShardTensorizer.NO_DATAPOINTS if not return_input_data else \
            (NativeShardTensorizer.LAZY_DATAPOINTS if lazy_input_data else NativeShardTensorizer.FULL_DATAPOINTS)
        return self._tensorizer.tensorize_files(paths, self._threads, self._rank, self._world, shard_elements, self._limit,
                                                mode)

    def update_model_metadata(self, model) -> bool:
        """The metadata pass (``model.compute_metadata``) over this data source in native code.  Returns False — nothing
        done, the caller iterates the raw datapoints instead — for models whose metadata the native pass does not cover
        (anything but a graph model with a string node-representation model, i.e. the sequence models)."""
        from buglab.utils.msgpackutils import select_shard_files

        gnn_model = getattr(model, "gnn_model", None)
        node_model = getattr(gnn_model, "node_representation_model", None)
        if gnn_model is None or not hasattr(node_model, "update_metadata_from_token_counts") or \
                not hasattr(gnn_model, "update_metadata_from_edge_types") or type(model).update_metadata_from is not \
                _graph_model_update_metadata_from(model):
            return False
        files, shard_elements = select_shard_files(self._path, self._shuffle, self._first_n, self._rank, self._world)
        paths = [f.to_local_path().path for f in files]
        counts, edge_types, _ = NativeMetadataPass(model).run(paths, self._threads, self._rank, self._world, shard_elements,
                                                              self._limit)
        node_model.update_metadata_from_token_counts(counts)
        gnn_model.update_metadata_from_edge_types(edge_types)
        return True

    @property
    def tensorizer(self) -> Optional[NativeShardTensorizer]:
        return self._tensorizer


def _graph_model_update_metadata_from(model):
    """``GnnBugLabModel.update_metadata_from`` — the one implementation whose effect the native pass reproduces (a subclass
    that overrides it collects something else and must see the raw datapoints)."""
    from buglab.models.gnn import GnnBugLabModel

    return GnnBugLabModel.update_metadata_from
