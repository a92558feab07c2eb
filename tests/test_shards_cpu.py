"""Native shard decoder (include/buglab_shards.h) against the host-language path.

The checker is the reference-shaped Python chain load_msgpack_l_gz -> BugLabData.as_graph_data -> GnnBugLabModel.tensorize
(itself pinned by tests/golden).  The bar is bit-identity: same arrays, same dtypes, same Python scalars, same order."""
import copy
import gzip
import io
import os
import random
import re
import sys

import msgpack
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

from buglab_b200 import shards  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def assert_same(a, b, where="sample"):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        assert isinstance(a, np.ndarray) and isinstance(b, np.ndarray), (where, type(a), type(b))
        assert a.dtype == b.dtype and a.shape == b.shape, (where, a.dtype, b.dtype, a.shape, b.shape)
        assert np.array_equal(a, b), where
    elif hasattr(a, "_fields"):
        assert type(a) is type(b), where
        for f in a._fields:
            assert_same(getattr(a, f), getattr(b, f), f"{where}.{f}")
    elif isinstance(a, dict):
        assert isinstance(b, dict) and list(a.keys()) == list(b.keys()), where
        for k in a:
            assert_same(a[k], b[k], f"{where}.{k}")
    elif isinstance(a, (list, tuple)):
        assert isinstance(b, (list, tuple)) and len(a) == len(b), (where, a, b)
        for i, (x, y) in enumerate(zip(a, b)):
            assert_same(x, y, f"{where}[{i}]")
    else:
        assert type(a) is type(b) and a == b, (where, a, b)


@pytest.fixture(scope="module")
def model():
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    m, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 16}, Path("/tmp/_buglab_shards_test.pkl.gz"))
    m.compute_metadata(SyntheticBugLabGenerator(seed=12345, mean_nodes=200, min_nodes=40).samples(48))
    return m


def host_tensorize_file(model, path):
    from buglab.utils.msgpackutils import load_msgpack_l_gz

    return [model.tensorize(dp) for dp in load_msgpack_l_gz(path) if dp is not None]


def write_objects(path, objects, **packer_kwargs):
    packer = msgpack.Packer(use_bin_type=True, **packer_kwargs)
    with gzip.GzipFile(path, "wb") as f:
        for o in objects:
            f.write(packer.pack(o))


# ---------------------------------------------------------------------------------------------- C ABI surface
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "buglab_shards.h")).read()
    declared = set(re.findall(r"\b(bl_[a-z0-9_]+)\s*\(", header))
    declared -= {"bl_sample_view"}
    assert len(declared) >= 16
    lib = shards.lib()
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in buglab_shards.h but not exported"
        assert name in shards._SIGNATURES, f"{name} has no ctypes signature"
    assert lib.bl_shards_version() == 1
    assert lib.bl_shards_error_string(2) == b"corrupt or truncated gzip stream"


# ---------------------------------------------------------------------------------------------- CPython set order
def test_set_iteration_order_matches_the_interpreter():
    rng = random.Random(3)
    for trial in range(400):
        n = rng.choice([0, 1, 2, 5, 8, 9, 21, 85, 341, 1500, 6000])
        hi = rng.choice([8, 64, 1000, 50_000, 10 ** 9, 2 ** 50])
        if trial % 2:
            start = rng.randrange(hi)
            values = []
            for i in range(n):  # NextToken-like chains (a, a+1), (a+1, a+2), ...
                values += [start + i, start + i + rng.choice([1, 1, 1, 3])]
        else:
            values = [rng.randrange(hi) for _ in range(n)]
        expected = set()
        for v in values:
            expected.add(v)
        assert shards.pyset_iteration_order(values) == list(expected)
    with pytest.raises(ValueError):
        shards.pyset_iteration_order([3, -1])


# ---------------------------------------------------------------------------------------------- tokeniser
LABELS = ["", "_", "___", "a", "A", "aB", "ABc", "AB", "HTTPServer", "fooBar_baz2", "__init__", "CamelCASE", "iOS",
          "x+y", " ", "a1B2c3", "snake_case_name", "UPPER_CASE", "Call", "Name", "<=", "'a string literal'", "f00Bar",
          "naïve", "日本語", "caféAu_lait", "\U0001F600smile", "aßb", "x" * 300, "aB" * 200,
          "Élan", "ΑΒΓ", "STRASSEẞ", "İstanbul", "mixedΣigma"]


@pytest.mark.parametrize("kind", ["subtoken", "token"])
def test_tokenizer_matches_host_ids(kind):
    from dpu_utils.mlutils import Vocabulary
    from ptgnn.neuralmodels.embeddings.strelementrepresentationmodel import StrElementRepresentationModel

    node_model = StrElementRepresentationModel(token_splitting=kind, embedding_size=8, vocabulary_size=500,
                                               min_freq_threshold=1, max_num_subtokens=6, subtoken_combination="max")
    for label in LABELS[::2] * 2:
        node_model.update_metadata_from(label)
    node_model.finalize_metadata()
    vocab: Vocabulary = node_model.vocabulary
    tok = shards.Tokenizer(vocab, kind, node_model.max_num_subtokens)
    needs_host = 0
    for label in LABELS:
        got = tok.ids(label)
        if got is None:
            needs_host += 1
            # only labels with case-variant or letter/digit non-ASCII code points are declined
            assert any(ord(c) >= 0x80 and (c.lower() != c or c.isalnum()) for c in label), label
            continue
        assert got == tuple(node_model._ids_of(label)), label
    assert needs_host == sum(any(ord(c) >= 0x80 and (c.lower() != c or c.isalnum()) for c in label) for label in LABELS)
    assert needs_host >= 5


def test_tokenizer_matches_host_on_random_identifiers():
    """20 000 random labels over an alphabet that stresses the splitter (case runs, digits, underscores, punctuation,
    non-ASCII letters, digits and symbols): native ids == host ids, or the native side declines — exactly when the label
    holds a non-ASCII code point that str.lower() changes or that is a letter/digit for the splitter's classes."""
    from ptgnn.neuralmodels.embeddings.strelementrepresentationmodel import StrElementRepresentationModel

    rng = random.Random(99)
    alphabet = list("abcxyzABCXYZ019__--+.( ") + ["é", "ß", "日", "Σ", "É", "ǅ", "😀", "٣", "→", "…"]
    labels = ["".join(rng.choice(alphabet) for _ in range(rng.randrange(0, 14))) for _ in range(20000)]
    node_model = StrElementRepresentationModel(token_splitting="subtoken", embedding_size=8, vocabulary_size=3000,
                                               min_freq_threshold=2, max_num_subtokens=5, subtoken_combination="max")
    for label in labels[:6000]:
        node_model.update_metadata_from(label)
    node_model.finalize_metadata()
    tok = shards.Tokenizer(node_model.vocabulary, "subtoken", node_model.max_num_subtokens)
    declined = 0
    for label in labels:
        got = tok.ids(label)
        if got is None:
            declined += 1
            assert any(ord(c) >= 0x80 and (c.lower() != c or c.isalnum()) for c in label), repr(label)
        else:
            assert got == tuple(node_model._ids_of(label)), repr(label)
            # never handles what it should decline
            assert not any(ord(c) >= 0x80 and (c.lower() != c or c.isalnum()) for c in label), repr(label)
    assert 2000 < declined < 16000


def test_vocabulary_without_unk_routes_misses_to_host():
    from dpu_utils.mlutils import Vocabulary

    vocab = Vocabulary(add_unk=False, add_pad=True)
    vocab.add_or_get_id("foo")
    tok = shards.Tokenizer(vocab, "subtoken", 4)
    assert tok.ids("foo_Foo") == (vocab.get_id_or_unk("foo"),) * 2
    assert tok.ids("foo_bar") is None


# ---------------------------------------------------------------------------------------------- whole samples
def test_synthetic_shard_is_bit_identical(model, tmp_path):
    from buglab_b200.synthetic import write_shards

    path = write_shards(str(tmp_path), 1, 24, seed=5, mean_nodes=300, min_nodes=30)[0]
    expected = host_tensorize_file(model, path)
    tensorizer = shards.NativeShardTensorizer(model)
    got = list(tensorizer.tensorize_shard(path))
    assert len(got) == len(expected) == 24
    assert tensorizer.num_native == 24 and tensorizer.num_host == 0
    for i, (a, b) in enumerate(zip(expected, got)):
        assert_same(a, b, f"sample{i}")


def test_golden_reference_samples_are_bit_identical(model):
    """The samples the golden vectors were generated from (tests/golden/make_golden.py), incl. the selector ones that
    carry candidate_rewrite_logprobs."""
    tensorizer = shards.NativeShardTensorizer(model)
    path = os.path.join(GOLDEN, "samples.msgpack.l.gz")
    for a, b in zip(host_tensorize_file(model, path), tensorizer.tensorize_shard(path)):
        assert_same(a, b)
    path = os.path.join(GOLDEN, "selector_samples.msgpack.l.gz")
    with model._tensorize_all_location_rewrites():
        expected = host_tensorize_file(model, path)
        got = list(tensorizer.tensorize_shard(path))
    assert len(expected) == len(got) > 0
    for a, b in zip(expected, got):
        assert_same(a, b)
        assert (a.rewrite_logprobs is None) == (b.rewrite_logprobs is None)
    assert tensorizer.num_host == 0


def _variants(base):
    """Hand-made irregular samples: each must come out exactly as the host path makes it (natively or by deferring)."""
    def v(mutate):
        s = copy.deepcopy(base)
        mutate(s)
        return s

    g = "graph"
    out = {
        "plain": v(lambda s: None),
        "no_next_token_keeps_file_has_subtoken": v(lambda s: (s[g]["edges"].pop("NextToken"),
                                                               s[g]["edges"].__setitem__("HasSubtoken", [[0, 1], [2, 3]]))),
        "stale_has_subtoken_is_replaced": v(lambda s: s[g]["edges"].__setitem__("HasSubtoken", [[0, 1]])),
        "empty_edge_lists": v(lambda s: [s[g]["edges"].__setitem__(k, []) for k in list(s[g]["edges"])
                                         if k not in ("NextToken", "Child")]),
        "unknown_edge_type": v(lambda s: s[g]["edges"].__setitem__("NeverSeenInMetadata", [[0, 1, "meta"], [1, 2]])),
        "edge_metadata_of_other_types": v(lambda s: s[g]["edges"].__setitem__(
            "NextToken", [[a, b, None] for a, b, *_ in s[g]["edges"]["NextToken"]])),
        "out_of_vocabulary_labels": v(lambda s: [s[g]["nodes"].__setitem__(i, f"zzQq{i}_neverSeen") for i in range(0, 12, 3)]),
        "non_ascii_symbols_stay_native": v(lambda s: s[g]["nodes"].__setitem__(1, "fooBar→baz…_qux😀")),
        "non_ascii_letters_go_to_host": v(lambda s: s[g]["nodes"].__setitem__(1, "caféCrème_日本")),
        "case_variant_unicode_goes_to_host": v(lambda s: s[g]["nodes"].__setitem__(1, "ÉlanVital")),
        "long_labels_str8_str16": v(lambda s: (s[g]["nodes"].__setitem__(2, "longName" * 9),
                                               s[g]["nodes"].__setitem__(3, "x" * 70000))),
        "extra_keys_everywhere": v(lambda s: (s.__setitem__("unheard_of", {"a": [1, 2.5, None, True, b"bytes"]}),
                                              s[g].__setitem__("more", [[], {}, -1, 2 ** 40, -2 ** 40, 1.5]))),
        "no_bug": v(lambda s: s.__setitem__("target_fix_action_idx", None)),
    }
    return out


@pytest.fixture(scope="module")
def base_sample():
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    return SyntheticBugLabGenerator(seed=77, mean_nodes=120, min_nodes=60).sample()


def test_irregular_samples_match_host_path(model, base_sample, tmp_path):
    variants = _variants(base_sample)
    path = str(tmp_path / "variants.msgpack.l.gz")
    # nil elements between the samples are skipped by both loaders (msgpackutils.py:38)
    write_objects(path, [None] + [x for s in variants.values() for x in (s, None)])
    expected = host_tensorize_file(model, path)
    tensorizer = shards.NativeShardTensorizer(model)
    got = list(tensorizer.tensorize_shard(path))
    assert len(expected) == len(got) == len(variants)
    for name, a, b in zip(variants, expected, got):
        assert_same(a, b, name)
    assert tensorizer.num_host == 2 and tensorizer.num_native == len(variants) - 2
    # the no-bug variant really exercised the nil target
    assert got[list(variants).index("no_bug")].target_location_node_idx is None


def test_host_path_exceptions_are_preserved(model, base_sample, tmp_path):
    """Samples the reference would crash on crash the same way here (the native side defers instead of guessing)."""
    broken = copy.deepcopy(base_sample)
    del broken["graph"]["edges"]["Child"]           # basemodel.py:84 indexes ["Child"]
    path = str(tmp_path / "broken.msgpack.l.gz")
    write_objects(path, [broken])
    tensorizer = shards.NativeShardTensorizer(model)
    with pytest.raises(KeyError):
        host_tensorize_file(model, path)
    with pytest.raises(KeyError):
        list(tensorizer.tensorize_shard(path))

    out_of_range = copy.deepcopy(base_sample)
    out_of_range["graph"]["edges"]["NextToken"].append([0, len(out_of_range["graph"]["nodes"]) + 10 ** 6])
    write_objects(path, [out_of_range])
    with pytest.raises(IndexError):
        host_tensorize_file(model, path)
    with pytest.raises(IndexError):
        list(tensorizer.tensorize_shard(path))


def test_oversized_graph_is_dropped_like_the_host_path(model, base_sample, tmp_path):
    path = str(tmp_path / "big.msgpack.l.gz")
    write_objects(path, [base_sample, base_sample])
    gnn_model = model.gnn_model
    saved = gnn_model.max_nodes_per_graph
    try:
        gnn_model.max_nodes_per_graph = 10
        assert host_tensorize_file(model, path) == [None, None]
        assert list(shards.NativeShardTensorizer(model).tensorize_shard(path)) == []
    finally:
        gnn_model.max_nodes_per_graph = saved


def mutate_sample(sample, rng):
    """One to three random structural mutations of a decoded sample (used by the differential test below)."""
    graph = sample["graph"]
    nodes, edges = graph["nodes"], graph["edges"]
    weird_labels = ["", "_", "ÉCOLE", "naïveBayes", "日本_語", "ΣΑΣ", "a" * 40, "HTTPServer2Go", "__x__", "\x00nul", "x y",
                    "İ", "ǅ", "ß", "tab\there", "😀", "Call", "call", "CALL_ME"]

    def random_edge_type():
        return rng.choice(list(edges))

    def m_label():
        nodes[rng.randrange(len(nodes))] = rng.choice(weird_labels)

    def m_label_type():
        nodes[rng.randrange(len(nodes))] = rng.choice([7, None, b"bytes", 1.5, ["x"]])

    def m_edge_value():
        lst = edges[random_edge_type()]
        if lst:
            e = list(lst[rng.randrange(len(lst))])
            e[rng.randrange(2)] = rng.choice([-1, -len(nodes), len(nodes), len(nodes) + 3, 2 ** 31, 2 ** 40, 0, 1.0, "3", None])
            lst[rng.randrange(len(lst))] = e

    def m_edge_arity():
        lst = edges[random_edge_type()]
        if lst:
            i = rng.randrange(len(lst))
            e = list(lst[i])
            lst[i] = rng.choice([e[:1], e[:2], e[:2] + ["args"], e[:2] + [b"args"], e[:2] + ["args", 1], e[:2] + [None], []])

    def m_drop_edge_type():
        edges.pop(random_edge_type())

    def m_edge_container():
        edges[random_edge_type()] = rng.choice([None, {}, "str", 5, [[0, 1]], []])

    def m_reference_nodes():
        refs = graph["reference_nodes"]
        if refs:
            refs[rng.randrange(len(refs))] = rng.choice([-1, 0, len(nodes) + 7, 2 ** 33, "1", None, 2.0])

    def m_target():
        sample["target_fix_action_idx"] = rng.choice([None, 0, -1, 10 ** 6, "0", 1.0, True])

    def m_drop_key():
        holder = rng.choice([sample, graph])
        holder.pop(rng.choice(list(holder)))

    def m_rewrites():
        key = rng.choice(["candidate_rewrites", "candidate_rewrite_metadata"])
        if sample.get(key):
            i = rng.randrange(len(sample[key]))
            sample[key][i] = rng.choice([None, [], ["only one"], ["ArgSwapRewriteScout", None], ["X", [0, 9]], 5])

    def m_call_label():
        child = edges.get("Child") or []
        marked = [e for e in child if len(e) == 3 and e[2] == "args"]
        if marked:
            nodes[marked[rng.randrange(len(marked))][0]] = rng.choice(["Call", "call", "Name"])

    def m_more_keys():
        sample[rng.choice(["zzz", "graph2", "candidate_rewrite_logprobs_x"])] = rng.choice([None, [1, 2], {"a": {"b": []}}])

    def m_next_token_gone():
        edges.pop("NextToken", None)

    moves = [m_label] * 4 + [m_label_type, m_edge_value, m_edge_value, m_edge_arity, m_edge_arity, m_drop_edge_type,
                             m_edge_container, m_reference_nodes, m_target, m_drop_key, m_rewrites, m_call_label,
                             m_call_label, m_more_keys, m_next_token_gone]
    for _ in range(rng.choice([1, 1, 2, 3])):
        try:
            rng.choice(moves)()
        except (KeyError, IndexError, TypeError, AttributeError, ValueError):
            pass  # an earlier mutation removed what this one wanted to touch
    return sample


def test_randomly_mutated_samples_behave_like_the_host_path(model):
    """Differential test: structurally mutated samples give bit-identical tensors, or raise the same exception type."""
    from buglab_b200.synthetic import SyntheticBugLabGenerator

    gen = SyntheticBugLabGenerator(seed=5, mean_nodes=60, min_nodes=30)
    bases = [gen.sample() for _ in range(4)]
    rng = random.Random(20210921)
    tensorizer = shards.NativeShardTensorizer(model)
    same = raised = 0
    for it in range(300):
        sample = mutate_sample(copy.deepcopy(rng.choice(bases)), rng)
        blob = msgpack.packb(sample, use_bin_type=True)
        try:
            expected, host_error = model.tensorize(msgpack.unpackb(blob, raw=False)), None
        except Exception as e:  # noqa: BLE001 - whatever the reference-shaped path raises is the expectation
            expected, host_error = None, e
        with shards.Shard(gz_bytes=gzip.compress(blob, 1)) as shard:
            assert len(shard) == 1
            try:
                got, native_error = tensorizer.tensorize_object(shard, 0), None
            except Exception as e:  # noqa: BLE001
                got, native_error = None, e
        if host_error is not None or native_error is not None:
            assert type(host_error) is type(native_error), (it, repr(host_error), repr(native_error))
            raised += 1
        else:
            assert_same(expected, got, f"mutation {it}")
            same += 1
    assert same > 100 and raised > 30 and tensorizer.num_native > 60 and tensorizer.num_host > 60


# ---------------------------------------------------------------------------------------------- container format
def test_multi_member_gzip_and_truncated_streams(model, base_sample, tmp_path):
    packer = msgpack.Packer(use_bin_type=True)
    one = packer.pack(base_sample)
    member = lambda payload: gzip.compress(payload)  # noqa: E731
    two_members = member(one + one) + member(one)
    with shards.Shard(gz_bytes=two_members) as shard:
        assert len(shard) == 3 and shard.status == 0
        assert shard.object_bytes(2) == one and shard.raw_bytes == 3 * len(one)
    # python's reader agrees on the concatenated-members semantics
    assert len(list(msgpack.Unpacker(gzip.GzipFile(fileobj=io.BytesIO(two_members)), raw=False))) == 3

    whole = member(one + one + one)
    with shards.Shard(gz_bytes=whole[: len(whole) - 12]) as shard:  # cut inside the deflate stream: no trailer, tail lost
        assert shard.status == 2 and 2 <= len(shard) <= 3
        assert shard.object_bytes(0) == one
    with shards.Shard(gz_bytes=member(one + one[: len(one) // 2])) as shard:  # complete gzip, incomplete last object
        assert len(shard) == 1 and shard.status == 4
    with shards.Shard(gz_bytes=member(one + b"\xc1" + one)) as shard:  # 0xc1 is never a valid msgpack type byte
        assert len(shard) == 1 and shard.status == 4
    with shards.Shard(gz_bytes=member(packer.pack({"k": "ok"}) + packer.pack({"bad": b"\xff\xfe".decode("latin1")})
                                      .replace("ÿþ".encode(), b"\xff\xfe\xfe\xfe"))) as shard:
        assert len(shard) == 1 and shard.status == 4          # invalid UTF-8 in a str, as raw=False rejects it
    with shards.Shard(gz_bytes=member(packer.pack({"k": 1}) + msgpack.packb({1: 2}))) as shard:
        assert len(shard) == 1 and shard.status == 4          # non-str map key (strict_map_key)
    with pytest.raises(RuntimeError):
        shards.Shard(path=str(tmp_path / "does_not_exist.msgpack.l.gz"))
    with shards.Shard(gz_bytes=b"this is not gzip") as shard:
        assert len(shard) == 0 and shard.status == 2


# ---------------------------------------------------------------------------------------------- dataset level
def test_shard_dataset_matches_host_loader(model, tmp_path):
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.synthetic import write_shards
    from dpu_utils.utils import RichPath

    write_shards(str(tmp_path / "d"), 5, 7, seed=9, mean_nodes=120, min_nodes=30)
    rich = RichPath.create(str(tmp_path / "d"))

    def host(**kw):
        return [t for t in (model.tensorize(dp) for dp in load_all_msgpack_l_gz(rich, **kw)) if t is not None]

    def native(threads, **kw):
        ds = shards.ShardDataset(rich, num_threads=threads, **kw)
        return [t for t, raw in ds.tensorized(model) if raw is None]

    cases = [dict(), dict(take_only_first_n_files=3), dict(limit_num_yielded_elements=10),
             dict(rank=1, world_size=2),             # 5 files >= 2 ranks: file-level sharding
             dict(rank=3, world_size=8),             # fewer files than ranks: element-level sharding
             dict(rank=0, world_size=8, limit_num_yielded_elements=2)]
    for kw in cases:
        expected = host(**kw)
        for threads in (1, 4):
            got = native(threads, **kw)
            assert len(got) == len(expected), kw
            for a, b in zip(expected, got):
                assert_same(a, b, str(kw))
    assert len(host()) == 35 and len(host(rank=1, world_size=2)) == 14 and len(host(limit_num_yielded_elements=10)) == 11

    # raw iteration (metadata pass) is the host loader itself
    assert len(list(shards.ShardDataset(rich))) == 35


def test_chunked_parallel_decode_keeps_file_order_and_cleans_up(model, tmp_path, monkeypatch, capsys):
    """Files are inflated whole and decoded chunk by chunk by a worker pool: several chunks per shard (CHUNK lowered to 3),
    nil objects, a truncated file and an unreadable one in the middle; every thread count yields the sequential order;
    a consumer that stops early leaves no worker running and no shard open."""
    import threading

    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.synthetic import SyntheticBugLabGenerator
    from dpu_utils.utils import RichPath

    monkeypatch.setattr(shards.NativeShardTensorizer, "CHUNK", 3)
    gen = SyntheticBugLabGenerator(seed=21, mean_nodes=90, min_nodes=30)
    directory = tmp_path / "d"
    directory.mkdir()
    for f in range(4):
        objects = [gen.sample() for _ in range(11 + f)]
        objects.insert(2, None)          # nil objects are skipped (msgpackutils.py:38) without shifting the chunk order
        objects.insert(7, None)
        write_objects(str(directory / f"shard{f}.msgpack.l.gz"), objects)
    (directory / "shard2.msgpack.l.gz").write_bytes(b"")          # unreadable / empty file in the middle
    rich = RichPath.create(str(directory))

    def native(threads, **kw):
        ds = shards.ShardDataset(rich, num_threads=threads, **kw)
        return [t for t, _ in ds.tensorized(model)]

    cases = (dict(), dict(rank=2, world_size=8), dict(limit_num_yielded_elements=17), dict(rank=1, world_size=2))
    for truncated in (False, True):
        if truncated:
            # a truncated stream: the objects before the break are served, then the error is reported (how many objects the
            # host-language reader salvages from a truncated stream depends on its read-ahead, so only the native orders
            # are compared with each other here)
            whole = (directory / "shard1.msgpack.l.gz").read_bytes()
            (directory / "shard1.msgpack.l.gz").write_bytes(whole[: len(whole) * 2 // 3])
            capsys.readouterr()
        for kw in cases:
            expected = native(1, **kw)
            assert len(expected) > 0, kw
            if not truncated:
                host = [t for t in (model.tensorize(dp) for dp in load_all_msgpack_l_gz(rich, **kw)) if t is not None]
                assert len(expected) == len(host), kw
                for a, b in zip(host, expected):
                    assert_same(a, b, f"sequential {kw}")
            for threads in (2, 3, 8):
                got = native(threads, **kw)
                assert len(got) == len(expected), (kw, threads)
                for a, b in zip(expected, got):
                    assert_same(a, b, f"{threads} threads {kw}")
    assert "shard1.msgpack.l.gz: corrupt or truncated gzip stream" in capsys.readouterr().out

    # early stop: close the generator after a few samples
    before = threading.active_count()
    tensorizer = shards.NativeShardTensorizer(model)
    opened = []
    real_open = shards.NativeShardTensorizer._open

    def tracking_open(path):
        shard = real_open(path)
        if shard is not None:
            opened.append(shard)
        return shard

    monkeypatch.setattr(shards.NativeShardTensorizer, "_open", staticmethod(tracking_open))
    paths = sorted(str(p) for p in directory.iterdir())
    it = tensorizer.tensorize_files(paths, num_threads=4)
    for _ in range(5):
        next(it)
    it.close()
    assert opened and all(s._h is None for s in opened)        # every shard that was opened has been closed
    assert threading.active_count() <= before                   # the pool's workers are gone
    assert tensorizer.num_native >= 5


def test_decode_many_equals_one_call_per_sample(model, tmp_path):
    """bl_sample_decode_many (one native call per chunk) fills the same views as bl_sample_decode per object."""
    import ctypes

    from buglab_b200.synthetic import write_shards

    path = write_shards(str(tmp_path / "d"), 1, 9, seed=4, mean_nodes=100, min_nodes=30)[0]
    tz = shards.NativeShardTensorizer(model)
    with shards.Shard(path) as shard:
        one_by_one = [tz.tensorize_object(shard, i) for i in range(len(shard))]
        chunked = tz._decode_chunk(shard, list(range(len(shard))))
        assert len(chunked) == len(one_by_one) == 9
        for a, b in zip(one_by_one, chunked):
            assert_same(a, b)
        # argument checking of the chunk entry point
        lib = shards.lib()
        assert lib.bl_sample_decode_many(shard._h, None, 0, tz._tokenizer._h, tz._edge_names, tz._num_edge_types, None, None) == 0
        assert lib.bl_sample_decode_many(shard._h, None, 2, tz._tokenizer._h, tz._edge_names, tz._num_edge_types, None, None) != 0
        bufs = shards._ChunkBuffers(2)
        bufs.indices[:2] = [0, len(shard)]                       # second index out of range: the call reports it
        assert lib.bl_sample_decode_many(shard._h, bufs.indices, 2, tz._tokenizer._h, tz._edge_names, tz._num_edge_types,
                                         bufs.handles, bufs.views) != 0


def test_predict_from_native_shards_equals_the_host_chain(tmp_path):
    """``model.predict`` over shard files through the reference-shaped loader, ShardDataset (unpacked datapoints) and
    ShardDataset (lazy datapoints: what ``python -m buglab.models.evaluate`` uses): identical predictions, datapoints and
    evaluation metrics, and the lazy datapoints are never unpacked by predict + evaluate.  Runs in a child process because the
    CPU oracle backend it computes with patches the package process-wide."""
    import json
    import subprocess

    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "predict_sources_child.py"), str(tmp_path)],
                          capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stderr[-3000:]
    line = json.loads([l for l in proc.stdout.splitlines() if l.startswith("{")][-1])
    assert line["samples"] == 120
    assert line["lazy_still_packed_after_predict"] == 120 and line["lazy_still_packed_after_metrics"] == 120


def test_lazy_datapoint_is_a_faithful_read_only_view(model, tmp_path):
    from buglab.utils.msgpackutils import load_msgpack_l_gz
    from buglab_b200.synthetic import write_shards

    path = write_shards(str(tmp_path / "d"), 1, 5, seed=8, mean_nodes=90, min_nodes=30)[0]
    host = [dp for dp in load_msgpack_l_gz(path)]
    for dp in host:
        model.tensorize(dp)          # the host chain extends the graph in place while tensorising (data.py:97-121)
    tz = shards.NativeShardTensorizer(model)
    pairs = list(tz.tensorize_files([path], 1, datapoints=tz.LAZY_DATAPOINTS))
    assert len(pairs) == len(host) == 5
    for (t, lazy), dp in zip(pairs, host):
        assert isinstance(lazy, shards.LazyDatapoint) and lazy._full is None
        assert lazy["target_fix_action_idx"] == dp["target_fix_action_idx"]
        assert lazy["candidate_rewrites"] == dp["candidate_rewrites"]
        assert lazy["candidate_rewrite_metadata"] == dp["candidate_rewrite_metadata"]
        graph = lazy["graph"]
        assert graph["reference_nodes"] == dp["graph"]["reference_nodes"] and type(graph["reference_nodes"]) is list
        assert lazy.get("target_fix_action_idx", "x") == dp["target_fix_action_idx"] and "graph" in lazy
        assert lazy._full is None and graph._full is None                    # nothing above needed the graph
        assert graph["nodes"] == dp["graph"]["nodes"]                        # ... this does: incl. the appended subtoken nodes
        assert lazy._full is not None
        assert lazy == dp and dict(lazy) == dp and list(lazy) == list(dp) and len(lazy) == len(dp)
        assert graph == dp["graph"] and lazy["graph"]["edges"]["HasSubtoken"] == dp["graph"]["edges"]["HasSubtoken"]
        assert lazy.get("no_such_key") is None and "no_such_key" not in lazy
        with pytest.raises(KeyError):
            lazy["no_such_key"]
        with pytest.raises(TypeError):
            lazy["x"] = 1                                                    # read-only
    full = list(tz.tensorize_files([path], 1, datapoints=tz.FULL_DATAPOINTS))
    for (t, dp_full), (t_lazy, _), dp in zip(full, pairs, host):
        assert type(dp_full) is dict and dp_full == dp
        assert_same(t, t_lazy)


def test_native_metadata_pass_equals_the_host_pass(tmp_path, base_sample):
    """``model.compute_metadata(ShardDataset(...))`` — subtoken counts and edge-type names counted by the native decoder's
    worker threads — builds the same vocabulary (ids included) and edge-type tuple as the pass over raw datapoints, for every
    thread count, element limit and rank sharding, with nil objects, samples the decoder declines (non-ASCII identifiers,
    stale HasSubtoken, no NextToken, ...), and token-level splitting."""
    from pathlib import Path

    from buglab.models.modelregistry import load_model
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.synthetic import SyntheticBugLabGenerator
    from dpu_utils.utils import RichPath

    gen = SyntheticBugLabGenerator(seed=31, mean_nodes=120, min_nodes=40)
    directory = tmp_path / "d"
    directory.mkdir()
    irregular = list(_variants(base_sample).values())
    for f in range(3):
        objects = [gen.sample() for _ in range(37 + f)]
        objects.insert(5, None)
        objects[10:10] = irregular[f::3]
        write_objects(str(directory / f"shard{f}.msgpack.l.gz"), objects)
    rich = RichPath.create(str(directory))

    def fresh(**spec):
        m, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 16, **spec}, Path(str(tmp_path / "m.pkl.gz")))
        return m

    def metadata_of(m):
        vocab = m.gnn_model.node_representation_model.vocabulary
        return dict(vocab.token_to_id), list(vocab.id_to_token), m.gnn_model.edge_types

    cases = [dict(), dict(limit_num_yielded_elements=40), dict(limit_num_yielded_elements=1), dict(rank=1, world_size=2),
             dict(rank=5, world_size=8, limit_num_yielded_elements=9), dict(take_only_first_n_files=2)]
    for spec in (dict(), dict(node_representations={"token_splitting": "token"})):
        for kw in cases:
            host = fresh(**spec)
            host.compute_metadata(load_all_msgpack_l_gz(rich, **kw))
            for threads in (1, 3):
                native = fresh(**spec)
                source = shards.ShardDataset(rich, num_threads=threads, **kw)
                assert source.update_model_metadata(fresh(**spec)) is True        # the pass is taken, not declined
                native.compute_metadata(source)
                assert metadata_of(native) == metadata_of(host), (spec, kw, threads)
    assert len(metadata_of(host)[0]) > 20

    # a malformed sample raises in the native pass exactly as in the host pass (it is handed to model.update_metadata_from)
    broken = copy.deepcopy(base_sample)
    del broken["graph"]["reference_nodes"]
    bad_dir = tmp_path / "bad"
    bad_dir.mkdir()
    write_objects(str(bad_dir / "shard0.msgpack.l.gz"), [gen.sample() for _ in range(40)] + [broken] + [gen.sample() for _ in range(40)])
    bad = RichPath.create(str(bad_dir))
    with pytest.raises(KeyError):
        fresh().compute_metadata(load_all_msgpack_l_gz(bad))
    for threads in (1, 3):
        with pytest.raises(KeyError):
            fresh().compute_metadata(shards.ShardDataset(bad, num_threads=threads))

    # models the native pass does not cover are told so and get the raw datapoints
    class Other:
        pass

    assert shards.ShardDataset(rich).update_model_metadata(Other()) is False


def test_metadata_accumulator_c_abi(model, tmp_path):
    import ctypes

    from buglab_b200.synthetic import write_shards

    path = write_shards(str(tmp_path / "d"), 1, 4, seed=6, mean_nodes=80, min_nodes=30)[0]
    lib = shards.lib()
    acc = ctypes.c_void_p()
    assert lib.bl_metadata_create(ctypes.byref(acc)) == 0
    splitter = shards.Tokenizer(None, "subtoken", 6)
    sample = shards._SampleBuffer()
    with shards.Shard(path) as shard:
        assert shard.non_nil_indices() == [0, 1, 2, 3]
        idx = (ctypes.c_int64 * 4)(0, 1, 2, 3)
        declined, n_declined = (ctypes.c_int32 * 4)(), ctypes.c_int32()
        assert lib.bl_metadata_add(acc, shard._h, idx, 4, splitter._h, sample.handle, declined, ctypes.byref(n_declined)) == 0
        assert n_declined.value == 0 and lib.bl_metadata_num_samples(acc) == 4
        bad = (ctypes.c_int64 * 1)(9)
        assert lib.bl_metadata_add(acc, shard._h, bad, 1, splitter._h, sample.handle, declined, ctypes.byref(n_declined)) != 0
        assert lib.bl_metadata_add(None, shard._h, idx, 4, splitter._h, sample.handle, declined, ctypes.byref(n_declined)) != 0
    nbytes = ctypes.c_int64()
    n_tokens = lib.bl_metadata_size(acc, 0, ctypes.byref(nbytes))
    n_types = lib.bl_metadata_size(acc, 1, None)
    assert n_tokens > 10 and nbytes.value > n_tokens and n_types >= 3 and lib.bl_metadata_size(acc, 2, None) == -1
    lib.bl_metadata_destroy(acc)


def test_trainer_consumes_self_tensorizing_datasets(model, tmp_path):
    """ModelTrainer.train asks a data source with ``tensorized`` for tensors; minibatches packed from them are the ones the
    host loader gives (host-side packing only — no device work)."""
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.synthetic import write_shards
    from dpu_utils.utils import RichPath

    write_shards(str(tmp_path / "d"), 2, 6, seed=4, mean_nodes=100, min_nodes=30)
    rich = RichPath.create(str(tmp_path / "d"))

    def pack(tensorized):
        mb = model.initialize_minibatch()
        for t in tensorized:
            model.extend_minibatch_with(t, mb)
        return mb

    host_mb = pack(model.tensorize(dp) for dp in load_all_msgpack_l_gz(rich))
    native_mb = pack(t for t, _ in shards.ShardDataset(rich, num_threads=2).tensorized(model))
    assert_same(host_mb, native_mb, "minibatch")


def test_native_dataset_through_the_trainers_producer_thread(model, tmp_path):
    """The exact chain ModelTrainer.train runs per epoch — ShardDataset.tensorized -> (streaming shuffle) ->
    minibatch_iterator -> _Prefetcher's producer thread — with the device work left out (device="cpu")."""
    import random

    import torch
    from buglab.utils.msgpackutils import load_all_msgpack_l_gz
    from buglab_b200.synthetic import write_shards
    from dpu_utils.utils import RichPath
    from ptgnn.baseneuralmodel.trainer import _buffered_shuffle, _Prefetcher

    write_shards(str(tmp_path / "d"), 3, 8, seed=11, mean_nodes=100, min_nodes=30)
    rich = RichPath.create(str(tmp_path / "d"))
    dataset = shards.ShardDataset(rich, shuffle=True)

    def make():
        tensors = _buffered_shuffle(dataset.tensorized(model), 8, random.Random(1))
        return model.minibatch_iterator(tensors, device="cpu", max_minibatch_size=5)

    for epoch in range(2):  # the tensoriser is created in the first epoch's thread and reused from another one
        batches = list(_Prefetcher(make, torch.device("cpu")))
        assert [len(raw) for _, raw in batches] == [5, 5, 5, 5, 4]
        assert sum(int(mb["graph_data"]["num_graphs"]) for mb, _ in batches) == 24
    total_nodes = sum(int(mb["graph_data"]["node_to_graph_idx"].shape[0]) for mb, _ in batches)
    expected_nodes = sum(model.tensorize(dp).graph_data.num_nodes for dp in load_all_msgpack_l_gz(rich))
    assert total_nodes == expected_nodes
    assert dataset.tensorizer.num_native == 48 and dataset.tensorizer.num_host == 0
