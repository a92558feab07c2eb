"""Seeded synthetic ``BugLabData`` samples and ``.msgpack.l.gz`` shards.

No real RandomBugs / PyPIBugs data or extractor is available offline, so every benchmark and
test is driven by schema-valid synthetic samples (schema: reference buglab/representations/data.py:14-20,130-137;
producer buglab/data/... buggydatacreation).  Shape knobs follow the survey: node count ~ LogNormal around
``mean_nodes`` clipped to [50, 35000]; ~4.5 forward edges per node over 7 named relation kinds (+ the
``HasSubtoken`` kind that ``as_graph_data`` derives = 8 forward kinds, 17 as seen by a layer with backward and
self edges); 80 % of edges stay within +-32 node ids, 20 % are uniform; ``NextToken`` is one chain; ``Child`` is a
tree of ``(parent, child, field)`` triples with ``Call`` nodes owning ``args`` children; 20-60 rewrite candidates
over three scout families; half of the samples have no bug.
"""
from typing import Any, Dict, Iterator, List, Optional, Sequence

import numpy as np

FORWARD_EDGE_KINDS = ("Child", "NextToken", "Sibling", "ControlFlowNext", "OccurrenceOf", "NextMayUse", "LastMayWrite",
                      "ComputedFrom", "AssignedFrom", "ReturnsFrom", "YieldsFrom", "MayFinalUseOf", "CandidateCall",
                      "MayFormalName")
_AST_LABELS = ("Name", "Attribute", "Assign", "BinaryOperation", "Comparison", "If", "For", "Return", "Arg", "Integer",
               "SimpleString", "Subscript", "FunctionDef", "IndentedBlock", "Expr", "BooleanOperation")
_PUNCT = ("(", ")", "=", ",", ".", ":", "[", "]", "+", "-", "==", "<", "return", "if", "for", "in")
_FIELDS = ("value", "body", "target", "func", "left", "right", "test", "params")
_TEXT_OPS = ("+", "-", "*", "/", "==", "!=", "<", "<=", ">", ">=", "and", "or", " is ", " is not ", " in ", " not in ",
             "+=", "-=", "0", "1", "True", "False")
_STEMS = ("get", "set", "value", "name", "index", "count", "data", "item", "list", "node", "path", "file", "result",
          "config", "user", "key", "size", "type", "error", "info", "line", "text", "args", "self", "state", "id",
          "token", "graph", "edge", "model", "batch", "loss", "step", "time", "max", "min", "sum", "len", "str", "obj")


def _identifier_vocabulary(rng: np.random.Generator, size: int) -> List[str]:
    """``size`` identifiers of 1-4 stems, mixed snake_case / camelCase so subtoken splitting yields 1-6 parts."""
    out = []
    for i in range(size):
        k = int(rng.integers(1, 5))
        parts = [_STEMS[int(j)] for j in rng.integers(0, len(_STEMS), size=k)]
        if i % 7 == 0:
            parts.append(str(int(rng.integers(0, 100))))
        if rng.random() < 0.5:
            out.append("_".join(parts))
        else:
            out.append(parts[0] + "".join(p.capitalize() for p in parts[1:]))
    return out


class SyntheticBugLabGenerator:
    def __init__(self, seed: int = 0, mean_nodes: int = 2000, num_named_edge_kinds: int = 7, vocabulary: int = 20000,
                 min_nodes: int = 50, max_nodes: int = 35000, sigma: float = 0.6, forward_edges_per_node: float = 4.5):
        assert 2 <= num_named_edge_kinds <= len(FORWARD_EDGE_KINDS)
        self.rng = np.random.default_rng(seed)
        self.mean_nodes, self.min_nodes, self.max_nodes, self.sigma = mean_nodes, min_nodes, max_nodes, sigma
        self.edge_kinds = FORWARD_EDGE_KINDS[:num_named_edge_kinds]
        self.edges_per_node = forward_edges_per_node
        self.identifiers = _identifier_vocabulary(self.rng, vocabulary)
        ranks = np.arange(1, vocabulary + 1, dtype=np.float64)
        self.zipf_p = ranks ** -1.2
        self.zipf_p /= self.zipf_p.sum()
        self._sample_idx = 0

    # ------------------------------------------------------------------------------------------
    def _num_nodes(self) -> int:
        mu = np.log(self.mean_nodes) - 0.5 * self.sigma ** 2
        return int(np.clip(self.rng.lognormal(mu, self.sigma), self.min_nodes, self.max_nodes))

    def _local_edges(self, n: int, count: int) -> np.ndarray:
        rng = self.rng
        src = rng.integers(0, n, size=count)
        local = np.clip(src + rng.integers(-32, 33, size=count), 0, n - 1)
        tgt = np.where(rng.random(count) < 0.8, local, rng.integers(0, n, size=count))
        return np.stack((src, tgt), axis=1)

    def sample(self, num_nodes: Optional[int] = None) -> Dict[str, Any]:
        rng = self.rng
        n = self._num_nodes() if num_nodes is None else int(num_nodes)
        n_tokens = max(8, n // 2)  # token nodes come first and form the NextToken chain
        labels: List[str] = []
        ident_ids = rng.choice(len(self.identifiers), size=n, p=self.zipf_p)
        kind = rng.random(n)
        for i in range(n):
            if i < n_tokens:
                labels.append(self.identifiers[ident_ids[i]] if kind[i] < 0.6 else _PUNCT[int(kind[i] * 997) % len(_PUNCT)])
            else:
                labels.append(_AST_LABELS[int(kind[i] * 991) % len(_AST_LABELS)])

        edges: Dict[str, List] = {}
        # Child: a tree; every node i>0 hangs off an earlier node within a window of 32
        parents = np.maximum(0, np.arange(1, n) - rng.integers(1, 33, size=n - 1))
        fields = rng.integers(0, len(_FIELDS), size=n - 1)
        child = [(int(p), int(c), _FIELDS[int(f)]) for p, c, f in zip(parents, np.arange(1, n), fields)]
        # a few Call nodes (AST region) with 2-4 positional args each
        num_calls = max(1, n // 200)
        call_nodes = rng.choice(np.arange(n_tokens, n), size=min(num_calls, n - n_tokens), replace=False)
        call_args: Dict[int, List[int]] = {}
        for c in call_nodes.tolist():
            labels[c] = "Call"
            args = rng.integers(0, n, size=int(rng.integers(2, 5))).tolist()
            call_args[c] = args
            child.extend((c, a, "args") for a in args)
        edges["Child"] = child
        edges["NextToken"] = [(i, i + 1) for i in range(n_tokens - 1)]
        others = [k for k in self.edge_kinds if k not in ("Child", "NextToken")]
        remaining = max(0, int(self.edges_per_node * n) - len(child) - (n_tokens - 1))
        # long-tailed split of the remaining edges over the other kinds
        w = np.array([0.5 ** i for i in range(len(others))]) if others else np.zeros(0)
        for k, share in zip(others, (w / w.sum() if others else w)):
            arr = self._local_edges(n, int(remaining * share))
            if k == "OccurrenceOf" and arr.shape[0] > 8:
                hubs = rng.integers(n_tokens, n, size=max(1, n // 100))  # symbol nodes with many occurrences
                arr[:, 1] = hubs[rng.integers(0, hubs.shape[0], size=arr.shape[0])]
            edges[k] = [(int(a), int(b)) for a, b in arr]

        # rewrite candidates: locations x family
        reference_nodes: List[int] = []
        rewrites: List[Any] = []
        metadata: List[Any] = []
        target_total = int(rng.integers(20, 61))
        call_list = list(call_args)
        while len(reference_nodes) < target_total:
            family = rng.random()
            if family < 0.15 and call_list:
                loc = int(call_list[int(rng.integers(0, len(call_list)))])
                nargs = len(call_args[loc])
                for _ in range(int(rng.integers(1, 4))):
                    a, b = rng.choice(nargs, size=2, replace=False)
                    reference_nodes.append(loc)
                    rewrites.append(("ArgSwap", (int(a), int(b))))
                    metadata.append(("ArgSwapRewriteScout", None))
            elif family < 0.55:
                loc = int(rng.integers(0, n_tokens))
                for _ in range(int(rng.integers(2, 7))):
                    reference_nodes.append(loc)
                    rewrites.append(("ReplaceText", self.identifiers[int(rng.integers(0, 50))]))
                    metadata.append(("VariableMisuseRewriteScout", int(rng.integers(0, n))))
            else:
                loc = int(rng.integers(0, n))
                for op in rng.choice(len(_TEXT_OPS), size=int(rng.integers(2, 6)), replace=False):
                    reference_nodes.append(loc)
                    rewrites.append(("ReplaceText", _TEXT_OPS[int(op)]))
                    metadata.append(("BinaryOperatorRewriteScout", None))
        target = None if rng.random() < 0.5 else int(rng.integers(0, len(reference_nodes)))
        self._sample_idx += 1
        return {
            "graph": {
                "nodes": labels, "edges": edges, "path": f"pkg/mod_{self._sample_idx}.py", "text": "",
                "reference_nodes": reference_nodes, "code_range": ((0, 0), (1, 0)),
            },
            "candidate_rewrites": rewrites,
            "candidate_rewrite_metadata": metadata,
            "candidate_rewrite_ranges": [((0, 0), (0, 1))] * len(rewrites),
            "target_fix_action_idx": target,
            "package_name": "synthetic",
        }

    def samples(self, count: int) -> Iterator[Dict[str, Any]]:
        for _ in range(count):
            yield self.sample()


def write_shards(directory: str, num_shards: int, graphs_per_shard: int, seed: int = 0, **generator_kwargs) -> List[str]:
    """Writes ``num_shards`` files ``shard_XXXX.msgpack.l.gz`` in the reference wire format."""
    import os

    from buglab.utils.msgpackutils import save_msgpack_l_gz

    os.makedirs(directory, exist_ok=True)
    gen = SyntheticBugLabGenerator(seed=seed, **generator_kwargs)
    paths = []
    for s in range(num_shards):
        path = os.path.join(directory, f"shard_{s:04d}.msgpack.l.gz")
        save_msgpack_l_gz(gen.samples(graphs_per_shard), path)
        paths.append(path)
    return paths


def packed_edge_batch(num_nodes: int, num_edges: int, num_edge_types: int, seed: int = 0):
    """BASELINE config 3: a directly synthesised packed batch (bypasses host packing) — int32 ``src, tgt, etype`` in
    type-major order with the locality mix and a long-tailed type distribution (3 kinds hold ~70 % of edges)."""
    rng = np.random.default_rng(seed)
    w = np.array([0.3, 0.25, 0.15] + [0.3 / max(1, num_edge_types - 3)] * max(0, num_edge_types - 3))[:num_edge_types]
    counts = np.floor(w / w.sum() * num_edges).astype(np.int64)
    counts[0] += num_edges - counts.sum()
    src = rng.integers(0, num_nodes, size=num_edges, dtype=np.int64)
    local = np.clip(src + rng.integers(-32, 33, size=num_edges), 0, num_nodes - 1)
    tgt = np.where(rng.random(num_edges) < 0.8, local, rng.integers(0, num_nodes, size=num_edges))
    etype = np.repeat(np.arange(num_edge_types), counts)
    return src.astype(np.int32), tgt.astype(np.int32), etype.astype(np.int32)
