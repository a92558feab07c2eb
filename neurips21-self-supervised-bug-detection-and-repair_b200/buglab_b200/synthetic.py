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


class SyntheticProgramGenerator:
    """Samples whose graphs have the STRUCTURE the sequence models rely on (reference buglab/models/seqmodel.py:442-624):
    a single ``NextToken`` chain over the leaf tokens in source order, a ``Child`` tree whose internal nodes are AST
    labels (``Assign`` owns an ``=`` token, ``BinaryOperation`` an operator token, ``ComparisonTarget`` a comparison
    token, ``Call`` its ``args``), symbol nodes outside the tree reached by ``OccurrenceOf`` edges, and data-/control-flow
    relations between tokens and statements.  Node ids follow a pre-order walk, so tokens and AST nodes interleave as in
    extracted graphs.  The same sample works for the graph models (it is schema-valid ``BugLabData``)."""

    _BIN_OPS = ("+", "-", "*", "/", "//", "%", "**", "|", "&")
    _CMP_OPS = ("<", "<=", "==", "!=", ">", ">=", "is", "in")

    def __init__(self, seed: int = 0, statements: int = 12, vocabulary: int = 400, max_depth: int = 3):
        self.rng = np.random.default_rng(seed)
        self.statements, self.max_depth = statements, max_depth
        self.identifiers = _identifier_vocabulary(self.rng, vocabulary)
        self._sample_idx = 0

    def sample(self, statements: Optional[int] = None) -> Dict[str, Any]:
        rng = self.rng
        nodes: List[str] = []
        child: List[Any] = []
        tokens: List[int] = []
        sibling: List[Any] = []
        occurrences: Dict[str, List[int]] = {}
        calls: Dict[int, List[int]] = {}
        binops: List[int] = []
        statements_nodes: List[int] = []
        local_names = [self.identifiers[int(i)] for i in rng.integers(0, len(self.identifiers), size=int(rng.integers(4, 10)))]

        def new(label: str) -> int:
            nodes.append(label)
            return len(nodes) - 1

        def token(label: str, parent: int, field: str) -> int:
            t = new(label)
            tokens.append(t)
            child.append((parent, t, field))
            return t

        def name(parent: int, field: str) -> int:
            n = new("Name")
            child.append((parent, n, field))
            ident = local_names[int(rng.integers(0, len(local_names)))]
            occurrences.setdefault(ident, []).append(token(ident, n, "value"))
            return n

        def expression(parent: int, field: str, depth: int) -> int:
            r = rng.random()
            if depth >= self.max_depth or r < 0.4:
                if r < 0.1:
                    lit = new("Integer")
                    child.append((parent, lit, field))
                    token(str(int(rng.integers(0, 100))), lit, "value")
                    return lit
                return name(parent, field)
            if r < 0.7:
                b = new("BinaryOperation")
                child.append((parent, b, field))
                binops.append(b)
                expression(b, "left", depth + 1)
                token(self._BIN_OPS[int(rng.integers(0, len(self._BIN_OPS)))], b, "operator")
                expression(b, "right", depth + 1)
                return b
            if r < 0.85:
                c = new("Comparison")
                child.append((parent, c, field))
                expression(c, "left", depth + 1)
                target = new("ComparisonTarget")
                child.append((c, target, "comparisons"))
                token(self._CMP_OPS[int(rng.integers(0, len(self._CMP_OPS)))], target, "operator")
                expression(target, "comparator", depth + 1)
                return c
            call = new("Call")
            child.append((parent, call, field))
            name(call, "func")
            token("(", call, "lpar")
            args = []
            for i in range(int(rng.integers(2, 5))):
                if i:
                    token(",", call, "comma")
                before = len(nodes)
                expression(call, "args", depth + 1)
                args.append(before)          # the argument's root node was created first (pre-order ids)
            token(")", call, "rpar")
            calls[call] = args
            return call

        module = new("Module")
        for _ in range(self.statements if statements is None else statements):
            r = rng.random()
            if r < 0.5:
                st = new("Assign")
                child.append((module, st, "body"))
                name(st, "targets")
                token("=", st, "equal")
                expression(st, "value", 0)
            elif r < 0.65:
                st = new("AugAssign")
                child.append((module, st, "body"))
                name(st, "target")
                token("+=", st, "operator")
                expression(st, "value", 1)
            elif r < 0.8:
                st = new("Return")
                child.append((module, st, "body"))
                token("return", st, "keyword")
                expression(st, "value", 0)
            elif r < 0.9:
                st = new("If")
                child.append((module, st, "body"))
                token("if", st, "keyword")
                expression(st, "test", 1)
                token(":", st, "colon")
            else:
                st = new("Expr")
                child.append((module, st, "body"))
                expression(st, "value", 1)
            statements_nodes.append(st)
        for parent in set(p for p, _, _ in child):
            kids = [c for p, c, _ in child if p == parent]
            sibling.extend((a, b) for a, b in zip(kids, kids[1:]))

        symbol_of: Dict[str, int] = {ident: new(ident) for ident in occurrences}      # symbol nodes live outside the tree
        occurrence_edges = [(t, symbol_of[ident]) for ident, toks in occurrences.items() for t in toks]
        ident_tokens = [t for toks in occurrences.values() for t in toks]

        def token_pairs(count: int):
            if len(ident_tokens) < 2:
                return []
            picks = rng.integers(0, len(ident_tokens), size=(count, 2))
            return [(int(ident_tokens[a]), int(ident_tokens[b])) for a, b in picks]

        edges: Dict[str, List] = {
            "Child": child,
            "NextToken": [(a, b) for a, b in zip(tokens, tokens[1:])],
            "Sibling": sibling,
            "OccurrenceOf": occurrence_edges,
            "NextMayUse": token_pairs(len(ident_tokens)),
            "LastMayWrite": token_pairs(len(ident_tokens) // 2),
            "ComputedFrom": token_pairs(len(ident_tokens) // 2),
            "ControlFlowNext": [(a, b) for a, b in zip(statements_nodes, statements_nodes[1:])],
        }

        reference_nodes: List[int] = []
        rewrites: List[Any] = []
        metadata: List[Any] = []
        symbols = list(symbol_of.values())
        for b in binops:                                   # operator rewrites at every binary operation
            for op in rng.choice(len(_TEXT_OPS), size=int(rng.integers(2, 5)), replace=False):
                reference_nodes.append(b)
                rewrites.append(("ReplaceText", _TEXT_OPS[int(op)]))
                metadata.append(("BinaryOperatorRewriteScout", None))
        for t in rng.permutation(ident_tokens)[: max(2, len(ident_tokens) // 3)].tolist():   # variable misuse at name tokens
            for sym in rng.permutation(symbols)[: int(rng.integers(2, 5))].tolist():
                reference_nodes.append(int(t))
                rewrites.append(("ReplaceText", nodes[sym]))
                metadata.append(("VariableMisuseRewriteScout", int(sym)))
        for call, args in calls.items():                  # argument swaps at calls
            for _ in range(int(rng.integers(1, 3))):
                a, b = rng.choice(len(args), size=2, replace=False)
                reference_nodes.append(call)
                rewrites.append(("ArgSwap", (int(a), int(b))))
                metadata.append(("ArgSwapRewriteScout", None))
        if not reference_nodes:                            # degenerate tiny program: one harmless candidate
            reference_nodes.append(tokens[0])
            rewrites.append(("ReplaceText", "0"))
            metadata.append(("LiteralRewriteScout", None))
        target = None if rng.random() < 0.4 else int(rng.integers(0, len(reference_nodes)))
        self._sample_idx += 1
        return {
            "graph": {"nodes": nodes, "edges": edges, "path": f"pkg/prog_{self._sample_idx}.py", "text": "",
                      "reference_nodes": reference_nodes, "code_range": ((0, 0), (1, 0))},
            "candidate_rewrites": rewrites,
            "candidate_rewrite_metadata": metadata,
            "candidate_rewrite_ranges": [((0, 0), (0, 1))] * len(rewrites),
            "target_fix_action_idx": target,
            "package_name": "synthetic",
        }

    def samples(self, count: int) -> Iterator[Dict[str, Any]]:
        for _ in range(count):
            yield self.sample()


def write_shards(directory: str, num_shards: int, graphs_per_shard: int, seed: int = 0, programs: bool = False,
                 **generator_kwargs) -> List[str]:
    """Writes ``num_shards`` files ``shard_XXXX.msgpack.l.gz`` in the reference wire format (``programs``: token-chain
    program graphs from :class:`SyntheticProgramGenerator`, which the sequence models can project onto token sequences)."""
    import os

    from buglab.utils.msgpackutils import save_msgpack_l_gz

    os.makedirs(directory, exist_ok=True)
    gen = (SyntheticProgramGenerator if programs else SyntheticBugLabGenerator)(seed=seed, **generator_kwargs)
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
