"""On-disk sample schema and its conversion to ``GraphData`` (reference: buglab/representations/data.py:14-20,
97-167).  Index bookkeeping here must match the reference bit for bit (tests/golden)."""
import re
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np
from dpu_utils.codeutils import split_identifier_into_parts
from dpu_utils.mlutils import Vocabulary
from ptgnn.neuralmodels.gnn import GraphData
from typing_extensions import TypedDict


class BugLabGraph(TypedDict):
    nodes: List[str]
    edges: Dict[str, List[Union[Tuple[int, int], Tuple[int, int, str]]]]
    path: str
    text: str
    reference_nodes: List[int]
    code_range: Tuple[Tuple[int, int], Tuple[int, int]]


_IDENTIFIER_START = re.compile(r"[a-zA-Z_][a-zA-Z0-9_]*")


def add_open_vocab_nodes_and_edges(graph: BugLabGraph) -> None:
    """Appends one node per distinct subtoken of the identifier tokens (tokens = endpoints of ``NextToken``) and a
    ``HasSubtoken`` edge token -> subtoken-node.  MUTATES ``graph`` exactly like the reference (data.py:97-121):
    a second call appends a fresh set of subtoken nodes and replaces the ``HasSubtoken`` list.
    Tokens are visited in the iteration order of a Python ``set`` filled edge by edge, as upstream does, because
    that order decides the ids of the new nodes."""
    next_token = graph["edges"].get("NextToken")
    if next_token is None:
        return
    token_nodes = set()
    for edge in next_token:
        token_nodes.add(edge[0])
        token_nodes.add(edge[1])
    nodes = graph["nodes"]
    subtoken_node: Dict[str, int] = {}
    has_subtoken: List[Tuple[int, int]] = []
    for token_idx in token_nodes:
        label = nodes[token_idx]
        if _IDENTIFIER_START.match(label) is None:
            continue
        for part in split_identifier_into_parts(label):
            part_idx = subtoken_node.get(part)
            if part_idx is None:
                part_idx = len(nodes)
                nodes.append(part)
                subtoken_node[part] = part_idx
            has_subtoken.append((token_idx, part_idx))
    graph["edges"]["HasSubtoken"] = has_subtoken


def _edge_array(adj_list) -> np.ndarray:
    if len(adj_list) == 0:
        return np.zeros((0, 2), dtype=np.int32)
    return np.array([(e[0], e[1]) for e in adj_list], dtype=np.int32)


class BugLabData(TypedDict):
    graph: BugLabGraph
    candidate_rewrites: List[Tuple[str, Any]]
    candidate_rewrite_metadata: List[Tuple[str, Any]]
    candidate_rewrite_ranges: List[Tuple[Tuple[int, int], Tuple[int, int]]]
    target_fix_action_idx: Optional[int]
    package_name: str
    candidate_rewrite_logprobs: Optional[List[float]]

    @classmethod
    def as_graph_data(cls, data: "BugLabData") -> Tuple[GraphData, Optional[int]]:
        """(graph, index of the target location among the sorted unique candidate nodes) — reference data.py:139-167."""
        graph = data["graph"]
        candidate_nodes, inverse = np.unique(graph["reference_nodes"], return_inverse=True)
        target = data["target_fix_action_idx"]
        target_node_idx = None if target is None else inverse[target]
        add_open_vocab_nodes_and_edges(graph)
        # The reference also materialises per-edge feature strings (data.py:158-161); gnn-mlp runs with
        # edge_feature_size == 0 (modelregistry.py:56) where they are never read, so they are not built here.
        return (
            GraphData(
                node_information=graph["nodes"],
                edges={name: _edge_array(adj) for name, adj in graph["edges"].items()},
                reference_nodes={"candidate_nodes": candidate_nodes},
            ),
            target_node_idx,
        )
