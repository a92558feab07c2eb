"""Model name -> constructor (reference: buglab/models/modelregistry.py:18-166): the graph family (``gnn-mlp`` on the B200
kernels; ``ggnn`` is named but its gated layer is not built) and the sequence family (``seq-*``)."""
import logging
import re
from functools import partial
from pathlib import Path
from typing import Any, Callable, Dict, Optional, Tuple, Union

from ptgnn.baseneuralmodel import AbstractNeuralModel, ModuleWithMetrics
from ptgnn.neuralmodels.embeddings.strelementrepresentationmodel import StrElementRepresentationModel
from ptgnn.neuralmodels.gnn import GraphNeuralNetworkModel

from buglab.models.gnn import GnnBugLabModel
from buglab.models.gnnlayerdefs import create_ggnn_mp_layers, create_mlp_mp_layers

LOGGER = logging.getLogger(__name__)

_WARMDOWN = re.compile(r"warmdown\(([0-9]+),\s?([0-9]*\.[0-9]+)\)")


def const_schedule(epoch_idx: int, const_weight: float) -> float:
    return const_weight


def linear_warmdown(epoch_idx: int, num_warmdown_epochs: int, target_weight: float) -> float:
    """1 -> target_weight linearly over ``num_warmdown_epochs`` epochs, then flat."""
    return max(target_weight, 1 + epoch_idx * (target_weight - 1) / num_warmdown_epochs)


def buggy_sample_weight_schedule(weight_spec: Union[str, int, float]) -> Callable[[int], float]:
    """A picklable epoch -> weight function from a number or a ``warmdown(n, w)`` spec (reference :22-41)."""
    if isinstance(weight_spec, (int, float)):
        return partial(const_schedule, const_weight=weight_spec)
    m = _WARMDOWN.match(weight_spec)
    if m:
        return partial(linear_warmdown, num_warmdown_epochs=int(m.group(1)), target_weight=float(m.group(2)))
    raise Exception(f"Unrecognized buggy sample weighting `{weight_spec}`")


class _LayerStack:
    """Picklable ``number of edge kinds -> layer list`` (upstream uses a lambda here)."""

    def __init__(self, mp_layer, hidden_state_size, dropout_rate, edge_feature_size):
        self.mp_layer, self.hidden, self.dropout, self.features = mp_layer, hidden_state_size, dropout_rate, edge_feature_size

    def __call__(self, n_edges: int):
        return self.mp_layer(self.hidden, self.dropout, n_edges, features_dimension=self.features)


# defaults of the node-label embedder when the spec does not override them (subtoken vocabulary, <= 6 parts, max-pooled)
_NODE_REPRESENTATION_DEFAULTS = {"token_splitting": "subtoken", "max_num_subtokens": 6, "subtoken_combination": "max",
                                 "vocabulary_size": 15000}


def gnn(*, mp_layer, add_self_edge: bool, use_all_gnn_layer_outputs: bool = False, hidden_state_size: int = 128,
        dropout_rate: float = 0.2, node_representations: Optional[Dict[str, Any]] = None,
        selector_loss_type="classify-max-loss", stop_extending_minibatch_after_num_nodes: int = 30000,
        max_nodes_per_graph: int = 35000, buggy_samples_weight_spec: Union[str, int, float] = 1.0,
        edge_feature_size: int = 0, **kwargs):
    """The graph family: a subtoken node embedder feeding ``mp_layer``'s message-passing stack (keyword surface of the
    reference's ``gnn``, modelregistry.py:44-94, because model specs are JSON dicts splatted into it)."""
    if edge_feature_size > 0:
        raise NotImplementedError("edge_feature_size > 0 is outside the gnn-mlp default path built here")
    embedder_spec = {**_NODE_REPRESENTATION_DEFAULTS, **(node_representations or {})}
    graph_model = GraphNeuralNetworkModel(
        node_representation_model=StrElementRepresentationModel(embedding_size=hidden_state_size, **embedder_spec),
        edge_representation_model=None,
        message_passing_layer_creator=_LayerStack(mp_layer, hidden_state_size, dropout_rate, edge_feature_size),
        add_self_edges=add_self_edge,
        max_nodes_per_graph=max_nodes_per_graph,
        stop_extending_minibatch_after_num_nodes=stop_extending_minibatch_after_num_nodes)
    return GnnBugLabModel(graph_model, use_all_gnn_layer_outputs=use_all_gnn_layer_outputs,
                          generator_loss_type=selector_loss_type,
                          buggy_samples_weight_schedule=buggy_sample_weight_schedule(buggy_samples_weight_spec))


def seq_transformer(*, layer_type: str, hidden_state_size: int = 256, dropout_rate: float = 0.1, vocab_size: int = 15000,
                    selector_loss_type: str = "classify-max-loss", num_layers: int = 5, num_heads: int = 8,
                    max_seq_size: int = 400, intermediate_dimension_size: int = 1024,
                    buggy_samples_weight_spec: Union[str, int, float] = 1.0, rezero_mode: str = "off",
                    normalisation_mode: str = "postnorm", **__):
    """The sequence family (keyword surface of the reference's ``seq_transformer``, modelregistry.py:97-126)."""
    from buglab.models.seqmodel import SeqBugLabModel

    encoder = dict(layer_type=layer_type, num_layers=num_layers, num_heads=num_heads, max_seq_size=max_seq_size,
                   intermediate_dimension_size=intermediate_dimension_size, rezero_mode=rezero_mode,
                   normalisation_mode=normalisation_mode)
    return SeqBugLabModel(hidden_state_size, max_subtoken_vocab_size=vocab_size, dropout_rate=dropout_rate,
                          generator_loss_type=selector_loss_type,
                          buggy_samples_weight_schedule=buggy_sample_weight_schedule(buggy_samples_weight_spec), **encoder)


# model name -> (family, keyword arguments the name fixes)
_MODEL_TABLE = {
    "gnn-mlp": ("graph", dict(mp_layer=create_mlp_mp_layers, add_self_edge=True)),
    "ggnn": ("graph", dict(mp_layer=create_ggnn_mp_layers, add_self_edge=False)),
    "seq-great": ("sequence", dict(layer_type="great")),
    "seq-rat": ("sequence", dict(layer_type="rat")),
    "seq-transformer": ("sequence", dict(layer_type="transformer")),
    "seq-gru": ("sequence", dict(layer_type="gru")),
}


class _ModelEntry:
    """``spec dict -> model`` for one registry name."""

    def __init__(self, constructor: Callable, fixed: Dict[str, Any]):
        self._constructor, self._fixed = constructor, fixed

    def __call__(self, spec: Dict[str, Any]):
        return self._constructor(**self._fixed, **spec)


def construct_model_dict(gnn_constructor: Callable, seq_constructor: Callable) -> Dict[str, Callable]:
    """Name -> ``callable(spec)``; the two constructors are parameters because the controllers wrap them upstream."""
    family = {"graph": gnn_constructor, "sequence": seq_constructor}
    return {name: _ModelEntry(family[kind], fixed) for name, (kind, fixed) in _MODEL_TABLE.items()}


def load_model(model_spec: Dict[str, Any], model_path: Path, restore_path: Optional[str] = None,
               restore_if_model_exists: bool = False, type_model: bool = False
               ) -> Tuple[AbstractNeuralModel, Optional[ModuleWithMetrics], bool]:
    """``(model, restored module or None, metadata still to be computed?)`` — a checkpoint (explicit ``restore_path``, or
    ``model_path`` itself when it exists and ``restore_if_model_exists``) wins over building from the spec."""
    assert model_path.name.endswith(".pkl.gz"), "MODEL_FILENAME must have a `.pkl.gz` suffix."
    checkpoint = Path(restore_path) if restore_path is not None else (
        model_path if restore_if_model_exists and model_path.exists() else None)
    if checkpoint is not None:
        import torch

        LOGGER.info("Resuming training from %s." % checkpoint)
        # restored on the host: ModelTrainer.train moves the module to this rank's device (restoring on cuda:0 would make
        # every torchrun rank materialise the model — and a CUDA context — on GPU 0 first)
        return (*AbstractNeuralModel.restore_model(checkpoint, torch.device("cpu")), False)
    registry = construct_model_dict(gnn, seq_transformer)
    name = model_spec["modelName"]
    if name not in registry:
        raise ValueError("Unknown model `%s`. Known models: %s" % (name, list(registry.keys())))
    return registry[name]({k: v for k, v in model_spec.items() if k != "modelName"}), None, True
