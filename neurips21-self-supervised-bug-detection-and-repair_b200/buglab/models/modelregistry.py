"""Model name -> constructor (reference: buglab/models/modelregistry.py:18-166).  Only the gnn-mlp family is built on
the B200 kernels; the other registry names are kept so callers get a clear error instead of a KeyError."""
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
    """Picklable ``n_edges -> layer list`` closure (the reference uses a lambda, modelregistry.py:85-87)."""

    def __init__(self, mp_layer, hidden_state_size, dropout_rate, edge_feature_size):
        self.mp_layer, self.hidden, self.dropout, self.features = mp_layer, hidden_state_size, dropout_rate, edge_feature_size

    def __call__(self, n_edges: int):
        return self.mp_layer(self.hidden, self.dropout, n_edges, features_dimension=self.features)


def gnn(*, mp_layer, add_self_edge: bool, use_all_gnn_layer_outputs: bool = False, hidden_state_size: int = 128,
        dropout_rate: float = 0.2, node_representations: Optional[Dict[str, Any]] = None,
        selector_loss_type="classify-max-loss", stop_extending_minibatch_after_num_nodes: int = 30000,
        max_nodes_per_graph: int = 35000, buggy_samples_weight_spec: Union[str, int, float] = 1.0,
        edge_feature_size: int = 0, **kwargs):
    node_representations = dict(node_representations or {})
    node_representations.setdefault("token_splitting", "subtoken")
    node_representations.setdefault("max_num_subtokens", 6)
    node_representations.setdefault("subtoken_combination", "max")
    node_representations.setdefault("vocabulary_size", 15000)
    if edge_feature_size > 0:
        raise NotImplementedError("edge_feature_size > 0 is outside the gnn-mlp default path built here")
    return GnnBugLabModel(
        GraphNeuralNetworkModel(
            node_representation_model=StrElementRepresentationModel(embedding_size=hidden_state_size, **node_representations),
            edge_representation_model=None,
            add_self_edges=add_self_edge,
            message_passing_layer_creator=_LayerStack(mp_layer, hidden_state_size, dropout_rate, edge_feature_size),
            stop_extending_minibatch_after_num_nodes=stop_extending_minibatch_after_num_nodes,
            max_nodes_per_graph=max_nodes_per_graph,
        ),
        use_all_gnn_layer_outputs=use_all_gnn_layer_outputs,
        generator_loss_type=selector_loss_type,
        buggy_samples_weight_schedule=buggy_sample_weight_schedule(buggy_samples_weight_spec),
    )


def seq_transformer(*, layer_type: str, hidden_state_size: int = 256, dropout_rate: float = 0.1, vocab_size: int = 15000,
                    selector_loss_type: str = "classify-max-loss", num_layers: int = 5, num_heads: int = 8,
                    max_seq_size: int = 400, intermediate_dimension_size: int = 1024,
                    buggy_samples_weight_spec: Union[str, int, float] = 1.0, rezero_mode: str = "off",
                    normalisation_mode: str = "postnorm", **__):
    """Reference modelregistry.py:97-126.  Host side only for now: the model tensorises and packs, its
    ``build_neural_module`` raises until the relational-transformer kernels exist."""
    from buglab.models.seqmodel import SeqBugLabModel

    return SeqBugLabModel(
        hidden_state_size, max_subtoken_vocab_size=vocab_size, dropout_rate=dropout_rate, layer_type=layer_type,
        generator_loss_type=selector_loss_type, intermediate_dimension_size=intermediate_dimension_size,
        buggy_samples_weight_schedule=buggy_sample_weight_schedule(buggy_samples_weight_spec), max_seq_size=max_seq_size,
        num_heads=num_heads, num_layers=num_layers, rezero_mode=rezero_mode, normalisation_mode=normalisation_mode)


def construct_model_dict(gnn_constructor: Callable, seq_constructor: Callable) -> Dict[str, Callable]:
    return {
        "gnn-mlp": lambda kwargs: gnn_constructor(mp_layer=create_mlp_mp_layers, add_self_edge=True, **kwargs),
        "ggnn": lambda kwargs: gnn_constructor(mp_layer=create_ggnn_mp_layers, add_self_edge=False, **kwargs),
        "seq-great": lambda kwargs: seq_constructor(layer_type="great", **kwargs),
        "seq-rat": lambda kwargs: seq_constructor(layer_type="rat", **kwargs),
        "seq-transformer": lambda kwargs: seq_constructor(layer_type="transformer", **kwargs),
        "seq-gru": lambda kwargs: seq_constructor(layer_type="gru", **kwargs),
    }


def load_model(model_spec: Dict[str, Any], model_path: Path, restore_path: Optional[str] = None,
               restore_if_model_exists: bool = False, type_model: bool = False
               ) -> Tuple[AbstractNeuralModel, Optional[ModuleWithMetrics], bool]:
    """(model, restored nn or None, whether metadata still has to be computed)."""
    assert model_path.name.endswith(".pkl.gz"), "MODEL_FILENAME must have a `.pkl.gz` suffix."
    if restore_path is not None or (restore_if_model_exists and model_path.exists()):
        import torch

        source = Path(restore_path) if restore_path is not None else model_path
        LOGGER.info("Resuming training from %s." % source)
        model, nn = AbstractNeuralModel.restore_model(source, torch.device("cuda:0" if torch.cuda.is_available() else "cpu"))
        return model, nn, False
    models = construct_model_dict(gnn, seq_transformer)
    if model_spec["modelName"] not in models:
        raise ValueError("Unknown model `%s`. Known models: %s" % (model_spec["modelName"], list(models.keys())))
    spec = {k: v for k, v in model_spec.items() if k != "modelName"}
    return models[model_spec["modelName"]](spec), None, True
