"""Layer stacks of the GNN models (reference: buglab/models/gnnlayerdefs.py:5-68).

gnn-mlp = two blocks of [remember state, 3 x MP(H->H), concat-residual (-> 2H), MP(2H -> H with 2H-wide messages)];
all eight message-passing layers aggregate with ``max`` and use the same dropout rate."""
from ptgnn.neuralmodels.gnn.messagepassing import GatedMessagePassingLayer, MlpMessagePassingLayer
from ptgnn.neuralmodels.gnn.messagepassing.residuallayers import ConcatResidualLayer


def create_mlp_mp_layers(hidden_state_size, dropout_rate, num_edges: int, features_dimension: int = 0):
    def mp(width_factor: int):
        return MlpMessagePassingLayer(
            input_state_dimension=width_factor * hidden_state_size,
            message_dimension=width_factor * hidden_state_size,
            output_state_dimension=hidden_state_size,
            num_edge_types=num_edges,
            message_aggregation_function="max",
            dropout_rate=dropout_rate,
            features_dimension=features_dimension,
        )

    layers = []
    for _ in range(2):
        residual = ConcatResidualLayer(hidden_state_size)
        layers += [residual.pass_through_dummy_layer(), mp(1), mp(1), mp(1), residual, mp(2)]
    return layers


def create_ggnn_mp_layers(hidden_state_size, dropout_rate, num_edges: int):
    raise NotImplementedError("the ggnn model is outside the B200 hot path (BASELINE.json north_star: gnn-mlp); "
                              f"{GatedMessagePassingLayer.__name__} has no kernel path")
