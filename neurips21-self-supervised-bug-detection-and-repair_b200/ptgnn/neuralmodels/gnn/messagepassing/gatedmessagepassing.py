from .abstractmessagepassing import AbstractMessagePassingLayer


class GatedMessagePassingLayer(AbstractMessagePassingLayer):
    """GGNN layer — importable because buglab/models/gnnlayerdefs.py:1 imports the name, but the ``ggnn`` model is
    outside the B200 hot path (BASELINE.json north_star names gnn-mlp only)."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        raise NotImplementedError("GatedMessagePassingLayer (ggnn) has no B200 kernel path; use gnn-mlp")
