"""Routes this repo's host-side mirror (``buglab.models.*`` over ``ptgnn`` / ``torch_scatter``) to the CPU oracle — TEST
INFRASTRUCTURE ONLY (imported by ``bench.py``'s CPU legs and by tests; never by the product, which has no CPU path).

The reference is pure Python whose arithmetic lives in ptgnn / torch_scatter / PyTorch CPU; neither
package is installable offline and /root/reference does not exist on the GPU box.  After :func:`install`, the kept entry
point ``buglab.models.train`` (reference buglab/models/train.py:54-138: ModelTrainer, 30 000-node minibatch budget of
modelregistry.py:53-54, Adam + clip 0.5 + warm-up) runs unchanged on the host cores with

  * ``MlpMessagePassingLayer`` / ``SubtokenUnitEmbedder``  -> oracle/mp_ref.py (per-edge Linear, GELU, scatter_max, ...)
  * ``torch_scatter.*`` and ``scatter_log_softmax``        -> oracle/scatter_ref.py
  * ``LayerNorm`` (only used inside the swapped layer), the flat fused optimiser -> torch.nn / torch.optim.Adam
  * no device plan.

This is exactly the substitution tests/golden/make_golden.py applies underneath the REAL reference modules to generate
the fixtures, so the CPU arm times the arithmetic the goldens pin.  The patch is process-wide: call it in a process that
does nothing else."""
import torch


def install() -> None:
    import buglab.models.gnnlayerdefs as gnnlayerdefs
    import buglab.models.train as train_mod
    import buglab.models.utils as utils_mod
    import ptgnn.neuralmodels.embeddings.strelementrepresentationmodel as srm
    import ptgnn.neuralmodels.gnn.graphneuralnetwork as gnn_mod
    import ptgnn.neuralmodels.gnn.messagepassing as mp_pkg
    import ptgnn.neuralmodels.gnn.messagepassing.mlpmessagepassing as mlp_mod
    from buglab_b200 import ops

    from . import mp_ref, scatter_ref

    for module in (mp_pkg, mlp_mod, gnnlayerdefs):
        module.MlpMessagePassingLayer = mp_ref.MlpMessagePassingLayer
    srm.SubtokenUnitEmbedder = mp_ref.SubtokenUnitEmbedder
    gnn_mod.plan_for = lambda adjacency_lists, num_nodes, block_nodes=None: None

    ops.layer_norm = lambda x, g, b, eps=1e-5: torch.nn.functional.layer_norm(x, (x.shape[-1],), g, b, eps)
    ops.segment_log_softmax = lambda src, index, eps=1e-12, num_segments=None: scatter_ref.scatter_log_softmax(src, index.long())
    ops.segment_minmax = lambda src, index, dim=-1, dim_size=None, is_min=False: (
        scatter_ref.scatter_min if is_min else scatter_ref.scatter_max)(src, index.long(), dim, dim_size)
    ops.segment_sum = lambda src, index, dim=-1, dim_size=None: scatter_ref.scatter_sum(src, index.long(), dim, dim_size)

    def adam(params, lr: float = 0.0001):
        return torch.optim.Adam(params, lr=lr)

    utils_mod.optimizer = adam
    train_mod.optimizer = adam
