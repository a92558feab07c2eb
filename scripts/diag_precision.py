"""Diagnostic (GPU): where the forward rounding noise of the gnn-mlp stack comes from.  One c2-shaped minibatch (hidden 256,
8 graphs of ~2 000 nodes), exact (fp64) and fp32 CPU oracle node states computed once, then the B200 forward under several
arithmetic configurations; prints, per configuration and per layer-list entry, max |gpu - fp64| next to the oracle's own
max |fp32 - fp64|.  One JSON line per configuration at the end."""
import copy, json, os, sys
from pathlib import Path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch
from buglab.models.modelregistry import load_model
from buglab_b200 import ops
from buglab_b200.synthetic import SyntheticBugLabGenerator
from oracle import model_ref

hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n_graphs = int(sys.argv[2]) if len(sys.argv) > 2 else 8
mean_nodes = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
device = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.manual_seed(3)
gen = SyntheticBugLabGenerator(seed=3, mean_nodes=mean_nodes, min_nodes=40)
data = [gen.sample() for _ in range(n_graphs)]
model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": 0.0}, Path("/tmp/diag_precision.pkl.gz"))
model.gnn_model.node_representation_model.dropout_rate = 0.0
model.compute_metadata(iter(copy.deepcopy(data)))
nn = model.build_neural_module().to(device).eval()
ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types, len(model.gnn_model.node_representation_model.vocabulary),
                                len(model._target_rewrite_ops))
ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
ref64 = copy.deepcopy(ref).double()
tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
mb, _ = next(model.minibatch_iterator(iter(tensors), device, n_graphs, parallelize=False))
mbc = model_ref.minibatch_to_cpu(mb)
with torch.no_grad():
    r32 = ref._gnn(mbc["graph_data"]["node_data"], mbc["graph_data"]["adjacency_lists"], return_all_states=True).double()
    r64 = ref64._gnn(mbc["graph_data"]["node_data"], mbc["graph_data"]["adjacency_lists"], return_all_states=True)
widths = [hidden] + [hidden, hidden, hidden, hidden, 2 * hidden, hidden] * 2
print(f"hidden={hidden} nodes={r64.shape[0]}; oracle fp32 vs fp64 per state: "
      + " ".join(f"{float((r32[:, sum(widths[:i]):sum(widths[:i + 1])] - r64[:, sum(widths[:i]):sum(widths[:i + 1])]).abs().max()):.1e}"
                 for i in range(len(widths))))

CONFIGS = [
    ("tma (default: TMA projections + TMA node update, operands pre-scaled)", dict(USE_TMA=True, PROJECTION_MODE="f16x3", DENSE_FORWARD_FP32_REFEREE=False, PRESCALE_OPERANDS=True)),
    ("tma, operands NOT pre-scaled (lo parts of the weights are fp16 subnormals)", dict(USE_TMA=True, PROJECTION_MODE="f16x3", DENSE_FORWARD_FP32_REFEREE=False, PRESCALE_OPERANDS=False)),
    ("tma projections, fp32 library GEMM node update", dict(USE_TMA=True, PROJECTION_MODE="f16x3", DENSE_FORWARD_FP32_REFEREE=True)),
    ("first-generation path (round 1)", dict(USE_TMA=False, PROJECTION_MODE="f16x3", DENSE_FORWARD_FP32_REFEREE=False)),
    ("fp32 library GEMMs everywhere (referee)", dict(USE_TMA=False, PROJECTION_MODE="fp32", DENSE_FORWARD_FP32_REFEREE=False)),
]
for label, flags in CONFIGS:
    for k, v in flags.items():
        setattr(ops, k, v)
    g = {k: v for k, v in mb["graph_data"].items() if k != "h2d_bytes"}
    g["adjacency_lists"].plan = None
    g["adjacency_lists"].block_nodes = ops.plan_block_nodes_for(getattr(model.gnn_model, "_mp_layer_dims", ()))
    with torch.no_grad():
        out = nn._gnn(**g, return_all_states=True).output_node_representations.cpu().double()
    per_state, off = [], 0
    for w in widths:
        per_state.append(float((out[:, off:off + w] - r64[:, off:off + w]).abs().max()))
        off += w
    final32 = float((out[:, -hidden:] - r32[:, -hidden:]).abs().max())
    print(json.dumps({"config": label, "block_nodes": g["adjacency_lists"].block_nodes, "max_abs_vs_fp64_per_state": [f"{v:.1e}" for v in per_state],
                      "final_vs_fp64": per_state[-1], "final_vs_fp32_oracle": final32,
                      "oracle_fp32_vs_fp64_final": float((r32[:, -hidden:] - r64[:, -hidden:]).abs().max())}), flush=True)
