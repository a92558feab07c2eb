"""Diagnostic (GPU): end-to-end node-state error of the GPU path vs the fp32 and fp64 CPU oracles, per layer."""
import copy, os, sys
from pathlib import Path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch
from buglab.models.modelregistry import load_model
from buglab_b200 import ops
from buglab_b200.synthetic import SyntheticBugLabGenerator
from oracle import model_ref

hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 32
mode = sys.argv[2] if len(sys.argv) > 2 else "f16x3"
ops.PROJECTION_MODE = mode
device = torch.device("cuda:0")
torch.manual_seed(0)
gen = SyntheticBugLabGenerator(seed=0, mean_nodes=250, min_nodes=40)
data = [gen.sample() for _ in range(10)]
model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": 0.0}, Path("/tmp/d.pkl.gz"))
model.gnn_model.node_representation_model.dropout_rate = 0.0
model.compute_metadata(iter(copy.deepcopy(data)))
nn = model.build_neural_module().to(device).train()
ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types, len(model.gnn_model.node_representation_model.vocabulary), len(model._target_rewrite_ops))
ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
ref64 = copy.deepcopy(ref).double()
tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
mb, _ = next(model.minibatch_iterator(iter(tensors), device, 5, parallelize=False))
mbc = model_ref.minibatch_to_cpu(mb)
with torch.no_grad():
    g = {k: v for k, v in mb["graph_data"].items() if k != "h2d_bytes"}
    out = nn._gnn(**g, return_all_states=True).output_node_representations.cpu()
    r32 = ref._gnn(mbc["graph_data"]["node_data"], mbc["graph_data"]["adjacency_lists"], return_all_states=True)
    r64 = ref64._gnn(mbc["graph_data"]["node_data"], mbc["graph_data"]["adjacency_lists"], return_all_states=True)
print(f"hidden={hidden} mode={mode} nodes={out.shape[0]}")
# states: [embed, after each of 12 layer-list entries]; widths: H except after concat (2H)
widths = [hidden] + [hidden, hidden, hidden, hidden, 2 * hidden, hidden] * 2
off = 0
for i, w in enumerate(widths):
    a, b, c = out[:, off:off + w].double(), r32[:, off:off + w].double(), r64[:, off:off + w]
    print(f"state {i:2d} width {w:4d}: max|gpu-ref32| {float((a-b).abs().max()):.2e}  max|gpu-ref64| {float((a-c).abs().max()):.2e}  "
          f"max|ref32-ref64| {float((b-c).abs().max()):.2e}  frac(|gpu-ref32|>1e-4) {float(((a-b).abs()>1e-4).double().mean()):.2e}  rms {float(c.pow(2).mean().sqrt()):.3f}")
    off += w
