"""Diagnostic (GPU): the smoke() workload (8 graphs of ~400 nodes, hidden 256) under several backward configurations;
per parameter tensor the relative L2 distance from the oracle's gradient under the same (forced) max-routing."""
import copy, json, os, sys
from pathlib import Path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch
from buglab.models.modelregistry import load_model
from buglab_b200 import ops
from buglab_b200.synthetic import SyntheticBugLabGenerator
from oracle import model_ref, parity

hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 256
mean_nodes = int(sys.argv[2]) if len(sys.argv) > 2 else 400
device = torch.device("cuda:0")
torch.manual_seed(0)
gen = SyntheticBugLabGenerator(seed=0, mean_nodes=mean_nodes, min_nodes=60)
data = [gen.sample() for _ in range(8)]
model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": 0.0}, Path("/tmp/diag_smoke.pkl.gz"))
model.gnn_model.node_representation_model.dropout_rate = 0.0
model.compute_metadata(iter(copy.deepcopy(data)))
nn = model.build_neural_module().to(device)
ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types, len(model.gnn_model.node_representation_model.vocabulary),
                                len(model._target_rewrite_ops))
ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
mb, _ = next(model.minibatch_iterator(iter(tensors), device, 8, parallelize=False))
mb_cpu = model_ref.minibatch_to_cpu(mb)
print("nodes", int(mb["graph_data"]["node_to_graph_idx"].shape[0]), "edges", sum(int(a[0].shape[0]) for a in mb["graph_data"]["adjacency_lists"]))

CONFIGS = [
    ("default", {}),
    ("no side-stream overlap", dict(OVERLAP_EDGE_BACKWARD=False)),
    ("round-1 edge backward (fp32 tables + REDs) feeding the TMA GEMMs", dict(USE_SPLIT_EDGE_BACKWARD=False)),
    ("no operand pre-scaling", dict(PRESCALE_OPERANDS=False)),
    ("type-major plan (no node blocks)", dict(PLAN_BLOCK_NODES=0)),
    ("first-generation kernels (BUGLAB_B200_TMA=0)", dict(USE_TMA=False)),
]
defaults = {k: getattr(ops, k) for _, flags in CONFIGS for k in flags}
for label, flags in CONFIGS:
    for k, v in defaults.items():
        setattr(ops, k, v)
    for k, v in flags.items():
        setattr(ops, k, v)
    adj = mb["graph_data"]["adjacency_lists"]
    adj.plan = None
    adj.block_nodes = ops.plan_block_nodes_for(getattr(model.gnn_model, "_mp_layer_dims", ()))
    nn.zero_grad(); ref.zero_grad(); nn.train()
    ops.WINNER_TRACE, ops.MINMAX_TRACE = [], []
    loss = nn(**mb)
    winners, ops.WINNER_TRACE = ops.WINNER_TRACE, None
    head_args, ops.MINMAX_TRACE = ops.MINMAX_TRACE, None
    loss.backward()
    torch.cuda.synchronize()
    ref.force_routing(winners, head_args)
    loss_ref = ref(**mb_cpu)
    loss_ref.backward()
    ref_grads = dict(ref.named_parameters())
    rows = []
    for n, p in nn.named_parameters():
        if p.grad is None or ref_grads[n].grad is None:
            continue
        frac, l2, mx = parity.grad_mismatch(p.grad, ref_grads[n].grad)
        rows.append((l2, frac, mx, n))
    rows.sort(reverse=True)
    print(json.dumps({"config": label, "block_nodes": adj.block_nodes, "loss_diff": abs(float(loss) - float(loss_ref)),
                      "worst": [{"rel_l2": f"{l2:.1e}", "frac_bad": f"{fr:.1e}", "max_abs": f"{mx:.1e}",
                                 "name": n.replace("_gnn._GraphNeuralNetwork__message_passing_layers.", "L").replace("_MlpMessagePassingLayer__", "")[-70:]}
                                for l2, fr, mx, n in rows[:6]],
                      "median_rel_l2": f"{sorted(r[0] for r in rows)[len(rows) // 2]:.1e}"}), flush=True)
