"""Diagnostic (GPU): per-parameter gradient differences vs the CPU oracle and per-layer winner mismatches."""
import copy, os, sys
from pathlib import Path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch
from buglab.models.modelregistry import load_model
from buglab_b200 import ops
from buglab_b200.synthetic import SyntheticBugLabGenerator
from oracle import model_ref
from oracle.mp_ref import edge_messages_ref
from oracle.scatter_ref import scatter_max

hidden = int(sys.argv[1]) if len(sys.argv) > 1 else 128
device = torch.device("cuda:0")
torch.manual_seed(0)
gen = SyntheticBugLabGenerator(seed=0, mean_nodes=400, min_nodes=60)
data = [gen.sample() for _ in range(8)]
model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": hidden, "dropout_rate": 0.0}, Path("/tmp/d.pkl.gz"))
model.gnn_model.node_representation_model.dropout_rate = 0.0
model.compute_metadata(iter(copy.deepcopy(data)))
nn = model.build_neural_module().to(device)
ref = model_ref.GnnBugLabModule(hidden, model.gnn_model.num_edge_types, len(model.gnn_model.node_representation_model.vocabulary), len(model._target_rewrite_ops))
ref.load_state_dict({k: v.cpu() for k, v in nn.state_dict().items()})
ref64 = copy.deepcopy(ref).double()
tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
mb, _ = next(model.minibatch_iterator(iter(tensors), device, 8, parallelize=False))
nn.train()
loss = nn(**mb); loss.backward()
mbc = model_ref.minibatch_to_cpu(mb)
loss_ref = ref(**mbc); loss_ref.backward()
loss64 = ref64(**mbc); loss64.backward()
print("loss gpu/ref32/ref64", float(loss), float(loss_ref), float(loss64))
r32, r64 = dict(ref.named_parameters()), dict(ref64.named_parameters())
rows = []
for n, p in nn.named_parameters():
    if p.grad is None or r32[n].grad is None: continue
    g, a, b = p.grad.cpu().double(), r32[n].grad.double(), r64[n].grad
    rows.append((float((g - a).abs().max()), float((g - b).abs().max()), float((a - b).abs().max()), float(b.abs().max()), n))
rows.sort(reverse=True)
print("max|gpu-ref32|  max|gpu-ref64|  max|ref32-ref64|  max|ref64|  name")
for r in rows[:12]: print("%.3e %.3e %.3e %.3e %s" % r)
# winner mismatches per layer with identical inputs
plan = mb["graph_data"]["adjacency_lists"].plan
adj_cpu = mbc["graph_data"]["adjacency_lists"]
state = ref._gnn._GraphNeuralNetwork__node_embedder(**mbc["graph_data"]["node_data"]).detach()
remembered = None
for i, layer in enumerate(ref._gnn.layers):
    if isinstance(layer, model_ref._NoParams):
        if layer.kind == "remember": remembered = state
        else: state = torch.cat((remembered, state), -1)
        continue
    lw = layer._MlpMessagePassingLayer__edge_message_transformation_layers
    W = torch.stack([l.weight for l in lw]).detach(); b = torch.stack([l.bias for l in lw]).detach()
    msgs, tg = edge_messages_ref(state, adj_cpu, W, b)
    val, arg = scatter_max(msgs, tg, dim=0, dim_size=state.shape[0])
    msgs64, _ = edge_messages_ref(state.double(), adj_cpu, W.double(), b.double())
    val64, arg64 = scatter_max(msgs64, tg, dim=0, dim_size=state.shape[0])
    # product winners
    from buglab_b200 import _lib
    hg, Wg, bg = state.to(device), W.to(device), b.to(device)
    D = state.shape[1]
    u = ops._project_pairs(ops._rows_gather(hg, plan.s_node), Wg, 0, plan.s_type_ptr_host, None)
    v = ops._project_pairs(ops._rows_gather(hg, plan.t_node), Wg, D, plan.t_type_ptr_host, bg)
    M = W.shape[1]; N = state.shape[0]
    agg = torch.empty(N, M, device=device); xw = torch.empty_like(agg); ew = torch.empty(N, M, device=device, dtype=torch.int32)
    _lib.check(_lib.load().bl_edge_segmax_fwd(u.data_ptr(), v.data_ptr(), plan.row_ptr.data_ptr(), plan.urow.data_ptr(), plan.vrow.data_ptr(), N, M, agg.data_ptr(), xw.data_ptr(), ew.data_ptr(), torch.cuda.current_stream().cuda_stream), "f")
    win = plan.e_perm.long().cpu()[ew.long().cpu().clamp(min=0)]
    mism32 = (win != arg); mism64 = (win != arg64); m3264 = (arg != arg64)
    # gap between best and the product's pick, in oracle messages
    gap = (val - msgs.gather(0, win.clamp(max=msgs.shape[0]-1))).abs()
    print(f"layer {i}: D={D} M={M} winners differ vs ref32 {mism32.float().mean():.2e} vs ref64 {mism64.float().mean():.2e}; ref32 vs ref64 {m3264.float().mean():.2e}; max value gap at mismatches {float(gap[mism32].max()) if mism32.any() else 0:.2e}; max|agg diff| {float((agg.cpu()-val).abs().max()):.2e}")
    state = layer(state, adj_cpu).detach()
