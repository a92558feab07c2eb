"""Diagnostic (GPU): per message-passing layer, the smoke() workload's activations and activation gradients against the
CPU oracle under the same forced routing; then the layer-1 operator alone against fp64 on the same graph."""
import copy, json, os, sys
from pathlib import Path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch
from buglab.models.modelregistry import load_model
from buglab_b200 import ops
from buglab_b200.synthetic import SyntheticBugLabGenerator
from oracle import model_ref, parity
from oracle.mp_ref import edge_messages_ref, typed_edge_message_max_ref

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
N = int(mb["graph_data"]["node_to_graph_idx"].shape[0])
print("nodes", N)


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-300)), float((a - b).abs().max()), float(b.abs().max())


def hooked(layers, store):
    handles = []
    for i, layer in enumerate(layers):
        def hook(mod, inp, out, i=i):
            if isinstance(out, torch.Tensor) and out.requires_grad:
                out.retain_grad()
                store[i] = out
        handles.append(layer.register_forward_hook(hook))
    return handles


gpu_layers = nn._gnn.message_passing_layers if hasattr(nn._gnn, "message_passing_layers") else list(getattr(nn._gnn, "_GraphNeuralNetwork__message_passing_layers"))
ref_layers = list(ref._gnn.layers)
for use_tma in (True, False):
    ops.USE_TMA = use_tma
    adj = mb["graph_data"]["adjacency_lists"]
    adj.plan = None
    adj.block_nodes = ops.plan_block_nodes_for(getattr(model.gnn_model, "_mp_layer_dims", ()))
    nn.zero_grad(); ref.zero_grad(); nn.train()
    got, exp = {}, {}
    hs = hooked(gpu_layers, got) + hooked(ref_layers, exp)
    ops.WINNER_TRACE, ops.MINMAX_TRACE = [], []
    loss = nn(**mb)
    winners, ops.WINNER_TRACE = ops.WINNER_TRACE, None
    head_args, ops.MINMAX_TRACE = ops.MINMAX_TRACE, None
    loss.backward()
    torch.cuda.synchronize()
    ref.force_routing(winners, head_args)
    loss_ref = ref(**mb_cpu)
    loss_ref.backward()
    for h in hs:
        h.remove()
    print(f"== USE_TMA={use_tma}  block_nodes={adj.block_nodes}  loss diff {abs(float(loss.detach()) - float(loss_ref.detach())):.1e}")
    for i in sorted(set(got) & set(exp)):
        f = rel(got[i], exp[i])
        g = rel(got[i].grad, exp[i].grad) if got[i].grad is not None and exp[i].grad is not None else (float("nan"),) * 3
        print(f"  layer {i:2d} {type(gpu_layers[i]).__name__:28s} out rel_l2 {f[0]:.1e} max_abs {f[1]:.1e} (scale {f[2]:.1e}) | "
              f"d_out rel_l2 {g[0]:.1e} max_abs {g[1]:.1e} (scale {g[2]:.1e})", flush=True)
ops.USE_TMA = True

# ---- the typed-edge operator alone on this graph (random h, fp64 reference under the kernel's own routing) ----
adj_cpu = [(a[0].cpu().long(), a[1].cpu().long()) for a in mb["graph_data"]["adjacency_lists"]]
K = len(adj_cpu)
for D, M, h_scale in ((hidden, hidden, 1.0), (hidden, hidden, 1.0)):
    for block_nodes in (8192, 0):
        g = torch.Generator().manual_seed(5)
        h = torch.randn(N, D, generator=g) * h_scale
        w = torch.randn(K, M, 2 * D, generator=g) / (2 * D) ** 0.5
        b = torch.randn(K, M, generator=g) * 0.1
        d_out = torch.randn(N, M, generator=g)
        h_ref, w_ref, b_ref = h.double().requires_grad_(), w.double().requires_grad_(), b.double().requires_grad_()
        agg_ref, arg_ref = typed_edge_message_max_ref(h_ref, adj_cpu, w_ref, b_ref)
        plan = ops.build_edge_plan([(s.to(device), t.to(device)) for s, t in adj_cpu], N, block_nodes=block_nodes)
        h_g, w_g, b_g = h.to(device).requires_grad_(), w.to(device).requires_grad_(), b.to(device).requires_grad_()
        ops.WINNER_TRACE = []
        agg = ops.typed_edge_message_max(h_g, w_g, b_g, plan)
        win, ops.WINNER_TRACE = ops.WINNER_TRACE[0], None
        agg.backward(d_out.to(device))
        messages, targets = edge_messages_ref(h_ref, adj_cpu, w_ref, b_ref)
        E = messages.shape[0]
        valid = win < E
        picked = messages.gather(0, win.clamp(max=max(E - 1, 0)))
        torch.where(valid, picked, torch.zeros_like(picked)).backward(d_out.double())
        print(f"== operator alone, block_nodes={block_nodes} P_s={plan.num_s_pairs} P_t={plan.num_t_pairs} E={E}: "
              f"agg {rel(agg, agg_ref)[0]:.1e}  d_h {rel(h_g.grad, h_ref.grad)[0]:.1e}  d_w {rel(w_g.grad, w_ref.grad)[0]:.1e}  "
              f"d_b {rel(b_g.grad, b_ref.grad)[0]:.1e}", flush=True)
        # per type: weight gradient halves
        for k in range(K):
            a = rel(w_g.grad[k, :, :D], w_ref.grad[k, :, :D]); bb = rel(w_g.grad[k, :, D:], w_ref.grad[k, :, D:])
            if max(a[0], bb[0]) > 1e-4:
                print(f"     type {k}: edges {adj_cpu[k][0].numel()}  dA rel {a[0]:.1e}  dB rel {bb[0]:.1e}")

# ---- the dense node-update Linear alone ----
for R in (N, 13168):
    g = torch.Generator().manual_seed(R)
    x = torch.randn(R, hidden, generator=g); wt = torch.randn(hidden, hidden, generator=g) / hidden ** 0.5; dy = torch.randn(R, hidden, generator=g)
    xg, wg = x.to(device).requires_grad_(), wt.to(device).requires_grad_()
    y = ops.dense_linear(xg, wg); y.backward(dy.to(device))
    xr, wr = x.double().requires_grad_(), wt.double().requires_grad_()
    yr = xr @ wr.t(); yr.backward(dy.double())
    print(f"== dense_linear R={R}: y {rel(y, yr)[0]:.1e}  dx {rel(xg.grad, xr.grad)[0]:.1e}  dw {rel(wg.grad, wr.grad)[0]:.1e}", flush=True)
