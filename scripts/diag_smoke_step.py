"""Diagnostic (GPU only, no oracle): one forward+backward of the smoke() workload; saves every parameter gradient and the
last message-passing layer's output gradient to argv[1].  argv[2:] = files of earlier runs to compare against."""
import copy, os, sys
from pathlib import Path
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "neurips21-self-supervised-bug-detection-and-repair_b200")]
import torch
from buglab.models.modelregistry import load_model
from buglab_b200 import ops
from buglab_b200.synthetic import SyntheticBugLabGenerator

device = torch.device("cuda:0")
torch.manual_seed(0)
gen = SyntheticBugLabGenerator(seed=0, mean_nodes=int(os.environ.get("DIAG_NODES", "400")), min_nodes=60)
data = [gen.sample() for _ in range(8)]
model, _, _ = load_model({"modelName": "gnn-mlp", "hidden_state_size": 256, "dropout_rate": 0.0}, Path("/tmp/diag_smoke.pkl.gz"))
model.gnn_model.node_representation_model.dropout_rate = 0.0
model.compute_metadata(iter(copy.deepcopy(data)))
nn = model.build_neural_module().to(device)
tensors = list(model.tensorize_dataset(iter(copy.deepcopy(data)), parallelize=False))
mb, _ = next(model.minibatch_iterator(iter(tensors), device, 8, parallelize=False))
nn.train()
store = {}
layers = nn._gnn.message_passing_layers
def hook(mod, inp, out):
    out.retain_grad(); store["last"] = out
layers[-1].register_forward_hook(hook)
loss = nn(**mb)
loss.backward()
torch.cuda.synchronize()
result = {n: p.grad.detach().cpu() for n, p in nn.named_parameters() if p.grad is not None}
result["__d_last__"] = store["last"].grad.detach().cpu()
result["__loss__"] = loss.detach().cpu()
torch.save(result, sys.argv[1])
print("saved", sys.argv[1], "loss", float(loss.detach()), "USE_TMA", ops.USE_TMA)
for other in sys.argv[2:]:
    ref = torch.load(other)
    worst = []
    for k, v in result.items():
        d = (v.double() - ref[k].double()).norm() / ref[k].double().norm().clamp_min(1e-300)
        worst.append((float(d), k))
    worst.sort(reverse=True)
    print(f"vs {other}: d_last rel {float((result['__d_last__'].double() - ref['__d_last__'].double()).norm() / ref['__d_last__'].double().norm()):.2e}; "
          f"worst params: " + ", ".join(f"{k[-40:]} {d:.1e}" for d, k in worst[:3]))
