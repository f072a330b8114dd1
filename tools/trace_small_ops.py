"""Which Python line launches each library (ATen) kernel of one eager gbm train step?  torch.profiler with stacks;
prints every CUDA-launching aten op with its innermost repo frames.  Used to hunt the 2 - 4 us elementwise launches
that sit between the hand-written kernels of the captured step."""
import os
import sys
import types

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multilevel_gnn_b200 as m  # noqa: E402
from multilevel_gnn_b200.train import Trainer  # noqa: E402

dev = torch.device("cuda", 0)
args = m.configs.make_args("gbm")
B = args.batch_size
torch.manual_seed(0)
model = m.MultilevelGNN(args)
m.synth.multilevel_params(model)
model.to(dev)
model.pathway_indexs = model.pathway_indexs.to(dev)
raw = m.synth.multilevel_batch(batch_size=B, seed=100)
n1 = 3 * m.MultilevelGNN.GENES
E1 = raw.edge_index.shape[1] // B
topo = m.data.FoldTopology(raw.edge_index[:, :E1], raw.edge_attr[:E1], raw.gene_pca_match[0], raw.raw_indice[0], n1)
patients = [types.SimpleNamespace(x=raw.x[i * n1:(i + 1) * n1], age=raw.age[i], y=raw.y[2 * i:2 * i + 2]) for i in range(B)]
host = m.data.collate(patients, topo, pin=True)
weight = torch.tensor([[0.8, 1.3]]).repeat(B, 1).to(dev)
tr = Trainer(model, args, weight, world_size=1)
resident = m.data.to_device(host, topo, dev)
for _ in range(3):
    tr._step_eager(resident)
torch.cuda.synchronize()
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    tr._step_eager(resident)
    torch.cuda.synchronize()
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
def dev_us(e):
    for a in ("self_device_time_total", "device_time_total", "self_cuda_time_total", "cuda_time_total"):
        v = getattr(e, a, 0) or 0
        if v > 0:
            return float(v)
    return 0.0


allev = list(prof.events())
leaf = [e for e in allev if e.name.startswith("aten::") and not any(c.name.startswith("aten::") for c in (e.cpu_children or []))]
evs = [e for e in leaf if dev_us(e) > 0] or leaf
print("events", len(allev), "leaf aten", len(leaf), "with device time", sum(1 for e in leaf if dev_us(e) > 0))
evs.sort(key=lambda e: e.time_range.start)
for e in evs:
    frames = [f for f in (e.stack or []) if root in f and "trace_small_ops" not in f][:3]
    print("%-34s %6.1f us  %s" % (e.name, dev_us(e), " <- ".join(f.replace(root + "/", "") for f in frames)))
