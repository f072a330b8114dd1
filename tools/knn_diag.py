"""Per-kernel timing of one kNN call on the tensor-core path (run under `ncu --metrics gpu__time_duration.sum`)."""
import sys
import torch
sys.path.insert(0, ".")
import multilevel_gnn_b200 as m

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
x = torch.randn(n, 64, device="cuda")
m.knn_graph_matrix(x, 16)
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
m.knn_graph_matrix(x, 16)
t1.record()
torch.cuda.synchronize()
print("kNN N=%d D=64 k=16: %.3f ms" % (n, t0.elapsed_time(t1)))
