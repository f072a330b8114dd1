"""Where the large DiffPool forward (b=1, N=10000 -> 2500 -> 625, C=1024) spends its time: torch profiler, top CUDA ops."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import multilevel_gnn_b200 as m
dev = "cuda:0"
args = m.configs.make_args("lgg")
torch.manual_seed(0)
big = m.DiffPool(1024, 2, 10000, 2, 1024, 1024, args).to(dev)
xb = torch.randn(1, 10000, 1024, device=dev)
ab = torch.rand(10000, 10000, device=dev)
with torch.no_grad():
    for _ in range(2):
        big(xb, ab)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        big(xb, ab)
    b.record(); torch.cuda.synchronize()
    print("forward ms", a.elapsed_time(b) / 5)
    with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
        big(xb, ab)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=22, max_name_column_width=70))
if "--train" in sys.argv:
    xg = xb.clone().requires_grad_()
    def fb():
        out, l, e = big(xg, ab)
        (out.sum() + l + e).backward()
    for _ in range(2): fb()
    torch.cuda.synchronize(); a.record()
    for _ in range(3): fb()
    b.record(); torch.cuda.synchronize()
    print("forward+backward ms", a.elapsed_time(b) / 3)
