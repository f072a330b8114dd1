"""Time mlg_gemm_tf32x3 / mlg_xty_tc on the SAGE update shapes (rows = 32 graphs x 15405 nodes)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import multilevel_gnn_b200 as m
from multilevel_gnn_b200 import functional as Fn
dev = "cuda"
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 492960
def tm(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for K, N in [(128, 64), (64, 128), (128, 128), (64, 64)]:
    a = torch.randn(rows, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev)
    ms = tm(lambda: Fn.tall_matmul(a, w, b, act=1, slope=0.2))
    print("tf32x3 rows=%d K=%d N=%d: %.1f us  %.0f GB/s" % (rows, K, N, ms * 1e3, 4 * rows * (K + N) / ms / 1e6))
