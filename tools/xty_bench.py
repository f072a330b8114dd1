import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, time
import multilevel_gnn_b200 as m
from multilevel_gnn_b200 import functional as Fn
dev='cuda'
for rows, M in [(313728,64),(313728,32),(627456,64),(2000000,128)]:
    a=torch.randn(rows,M,device=dev); x=torch.randn(rows,128,device=dev)
    for tc in (False, True):
        Fn.USE_TF32X3 = tc
        for _ in range(3): Fn.xty(a,x,True)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): Fn.xty(a,x,True)
        e1.record(); torch.cuda.synchronize()
        ms=e0.elapsed_time(e1)/20
        print(rows,M,'tc' if tc else 'simt', '%.1f us'%(ms*1e3), '%.0f GB/s'%(4*rows*(M+128)/ms/1e6))
print("-- small shapes: xty vs torch matmul")
Fn.USE_TF32X3 = True
def tm(f, n=50):
    for _ in range(5): f()
    torch.cuda.synchronize()
    e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n*1e3
for rows, M, K in [(28032,32,32),(28032,64,32),(32,256,6913),(28032,32,64),(28032,64,64)]:
    a=torch.randn(rows,M,device=dev); x=torch.randn(rows,K,device=dev)
    print(rows,M,K,'xty %.1f us'%tm(lambda: Fn.xty(a,x,True)), 'torch %.1f us'%tm(lambda: (a.t()@x, a.sum(0))))
