#!/usr/bin/env python
"""Per-kernel microbenchmarks (CUDA events, L2-exceeding inputs) -- used for ncu captures and tuning.

    python tools/microbench.py genconv --H 128 [--bwd]
    python tools/microbench.py sage | pool | knn --n 10000 --d 64
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import multilevel_gnn_b200 as m  # noqa: E402
from multilevel_gnn_b200 import functional as Fn, graph  # noqa: E402

DEV = "cuda:0"
HBM = 6454.0


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def genconv(a):
    n, k, H = a.n, a.k, a.H
    g = torch.Generator().manual_seed(0)
    pts = None
    src = torch.randint(0, max(1, int(n * a.src_frac)), (n * k,), generator=g)   # --src-frac < 1: smaller gather / reduce hot set
    dst = torch.arange(n).repeat_interleave(k)
    ei = torch.stack([src, dst]).to(DEV)
    x = torch.randn(n, H, generator=g).to(DEV).requires_grad_()
    e = torch.randn(n * k, H, generator=g).to(DEV).requires_grad_()
    topo = graph.topology(ei, n)
    topo.bwd
    scale = torch.ones(1, device=DEV, requires_grad=True)
    t = torch.ones(1, device=DEV, requires_grad=True)
    fwd = lambda: Fn.GenAggregate.apply(x, e, t, 1.0, None, scale, topo, a.aggr, 1e-7, Fn.EPI_MSGNORM, True)
    ms = timeit(fwd)
    nb = 4 * H * (n * k + 2 * n) + 4 * n * k + 4 * (n + 1)
    out = {"kernel": "gen_aggr_fwd", "ms": ms, "GBps": nb / ms / 1e6, "frac": nb / ms / 1e6 / HBM}
    with torch.no_grad():
        ms_inf = timeit(fwd)
    out.update({"inference_ms": ms_inf, "inference_GBps": nb / ms_inf / 1e6, "inference_frac": nb / ms_inf / 1e6 / HBM})
    with torch.no_grad():
        ms_x = timeit(lambda: Fn.GenAggregate.apply(x, None, t, 1.0, None, scale, topo, a.aggr, 1e-7, Fn.EPI_MSGNORM, True))
        ms_e = timeit(lambda: Fn.GenAggregate.apply(None, e, t, 1.0, None, None, topo, a.aggr, 0.0, Fn.EPI_NONE, True))
    out.update({"x_only_ms": ms_x, "x_only_L2_GBps": 4 * H * n * k / ms_x / 1e6, "e_only_ms": ms_e,
                "e_only_GBps": 4 * H * n * k / ms_e / 1e6})
    if a.bwd:
        h = fwd()
        gout = torch.randn_like(h)
        ms_fb = timeit(lambda: torch.autograd.grad(fwd(), [x, e, t, scale], gout))
        nbb = 4 * H * (2 * n * k + 3 * n) + 8 * n * k
        out["fwd+bwd_ms"] = ms_fb
        out["bwd_ms_est"] = ms_fb - ms
        out["bwd_GBps_alg"] = nbb / (ms_fb - ms) / 1e6
    print(json.dumps(out))


def sage(a):
    b = m.synth.multilevel_batch(batch_size=a.B, seed=0).to(DEV)
    n = b.x.shape[0]
    C = a.H
    x = torch.randn(n, C, device=DEV, requires_grad=True)
    topo = graph.Topology(b.edge_index, n, self_loops=True, edge_weight=b.edge_attr, period=None if a.generic else 15405)
    print("replicas", topo.replicas)
    topo.bwd, topo.bwd_val, topo.inv_cnt
    fwd = lambda: Fn.SageAggregate.apply(x, topo, False)
    ms = timeit(fwd)
    nnz = topo.fwd.cap
    nb = 8 * C * n + 8 * nnz
    out = {"kernel": "sage_aggr_fwd", "ms": ms, "GBps": nb / ms / 1e6, "frac": nb / ms / 1e6 / HBM}
    y = fwd()
    gout = torch.randn_like(y)
    ms2 = timeit(lambda: torch.autograd.grad(fwd(), [x], gout))
    out["bwd_ms_est"] = ms2 - ms
    print(json.dumps(out))


def knn(a):
    x = torch.randn(a.n, a.d, device=DEV)
    ms = timeit(lambda: m.knn_graph_matrix(x, a.k), reps=3, warm=1)
    fl = 2.0 * a.n * a.n * a.d
    print(json.dumps({"kernel": "knn", "n": a.n, "d": a.d, "k": a.k, "ms": ms, "TFLOPs_fp32": fl / ms / 1e9}))


def gemm(a):
    """tcgen05 bf16 GEMM on pre-cast K-major operands: C[M,N] = A[M,K] . B[N,K]^T."""
    import ctypes
    from multilevel_gnn_b200 import _cabi
    L = _cabi.lib()
    for (M, N, K, what) in [(2500, 10000, 10000, "S^T.A"), (2500, 1024, 10000, "S^T.X"), (2500, 2500, 10000, "(S^T.A).S"),
                            (10000, 1024, 10000, "A.X"), (10000, 10000, 2504, "S.S^T (K padded to 8)"), (8192, 8192, 8192, "square")]:
        A = torch.randn(M, K, device=DEV).bfloat16()
        B = torch.randn(N, K, device=DEV).bfloat16()
        C = torch.empty(M, N, device=DEV)
        ws = torch.zeros(L.mlg_gemm_bf16_workspace_bytes(), dtype=torch.uint8, device=DEV)
        run = lambda: _cabi.check(L.mlg_gemm_bf16_ws(ctypes.c_void_p(A.data_ptr()), K, 0, ctypes.c_void_p(B.data_ptr()), K, 0,
                                                     _cabi.fptr(C), N, 0, M, N, K, 1, 1.0, ctypes.c_void_p(ws.data_ptr()),
                                                     ws.numel(), _cabi.stream_ptr()), "mlg_gemm_bf16_ws")
        ms = timeit(run, reps=10, warm=2)
        ref = timeit(lambda: torch.matmul(A, B.t()), reps=10, warm=2)
        tf = 2.0 * M * N * K / ms / 1e9
        print(json.dumps({"kernel": "gemm_bf16", "what": what, "M": M, "N": N, "K": K, "ms": round(ms, 4), "TFLOPs": round(tf, 1),
                          "frac_of_1650": round(tf / 1650.5, 3), "cublas_bf16_ms": round(ref, 4),
                          "cublas_TFLOPs": round(2.0 * M * N * K / ref / 1e9, 1)}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["genconv", "sage", "knn", "gemm"])
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--k", type=int, default=16)
    ap.add_argument("--H", type=int, default=128)
    ap.add_argument("--d", type=int, default=64)
    ap.add_argument("--B", type=int, default=32)
    ap.add_argument("--aggr", default="softmax")
    ap.add_argument("--bwd", action="store_true")
    ap.add_argument("--src-frac", type=float, default=1.0)
    ap.add_argument("--generic", action="store_true")
    a = ap.parse_args()
    {"genconv": genconv, "sage": sage, "knn": knn, "gemm": gemm}[a.what](a)
