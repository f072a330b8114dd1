#!/usr/bin/env python
"""Per-component measurement of every hot-path row: one JSON line each with the device time
(CUDA events, warm, inputs > L2 where the shape allows), the roofline fraction against MEASURED_PEAKS.json and
the CPU oracle (oracle/restated.py, torch CPU threads = all cores) on a bounded sample.

    python tools/bench_components.py [--quick] > profiles/r01_components.jsonl
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import multilevel_gnn_b200 as m  # noqa: E402
from multilevel_gnn_b200 import functional as Fn, graph  # noqa: E402
from oracle import restated as R  # noqa: E402

DEV = "cuda:0"
PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) \
    else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}
HBM, BF16 = PEAKS["hbm_gbs"], PEAKS["bf16_tflops"]
FP32_PEAK = 148 * 128 * 2 * 1.965e9 / 1e12      # nominal non-tensor fp32 FMA peak at max clock, TFLOP/s


def gpu_ms(fn, reps=10, warm=2):
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


def cpu_ms(fn, reps=1):
    fn()
    t = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t) / reps * 1e3


def emit(**kw):
    print(json.dumps(kw), flush=True)


def genconv(quick):
    n, k, H = (100000, 16, 128)
    g = torch.Generator().manual_seed(0)
    ei = torch.stack([torch.randint(0, n, (n * k,), generator=g), torch.arange(n).repeat_interleave(k)])
    x, e = torch.randn(n, H, generator=g), torch.randn(n * k, H, generator=g)
    torch.manual_seed(0)
    conv = m.GENConv(H, H, aggr="softmax", learn_t=True, msg_norm=True, encode_edge=False, norm="layer").to(DEV)
    xg, eg, eig = x.to(DEV).requires_grad_(), e.to(DEV).requires_grad_(), ei.to(DEV)
    topo = graph.topology(eig, n)
    topo.bwd
    agg = lambda: Fn.GenAggregate.apply(xg, eg, conv.t, 1.0, None, conv.msg_norm.msg_scale, topo, "softmax", 1e-7,
                                        Fn.EPI_MSGNORM, True)
    ms_f = gpu_ms(agg, reps=20)
    go = torch.randn(n, H, device=DEV)
    ms_fb = gpu_ms(lambda: torch.autograd.grad(agg(), [xg, eg, conv.t, conv.msg_norm.msg_scale], go), reps=10)
    bf = 4 * H * (n * k + 2 * n) + 4 * n * k + 4 * (n + 1)
    bb = 4 * H * (2 * n * k + 3 * n) + 8 * n * k
    ns = 10000 if quick else 25000                      # CPU sample: first ns nodes' rows
    sub = ei[:, :ns * k]
    sub = torch.stack([sub[0] % ns, sub[1]])
    sd = {kk: v.detach().cpu() for kk, v in conv.state_dict().items()}
    xs, es = x[:ns], e[:ns * k]
    c_ms = cpu_ms(lambda: R.msg_norm(xs, R.gen_aggregate(R.gen_message(xs, sub, es), sub[1], ns, "softmax", sd["t"], False),
                                     sd["msg_norm.msg_scale"]))
    emit(row="a1-a4 GENConv message+softmax aggregate+MsgNorm+residual (fwd)", shape="N=100k k=16 H=128", ms=round(ms_f, 4),
         bound="hbm", achieved_GBps=round(bf / ms_f / 1e6, 1), peak_GBps=HBM, frac=round(bf / ms_f / 1e6 / HBM, 4),
         cpu_oracle_ms=round(c_ms * n / ns, 1), cpu_sample="%d of %d rows, scaled" % (ns, n), cores=torch.get_num_threads())
    emit(row="a14 GENConv aggregate backward (edge + node + t + msg_scale grads)", shape="N=100k k=16 H=128",
         ms=round(ms_fb - ms_f, 4), bound="hbm", achieved_GBps=round(bb / (ms_fb - ms_f) / 1e6, 1), peak_GBps=HBM,
         frac=round(bb / (ms_fb - ms_f) / 1e6 / HBM, 4))


def sage_and_pool(quick):
    B = 32
    args = m.configs.make_args("gbm")
    b = m.synth.multilevel_batch(batch_size=B, seed=0)
    bd = b.to(DEV)
    n = b.x.shape[0]
    torch.manual_seed(0)
    conv = m.GraphConv(64, 64, conv="sage", act="leakyrelu", mlp_norm="none").to(DEV)
    x = torch.randn(n, 64, device=DEV, requires_grad=True)
    graph.topology(bd.edge_index, n, self_loops=True, edge_weight=bd.edge_attr, period=15405)
    f = lambda: conv(x, bd.edge_index, bd.edge_attr)
    ms_f = gpu_ms(f)
    go = torch.randn(n, 64, device=DEV)
    ms_fb = gpu_ms(lambda: torch.autograd.grad(f(), [x, conv.gconv.lin_r.weight, conv.gconv.nn[0].weight], go))
    bs = 4 if quick else 8
    small = m.synth.multilevel_batch(batch_size=bs, seed=0)
    sd = {k: v.detach().cpu() for k, v in conv.state_dict().items()}
    xc = torch.randn(small.x.shape[0], 64)
    c_ms = cpu_ms(lambda: R.sage_forward(sd, xc, small.edge_index, small.edge_attr))
    emit(row="a6/a7 SAGE layer 64->64 (aggregate + folded GEMM + bias/LeakyReLU), fwd", shape="gbm B=32 N=15405 E'=107835/graph",
         ms=round(ms_f, 4), fwd_bwd_ms=round(ms_fb, 4), bound="hbm+fp32 GEMM", cpu_oracle_ms=round(c_ms * B / bs, 1),
         cpu_sample="%d of %d graphs, scaled" % (bs, B), cores=torch.get_num_threads())
    # pool
    G, P, C = 25015, 2, 32
    lay = graph.pool_layout(bd.gene_pca_match, bd.raw_indice, 15405, 438)
    xp = torch.randn(n, C, device=DEV, requires_grad=True)
    w = torch.randn(G, P, device=DEV, requires_grad=True)
    vm = bd.x.reshape(-1).contiguous()
    pf = lambda: Fn.PathwayPool.apply(xp, w, vm, lay)
    ms_p = gpu_ms(pf)
    gp = torch.randn(B, C, 438, P, device=DEV)
    ms_pb = gpu_ms(lambda: torch.autograd.grad(pf(), [xp, w], gp))
    pb = 4 * C * B * n // B + 4 * n + 12 * G + 4 * B * C * 438 * P
    wc, mk = torch.randn(G, P), torch.ones(G, 1)
    xs = torch.randn(small.x.shape[0], C) * small.x
    c_ms = cpu_ms(lambda: R.multilevel_pool(xs, small.gene_pca_match, small.raw_indice, wc, mk, 15405, 438))
    emit(row="a8 gene->pathway pool (value mask + gather + project + segment sum), fwd", shape="gbm B=32 G=25015 P=2 C=32",
         ms=round(ms_p, 4), fwd_bwd_ms=round(ms_pb, 4), bound="hbm", achieved_GBps=round(pb / ms_p / 1e6, 1), peak_GBps=HBM,
         frac=round(pb / ms_p / 1e6 / HBM, 4), cpu_oracle_ms=round(c_ms * B / bs, 1), cpu_sample="%d of %d graphs, scaled" % (bs, B))


def knn(quick):
    for (n, d) in [(10000, 1024), (100000, 64)]:
        x = torch.randn(n, d, device=DEV)
        ms = gpu_ms(lambda: m.knn_graph_matrix(x, 16), reps=2, warm=1)
        fl = 2.0 * n * n * d
        ns = 2000 if quick else 4000
        xc = x[:ns].cpu()
        c_ms = cpu_ms(lambda: R.knn_graph_matrix(xc, 16))
        path = ("3xTF32 tcgen05 candidate search + fp32 re-evaluation / certificate" if n >= 4096 and d <= 128
                else "tiled fp32 distance + warp top-k")
        emit(row="a10 kNN graph (%s)" % path, shape="N=%d D=%d k=16" % (n, d), ms=round(ms, 3),
             bound="fp32-fma", achieved_TFLOPs=round(fl / ms / 1e9, 2), peak_TFLOPs=round(FP32_PEAK, 1),
             frac=round(fl / ms / 1e9 / FP32_PEAK, 4), cpu_oracle_ms=round(c_ms * (n / ns) ** 2, 1),
             cpu_sample="%d points, scaled by (N/%d)^2 (the reference materialises [N,N]: %.1f GB at this N)" % (ns, ns, n * n * 4 / 1e9))


def diffpool(quick):
    args = m.configs.make_args("lgg")
    torch.manual_seed(0)
    dp = m.DiffPool(32, 2, 146, 2, 32, 64, args).to(DEV)
    x, adj = m.synth.diffpool_inputs(576, 146, 32)
    xg, ag = x.to(DEV).requires_grad_(), adj.to(DEV)

    def fb():
        out, l, e = dp(xg, ag)
        (out.sum() + l + e).backward()
    ms = gpu_ms(fb, reps=5)
    sd = {k: v.detach().cpu() for k, v in dp.state_dict().items()}
    bs = 64 if quick else 144
    c_ms = cpu_ms(lambda: R.diffpool_forward(sd, x[:bs], adj))
    emit(row="a11 DiffPool at the reference size (one fused fp32 kernel per direction, a sample per CTA)", shape="b=576 N=146->37->10 C=32->32->64",
         fwd_bwd_ms=round(ms, 3), cpu_oracle_fwd_ms=round(c_ms * 576 / bs, 1),
         cpu_sample="%d of 576 samples, scaled" % bs)
    import ctypes
    from multilevel_gnn_b200 import _cabi
    L = _cabi.lib()
    for (M, N, K, what) in [(2500, 10000, 10000, "S^T.A"), (2500, 2500, 10000, "(S^T.A).S"), (10000, 1024, 10000, "A.X"),
                            (2500, 1024, 10000, "S^T.X")]:
        A = torch.randn(M, K, device=DEV).bfloat16()
        Bm = torch.randn(N, K, device=DEV).bfloat16()
        C = torch.empty(M, N, device=DEV)
        ws = torch.zeros(L.mlg_gemm_bf16_workspace_bytes(), dtype=torch.uint8, device=A.device)   # stream-K partials + flags
        run = lambda: _cabi.check(L.mlg_gemm_bf16_ws(ctypes.c_void_p(A.data_ptr()), K, 0, ctypes.c_void_p(Bm.data_ptr()), K, 0,
                                                     _cabi.fptr(C), N, 0, M, N, K, 1, 1.0, ctypes.c_void_p(ws.data_ptr()),
                                                     ws.numel(), _cabi.stream_ptr()), "mlg_gemm_bf16_ws")
        ms = gpu_ms(run, reps=10)
        tf = 2.0 * M * N * K / ms / 1e9
        emit(row="a11 DiffPool contraction %s on tcgen05 (bf16 in, fp32 accumulate)" % what, shape="M=%d N=%d K=%d" % (M, N, K),
             ms=round(ms, 4), bound="tensor", achieved_TFLOPs=round(tf, 1), peak_TFLOPs=BF16, frac=round(tf / BF16, 4))


def deepergcn(quick):
    n, k, H, layers = 100000, 16, 128, (4 if quick else 28)
    pts = m.synth.knn_points(n, 64).to(DEV)
    ei = m.knn_graph_matrix(pts, k)
    args = m.configs.deepergcn_args(hidden=H, layers=layers)
    torch.manual_seed(0)
    model = m.DeeperGCN(args).to(DEV)
    b = m.synth.deepergcn_batch(ei.cpu(), n)
    bd = b.to(DEV)
    bd.node_size = b.node_size

    def fb():
        model.zero_grad(set_to_none=True)
        out = model(bd)
        out[:, 0].sum().backward()
    for affine in (True, False):
        model.AFFINE_EDGE = affine
        ms = gpu_ms(fb, reps=2, warm=1)
        emit(row="a5 DeeperGCN train step (res+, softmax, learn_t, msg_norm, per-layer edge encoder), edge term %s"
             % ("factored a_e*p+q (no [E,H] tensor)" if affine else "materialised [E,H] + per-layer edge GEMM"),
             shape="N=100k(+146) k=16 H=128 L=%d" % layers, fwd_bwd_ms=round(ms, 2))
        torch.cuda.empty_cache()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    m._cabi.lib()
    for fn in (genconv, sage_and_pool, knn, diffpool, deepergcn):
        if a.only and a.only not in fn.__name__:
            continue
        fn(a.quick)
        torch.cuda.empty_cache()
