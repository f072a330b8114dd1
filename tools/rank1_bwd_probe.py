#!/usr/bin/env python
"""Probe for DESIGN.md section 6 item 1: mlg_sage_rank1_bwd_rows (by-row backward of the factored first SAGE layer) ran at
~0.18 ms per launch when its input gradient came straight out of the dX GEMM (mlg_gemm_tf32x3) and at ~0.06 ms when a
library elementwise pass sat in between (profiles/r01_ab_layer2.log).  This script times ONLY that kernel (CUDA events
around it, on the launching stream) after different producers of its input, on the gbm shape:

    python tools/rank1_bwd_probe.py            # one JSON line per case, a few seconds of GPU time
    ncu --set full -k regex:sage_rank1_bwd_rows -c 4 python tools/rank1_bwd_probe.py --reps 1

Cases: input written long ago (L2 flushed), by the tensor-core GEMM, by the GEMM followed by aten leaky_relu_backward,
by the GEMM followed by a device-to-device copy, by an elementwise kernel alone; each with the three mask modes of the
kernel (none / y / sign bits)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32)
    a = ap.parse_args()
    import multilevel_gnn_b200 as m
    from multilevel_gnn_b200 import _cabi, functional as Fn, graph, synth
    dev = torch.device("cuda", 0)
    L = _cabi.lib()
    b = synth.multilevel_batch(batch_size=a.batch, seed=0).to(dev)
    n = b.x.shape[0]
    topo = graph.topology(b.edge_index, n, self_loops=True, edge_weight=b.edge_attr, period=3 * m.MultilevelGNN.GENES)
    fw, n1, B, C = topo.fwd, topo.n_single, topo.replicas, 64
    assert B == a.batch and n == n1 * B
    xs = b.x.reshape(-1).float().contiguous()
    g = torch.Generator(device="cpu").manual_seed(0)
    y = torch.randn(n, C, generator=g).to(dev)                       # the layer's activation (mask source)
    guv = torch.randn(n, C, generator=g).to(dev)                     # dX GEMM operand [gz | A^T gz]
    wt = (torch.randn(C, C, generator=g) * 0.1).to(dev)
    bits = torch.randint(-2 ** 62, 2 ** 62, (n1 * B,), generator=g, dtype=torch.int64).to(dev)
    h = torch.empty(fw.cap, C, device=dev)
    g12 = torch.empty(n1, 2 * C, device=dev)
    gbr = torch.empty(n1, C, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > L2

    def rows(gz, mode):
        _cabi.check(L.mlg_sage_rank1_bwd_rows(
            _cabi.fptr(gz), C, _cabi.fptr(y if mode == "y" else None, True), _cabi.lptr(bits if mode == "bits" else None, True),
            0.2, _cabi.fptr(xs), 0, _cabi.iptr(fw.rowptr), _cabi.iptr(fw.col), _cabi.fptr(topo.fwd_val, True),
            _cabi.iptr(topo.fwd_order, True), n1, C, B, _cabi.fptr(h), _cabi.fptr(g12), 2 * C, _cabi.fptr(gbr),
            _cabi.stream_ptr()), "mlg_sage_rank1_bwd_rows")

    def gemm():
        return Fn.tall_matmul(guv, wt, tag="probe_gemm")

    producers = {
        "cold (written long ago, L2 flushed)": lambda: (flush.zero_(), static)[1],
        "tensor-core GEMM": gemm,
        "GEMM -> aten leaky_relu_backward": lambda: torch.ops.aten.leaky_relu_backward(gemm(), y, 0.2, True),
        "GEMM -> clone": lambda: gemm().clone(),
        "elementwise kernel": lambda: guv * 1.5,
    }
    static = torch.randn(n, C, generator=g).to(dev)
    for name, make in producers.items():
        for mode in ("none", "y", "bits"):
            for _ in range(2):
                rows(make(), mode)
            torch.cuda.synchronize()
            tot = 0.0
            for _ in range(a.reps):
                gz = make()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                rows(gz, mode)
                e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            print(json.dumps({"producer": name, "mask": mode, "rows_kernel_us": round(tot / a.reps * 1e3, 1),
                              "input_ptr_mod_2MB": int(gz.data_ptr() % (2 << 20))}))


if __name__ == "__main__":
    main()
