"""tcgen05 / TMA bf16 GEMM (mlg_gemm_bf16) and the DiffPool tensor-core path.

Tolerances: operands are rounded to bf16 (8-bit mantissa), accumulation is fp32 in tensor memory.
 (1) against the SAME bf16-rounded operands multiplied in fp64 the only difference is summation order:
     rtol 1e-4 / atol 1e-4*scale;
 (2) against the fp32 product the bf16 rounding of both operands gives a relative error ~2^-8 per term:
     |err| <= 2^-7 * sqrt(K) * rms(a) * rms(b) * 4 is asserted (a loose statistical bound)."""
import math

import pytest
import torch

from conftest import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module")
def gemm():
    from multilevel_gnn_b200 import gemm as g
    return g


@pytest.mark.parametrize("shape", [(128, 128, 64), (256, 384, 512), (130, 70, 200), (1000, 37, 146), (2048, 2048, 4096),
                                   (300, 1024, 1000)])
def test_gemm_bf16_matches_rounded_reference(gemm, shape):
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = torch.randn(K, N, generator=g).to(DEV)
    c = gemm.matmul_bf16(a, b)
    ref = (a.bfloat16().double() @ b.bfloat16().double()).float()
    assert_close(c, ref, rtol=1e-4, atol=1e-4, what="bf16-rounded reference %s" % (shape,))
    err = (c - a @ b).abs().max().item()
    assert err <= 2 ** -7 * math.sqrt(K) * 4, "vs fp32 product: %g" % err


@pytest.mark.parametrize("shape", [(512, 512, 4096, 1), (256, 256, 10000, 1), (2500, 1024, 10000, 1), (300, 260, 100, 1),
                                   (1000, 700, 2100, 3), (2500, 2500, 3000, 1), (257, 513, 64, 5)])
def test_gemm_bf16_stream_k(gemm, shape):
    """M, N >= 256: persistent stream-K grid of SM pairs (mlg_gemm_bf16_ws).  Shapes where tiles are split between
    pairs (few tiles x long K), pairs that span several tiles (many tiles x short K), ragged edges, batches.  Each is
    launched three times on the same workspace (the consumers reset the flags) and must give bit-identical
    results (partials are added in pair order)."""
    M, N, K, B = shape
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(B, M, K, generator=g).to(DEV)
    b = torch.randn(B, K, N, generator=g).to(DEV)
    ref = torch.matmul(a.bfloat16().double(), b.bfloat16().double()).float()
    first = None
    for _ in range(3):
        c = gemm.matmul_bf16(a, b)
        assert_close(c, ref, rtol=1e-4, atol=1e-4, what="stream-K %s" % (shape,))
        if first is None:
            first = c.clone()
        assert torch.equal(c, first)


def test_gemm_bf16_batched_and_grad(gemm):
    g = torch.Generator().manual_seed(5)
    a = torch.randn(3, 200, 320, generator=g).to(DEV).requires_grad_()
    b = torch.randn(3, 320, 150, generator=g).to(DEV).requires_grad_()
    c = gemm.matmul_bf16(a, b)
    ref = torch.matmul(a.detach().bfloat16().double(), b.detach().bfloat16().double()).float()
    assert_close(c, ref, rtol=1e-4, atol=1e-4, what="batched")
    go = torch.randn(c.shape, generator=g).to(DEV)
    ga, gb = torch.autograd.grad(c, [a, b], go)
    ra = torch.matmul(go.bfloat16().double(), b.detach().bfloat16().double().transpose(1, 2)).float()
    rb = torch.matmul(a.detach().bfloat16().double().transpose(1, 2), go.bfloat16().double()).float()
    assert_close(ga, ra, rtol=1e-4, atol=1e-4, what="grad a")
    assert_close(gb, rb, rtol=1e-4, atol=1e-4, what="grad b")
    # shared (2-D) left operand broadcast over the batch, as DiffPool's adj
    adj = torch.randn(200, 200, generator=g).to(DEV)
    x = torch.randn(3, 200, 96, generator=g).to(DEV)
    y = gemm.matmul_bf16(adj, x)
    assert_close(y, torch.matmul(adj.bfloat16().double(), x.bfloat16().double()).float(), rtol=1e-4, atol=1e-4, what="2D x 3D")


@pytest.mark.parametrize("rows,K,N,bias", [(4096, 1024, 2500, True), (2500, 1024, 625, False), (1030, 260, 300, True)])
def test_linear_bf16_and_operand_cache(gemm, rows, K, N, bias):
    """gemm.linear_bf16 (x @ W^T with the weight as the K-major B operand: DenseSAGEConv's projections on the tensor cores)
    against the bf16-rounded fp64 product, forward and both gradients; inside gemm.operand_cache() the bf16 copy of an
    operand is made once and the results are bit-identical to the uncached calls."""
    g = torch.Generator().manual_seed(rows + N)
    x = torch.randn(2, rows // 2, K, generator=g).to(DEV).requires_grad_()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV).requires_grad_()
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    y = gemm.linear_bf16(x, w, b)
    ref = torch.matmul(x.detach().bfloat16().double(), w.detach().bfloat16().double().t())
    ref = (ref + b.double()) if bias else ref
    assert_close(y, ref.float(), rtol=1e-4, atol=1e-4, what="linear_bf16")
    go = torch.randn(y.shape, generator=g).to(DEV)
    gx, gw = torch.autograd.grad(y, [x, w], go)
    rx = torch.matmul(go.bfloat16().double(), w.detach().bfloat16().double()).float()
    rw = torch.matmul(go.reshape(-1, N).bfloat16().double().t(), x.detach().reshape(-1, K).bfloat16().double()).float()
    assert_close(gx, rx, rtol=1e-4, atol=1e-4, what="linear_bf16 grad x")
    assert_close(gw, rw, rtol=1e-4, atol=1e-4 * math.sqrt(rows), what="linear_bf16 grad w")
    calls = []
    real = gemm._cast_now
    gemm._cast_now = lambda src, tr: (calls.append((src.data_ptr(), tr)), real(src, tr))[1]
    try:
        with gemm.operand_cache():
            y1 = gemm.linear_bf16(x.detach(), w.detach(), b)
            y2 = gemm.linear_bf16(x.detach(), w.detach(), b)
            z = gemm.matmul_bf16(x.detach()[0], w.detach().t())          # w transposed: another orientation of the same data
    finally:
        gemm._cast_now = real
    assert torch.equal(y1, y.detach()) and torch.equal(y2, y1)
    assert len([c for c in calls if c[0] == w.data_ptr()]) == 1, "the weight is cast once inside the block"
    assert_close(z, y1[0] - (b if bias else 0), rtol=1e-4, atol=1e-4, what="matmul_bf16 vs linear_bf16")


def test_diffpool_tensor_core_path_close_to_fp32():
    """DiffPool at a size where dense_ops takes the tcgen05 path (N=1024 nodes): outputs within the bf16
    tolerance of the fp32 library path on the same weights."""
    import multilevel_gnn_b200 as m
    from multilevel_gnn_b200 import dense_ops
    torch.manual_seed(0)
    args = m.configs.make_args("lgg")
    n, c = 4096, 1024
    dp = m.DiffPool(c, 2, n, 1, 1024, 1024, args).to(DEV)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, n, c, generator=g).to(DEV)
    a = torch.rand(n, n, generator=g)
    adj = ((a + a.t()) * 0.5 + torch.eye(n)).to(DEV)
    dense_ops.FORCE_FP32 = True
    ref, l0, e0 = dp(x, adj)
    dense_ops.FORCE_FP32 = False
    assert dense_ops.use_tensor_cores(adj, x[0])
    out, l1, e1 = dp(x, adj)
    assert_close(out, ref, rtol=3e-2, atol=3e-2, what="DiffPool bf16 vs fp32")
    assert abs(float(l1) - float(l0)) <= 2e-2 * abs(float(l0)) + 1e-6
    assert abs(float(e1) - float(e0)) <= 2e-2 * abs(float(e0)) + 1e-4



@pytest.mark.parametrize("shape", [(1000, 64, 128), (128, 32, 64), (5000, 128, 64), (12345, 64, 128), (777, 128, 128),
                                   (300, 16, 32), (70000, 32, 128), (4096, 256, 32)])
def test_gemm_tf32x3_is_fp32_accurate(shape):
    """3xTF32 split GEMM (mlg_gemm_tf32x3) vs an fp64 product: relative error far below the rtol-1e-4 parity bar;
    fused bias + LeakyReLU epilogue."""
    from multilevel_gnn_b200 import functional as Fn
    M, N, K = shape
    g = torch.Generator().manual_seed(M + N)
    a = (torch.randn(M, K, generator=g) * torch.rand(M, 1, generator=g) * 3).to(DEV)
    w = torch.randn(N, K, generator=g).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    assert Fn._cabi.lib().mlg_gemm_tf32x3_supported(M, N, K)
    out = Fn.tall_matmul(a, w, b, act=1, slope=0.2)
    ref = torch.nn.functional.leaky_relu(a.double() @ w.double().t() + b.double(), 0.2).float()
    assert_close(out, ref, rtol=2e-5, atol=2e-6, what="tf32x3 %s" % (shape,))
    out2 = Fn.tall_matmul(a, w)
    scale = (a.double().abs() @ w.double().abs().t()).float()       # error bound relative to sum |a||w|
    err = (out2 - (a.double() @ w.double().t()).float()).abs()
    assert float((err / scale).max()) < 3e-6
    # strided A (left half of a wider buffer)
    big = torch.randn(M, 2 * K, generator=g).to(DEV)
    out3 = Fn.tall_matmul(big[:, :K], w)
    assert_close(out3, (big[:, :K].double() @ w.double().t()).float(), rtol=2e-5, atol=2e-5, what="strided A")


@pytest.mark.gpu
def test_tall_matmul_splits_wide_outputs():
    """N = 256 with K = 128 does not fit the resident split weight: two 128-column launches writing column slices."""
    from multilevel_gnn_b200 import functional as Fn, _cabi
    g = torch.Generator().manual_seed(77)
    M, K, N = 100146, 128, 256
    a = torch.randn(M, K, generator=g).to(DEV)
    w = torch.randn(N, K, generator=g).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    assert not _cabi.lib().mlg_gemm_tf32x3_supported(M, N, K) and _cabi.lib().mlg_gemm_tf32x3_supported(M, 128, K)
    _cabi.TIMER = _cabi.KernelTimer()
    out = Fn.tall_matmul(a, w, b, act=1, slope=0.0, tag="wide")
    launches = _cabi.TIMER.summary()["wide"]["launches"]
    _cabi.TIMER = None
    assert launches == 2
    ref = torch.relu(a.double() @ w.double().t() + b.double()).float()
    assert_close(out, ref, rtol=2e-5, atol=2e-5, what="wide tall_matmul")
    # K = 256 -> N = 128: two 64-column launches
    a2 = torch.randn(M, 256, generator=g).to(DEV)
    w2 = torch.randn(128, 256, generator=g).to(DEV)
    assert not _cabi.lib().mlg_gemm_tf32x3_supported(M, 128, 256) and _cabi.lib().mlg_gemm_tf32x3_supported(M, 64, 256)
    out2 = Fn.tall_matmul(a2, w2, b[:128])
    assert_close(out2, (a2.double() @ w2.double().t() + b[:128].double()).float(), rtol=2e-5, atol=4e-5, what="K=256 tall_matmul")
