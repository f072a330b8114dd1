"""Dense contractions of DiffPool (S^T.X, S^T.A.S, A.X, S.S^T).

``matmul(a, b)`` dispatches by size: below ``TENSOR_CORE_MIN`` rows the reference-sized problems
(146 x 37 x 32) are launch-bound and go to cuBLAS in fp32 (bit-for-bit the reference's own arithmetic);
at or above it, on CUDA tensors, the product runs on the hand-written tcgen05 / TMA bf16 GEMM
(``mlg_gemm_bf16``, fp32 accumulation in TMEM) with its own autograd (the two backward products are the
same kernel).  The bf16 path has its own tolerance (stated in tests/test_gpu_parity.py) -- it cannot
meet fp32 rtol 1e-4 and is never used for parity runs at the reference's shapes.
"""
import torch

TENSOR_CORE_MIN = 512       # smallest M, N and K for which the tensor-core path is taken (the second pooling layer of the
                            # N = 10 000 shape has 2 500 nodes -> 625 clusters)
LINEAR_MIN_FEATURES = 256   # nn.Linear projections: tensor cores from this width on (with >= TENSOR_CORE_MIN rows)
FORCE_FP32 = False          # parity switch: keep everything on the fp32 library path


def use_tensor_cores(a, b):
    if FORCE_FP32 or not (a.is_cuda and b.is_cuda):
        return False
    m, k, n = a.shape[-2], a.shape[-1], b.shape[-1]
    return min(m, k, n) >= TENSOR_CORE_MIN


def matmul(a, b):
    """a [..., M, K] @ b [..., K, N] with broadcasting over leading dims."""
    if use_tensor_cores(a, b):
        from . import gemm
        return gemm.matmul_bf16(a, b)
    return torch.matmul(a, b)


def linear(x, lin):
    """``lin(x)`` for an nn.Linear: on the tensor cores (bf16 operands, fp32 accumulate) in the same size regime as
    ``matmul`` -- at N = 10 000 nodes the fp32 SIMT library GEMMs of the two DenseSAGEConv projections were 45 % of
    the large DiffPool forward -- and the module itself below it."""
    rows = x.numel() // x.shape[-1]
    if (not FORCE_FP32 and x.is_cuda and lin.weight.is_cuda and not lin._forward_hooks
            and rows >= TENSOR_CORE_MIN and min(lin.in_features, lin.out_features) >= LINEAR_MIN_FEATURES):
        from . import gemm
        return gemm.linear_bf16(x, lin.weight, lin.bias)
    return lin(x)


def operand_cache():
    from . import gemm
    return gemm.operand_cache()
