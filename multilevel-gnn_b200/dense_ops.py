"""Dense contractions of DiffPool (S^T.X, S^T.A.S, A.X, S.S^T).

``matmul(a, b)`` dispatches by size: below ``TENSOR_CORE_MIN`` rows the reference-sized problems
(146 x 37 x 32) are launch-bound and go to cuBLAS in fp32 (bit-for-bit the reference's own arithmetic);
at or above it, on CUDA tensors, the product runs on the hand-written tcgen05 / TMA bf16 GEMM
(``mlg_gemm_bf16``, fp32 accumulation in TMEM) with its own autograd (the two backward products are the
same kernel).  The bf16 path has its own tolerance (stated in tests/test_gpu_parity.py) -- it cannot
meet fp32 rtol 1e-4 and is never used for parity runs at the reference's shapes.
"""
import torch

TENSOR_CORE_MIN = 1024      # smallest M, N and K for which the tensor-core path is taken
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
