"""Tensor-core matmul for DiffPool's dense contractions: fp32 in / fp32 out, bf16 operands, fp32 accumulate.

``matmul_bf16(a, b)`` = a @ b with a [..., M, K], b [..., K, N] (leading dims broadcast like torch.matmul for
the 2-D x 3-D and 3-D x 3-D cases DiffPool uses).  The kernel wants both operands K-major, so a is cast as is and
b is cast + transposed (one pass each, ``mlg_cast_bf16``).  Backward: dA = dC @ b^T and dB = a^T @ dC are the same
kernel (fp32 gradients, bf16 operands)."""
import ctypes

import torch

from . import _cabi


def _pad8(n):
    return (n + 7) // 8 * 8


_OPERANDS = None     # {key: (source tensor kept alive, bf16 operand, ld)} while an operand_cache() block is open


class operand_cache:
    """``with gemm.operand_cache():`` -- inside the block the K-major bf16 copy of an fp32 operand is made once per
    (tensor, version, orientation) and shared by every product that reads it: DiffPool multiplies the SAME adjacency
    four times per layer (A.X for the assignment and the embedding conv, S^T.A, and A's transpose in backward), and each
    cast of a 10 000 x 10 000 matrix is a 600 MB pass.  Re-entrant; the copies are dropped when the outermost block ends."""

    def __enter__(self):
        global _OPERANDS
        self.outer = _OPERANDS
        if _OPERANDS is None:
            _OPERANDS = {}
        return self

    def __exit__(self, *exc):
        global _OPERANDS
        _OPERANDS = self.outer
        return False


def _cast(src, transpose):
    """fp32 [b, R, C] (rows contiguous) -> bf16 K-major [b, R, ld] (or [b, C, ld] when transposed)."""
    if _OPERANDS is not None:
        key = (src.data_ptr(), tuple(src.shape), tuple(src.stride()), src._version, bool(transpose))
        hit = _OPERANDS.get(key)
        if hit is None:
            hit = _OPERANDS[key] = (src,) + _cast_now(src, transpose)
        return hit[1], hit[2]
    return _cast_now(src, transpose)


def _cast_now(src, transpose):
    L = _cabi.lib()
    b, R, C = src.shape
    rows, cols = (C, R) if transpose else (R, C)
    ld = _pad8(cols)
    dst = torch.zeros(b, rows, ld, dtype=torch.bfloat16, device=src.device) if ld != cols else \
        torch.empty(b, rows, ld, dtype=torch.bfloat16, device=src.device)
    with torch.cuda.device(src.device):
        _cabi.check(L.mlg_cast_bf16(_cabi.fptr(src), src.stride(1), R, C, b, int(transpose),
                                    ctypes.c_void_p(dst.data_ptr()), ld, _cabi.stream_ptr()), "mlg_cast_bf16")
    return dst, ld


_WORKSPACES = {}


def _workspace(device):
    """Stream-K workspace of the current stream (partial tiles + flags), zeroed once: the kernel's consumers reset
    the flags they read, so it stays reusable across launches on that stream."""
    key = (device.index, torch.cuda.current_stream(device).cuda_stream)
    ws = _WORKSPACES.get(key)
    if ws is None:
        ws = _WORKSPACES[key] = torch.zeros(_cabi.lib().mlg_gemm_bf16_workspace_bytes(), dtype=torch.uint8, device=device)
    return ws


def _gemm_tn(a_bf, lda, b_bf, ldb, M, N, K, batch, a_batched, b_batched):
    """C[b] = A[b] (M x K) . B[b]^T (N x K), operands already bf16 K-major."""
    L = _cabi.lib()
    ws = _workspace(a_bf.device)
    wsp, wsb = ctypes.c_void_p(ws.data_ptr()), ws.numel()
    c = torch.empty(batch, M, N, dtype=torch.float32, device=a_bf.device)
    sa = a_bf.stride(0) if a_batched else 0
    sb = b_bf.stride(0) if b_batched else 0
    flops = 2.0 * M * N * K * batch
    with torch.cuda.device(a_bf.device), _cabi.span("gemm_bf16", flops):
        if batch > 1 and (not a_batched or not b_batched):
            # TMA maps carry one batch stride; a shared operand is looped over (rare: only adj is shared)
            for i in range(batch):
                ai = a_bf[i] if a_batched else a_bf[0]
                bi = b_bf[i] if b_batched else b_bf[0]
                _cabi.check(L.mlg_gemm_bf16_ws(ctypes.c_void_p(ai.data_ptr()), lda, 0, ctypes.c_void_p(bi.data_ptr()), ldb,
                                               0, _cabi.fptr(c[i]), N, 0, M, N, K, 1, 1.0, wsp, wsb, _cabi.stream_ptr()),
                            "mlg_gemm_bf16_ws")
        else:
            _cabi.check(L.mlg_gemm_bf16_ws(ctypes.c_void_p(a_bf.data_ptr()), lda, sa, ctypes.c_void_p(b_bf.data_ptr()), ldb,
                                           sb, _cabi.fptr(c), N, M * N, M, N, K, batch, 1.0, wsp, wsb, _cabi.stream_ptr()),
                        "mlg_gemm_bf16_ws")
    return c


def _as3(t):
    return t.unsqueeze(0) if t.dim() == 2 else t.reshape(-1, t.shape[-2], t.shape[-1])


def _mm(a, b):
    """a [ba, M, K] @ b [bb, K, N] -> [max(ba, bb), M, N] (ba, bb in {1, B})."""
    a3, b3 = _as3(a).float(), _as3(b).float()
    a3 = a3 if a3.stride(2) == 1 else a3.contiguous()
    b3 = b3 if b3.stride(2) == 1 else b3.contiguous()
    batch = max(a3.shape[0], b3.shape[0])
    M, K, N = a3.shape[1], a3.shape[2], b3.shape[2]
    a_bf, lda = _cast(a3, False)
    b_bf, ldb = _cast(b3, True)
    return _gemm_tn(a_bf, lda, b_bf, ldb, M, N, K, batch, a3.shape[0] > 1, b3.shape[0] > 1)


class _MatmulBf16(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _cabi.require_cuda(a, b)
        ctx.save_for_backward(a, b)
        out = _mm(a.detach(), b.detach())
        ctx.out_dim = max(a.dim(), b.dim())
        return out if ctx.out_dim == 3 else out[0]

    @staticmethod
    def backward(ctx, g):
        a, b = ctx.saved_tensors
        g3 = _as3(g)
        ga = gb = None
        if ctx.needs_input_grad[0]:
            ga = _mm(g3, _as3(b).transpose(1, 2))                # [B, M, K]
            ga = ga.sum(0, keepdim=True) if _as3(a).shape[0] == 1 and ga.shape[0] > 1 else ga
            ga = ga.reshape(a.shape)
        if ctx.needs_input_grad[1]:
            gb = _mm(_as3(a).transpose(1, 2), g3)                # [B, K, N]
            gb = gb.sum(0, keepdim=True) if _as3(b).shape[0] == 1 and gb.shape[0] > 1 else gb
            gb = gb.reshape(b.shape)
        return ga, gb


def matmul_bf16(a, b):
    if a.dim() > 3 or b.dim() > 3:
        raise NotImplementedError("matmul_bf16 handles 2-D / 3-D operands")
    return _MatmulBf16.apply(a, b)


class _LinearBf16(torch.autograd.Function):
    """x [..., K] @ w[N, K]^T: the weight is already K-major, so neither operand is transposed on the way in."""

    @staticmethod
    def forward(ctx, x, w):
        _cabi.require_cuda(x, w)
        ctx.save_for_backward(x, w)
        x3 = _as3(x.detach()).float()
        x3 = x3 if x3.stride(2) == 1 and x3.stride(1) >= x3.shape[2] else x3.contiguous()
        w3 = w.detach().float().contiguous().unsqueeze(0)
        rows = x3.shape[0] * x3.shape[1]
        x2 = x3.reshape(1, rows, x3.shape[2])
        a_bf, lda = _cast(x2, False)
        b_bf, ldb = _cast(w3, False)
        out = _gemm_tn(a_bf, lda, b_bf, ldb, rows, w.shape[0], x3.shape[2], 1, False, False)
        return out.reshape(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        g2 = g.reshape(-1, g.shape[-1])
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = _mm(g2, w.detach())[0].reshape(x.shape)                       # [rows, N] @ [N, K]
        if ctx.needs_input_grad[1]:
            gw = _mm(g2.t(), x.detach().reshape(-1, x.shape[-1]))[0]           # [N, rows] @ [rows, K]
        return gx, gw


def linear_bf16(x, weight, bias=None):
    out = _LinearBf16.apply(x, weight)
    return out if bias is None else out + bias
