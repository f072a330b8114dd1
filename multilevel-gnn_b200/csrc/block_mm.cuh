// Block-level building blocks shared by the fused small-graph kernels (diffpool_fused.cu, decoder_grouped.cu): a
// register-tiled C = alpha A B (+ C) over operands addressed by (row stride, column stride) in shared or global memory,
// and a fixed-tree block sum.
//
// Every phase is written as "parallel for over work items, then barrier" (MLG_PFOR / MLG_SYNC) without warp-level
// primitives, so the SAME source also compiles as plain host C++ (-DMLG_HOST_EMU, tests/ only) where the items of a phase
// run sequentially: the algebra of a kernel built from these blocks is then checked against the CPU oracle without a GPU.
#pragma once
#ifndef MLG_HOST_EMU
#include "common.cuh"
#include "../../include/mlg_b200.h"
#define MLG_DEV __device__ __forceinline__
#define MLG_DEV_CALL static __device__ __noinline__   /* big phase-level routines: one copy per kernel (compile time, code size) */
#define MLG_PFOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define MLG_SYNC() __syncthreads()
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define MLG_DEV static inline
#define MLG_DEV_CALL static
#define MLG_PFOR(i, n) for (int i = 0; i < (n); ++i)
#define MLG_SYNC() ((void)0)
#define __restrict__
#endif


namespace bmm {

// ---------------------------------------------------------------------------------------------------------------------
// Shared-memory matrices are stored with leading dimensions rounded up to a multiple of 4 floats (P4) at 16-byte aligned
// offsets, so that 4 x 4 operand blocks are four 128-bit loads; global operands (x, weights, gradients) keep their natural
// layout and take the vector path only when it happens to be aligned (c = 32 / 64 in the shipped sizes).
MLG_DEV int P4(int v) { return (v + 3) & ~3; }

struct alignas(16) F4 { float x, y, z, w; };
MLG_DEV F4 ld4(const float* p) { return *reinterpret_cast<const F4*>(p); }

// Loads of an operand known to live in shared memory (SH) -- inside an out-of-line routine the pointers are generic, and
// generic loads are measurably slower than ld.shared here (0.63 -> 0.80 ms for the DiffPool forward kernel).
template <bool SH> MLG_DEV float ldf(const float* p) {
#ifndef MLG_HOST_EMU
  if (SH) {
    float v;
    asm("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
  }
#endif
  return *p;
}
template <bool SH> MLG_DEV F4 ldv(const float* p) {
#ifndef MLG_HOST_EMU
  if (SH) {
    F4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
        : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
  }
#endif
  return ld4(p);
}
MLG_DEV bool vec_ok(const float* p, int ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld & 3) == 0; }

// C[M x N] (ldc) = alpha * A[M x K] * B[K x N] (+ C when acc); A, B addressed by (row stride, column stride): a transposed
// operand is just swapped strides.  4 x 4 register tile per work item, 4 k per step.
// AM / BM: how the operand's 4 x 4 block is fetched -- 0: sixteen scalar loads (any strides); 1: contiguous along K (four
// 128-bit loads, one per row / column of the tile); 2: contiguous along the tile's own dimension (M for A, N for B: four
// 128-bit loads, one per k).  Rows / columns past M / N are clamped (mode 0, 1) or read from the row's padding (mode 2) and
// never stored; the K tail (K % 4) runs scalar.
template <int AM, int BM, bool AS, bool BS>
MLG_DEV void mm_tile(float* C, int ldc, const float* A, int rsA, int csA, const float* B, int rsB, int csB, int M, int N, int K,
                     float alpha, bool acc) {
  const int tm = (M + 3) >> 2, tn = (N + 3) >> 2;
  const int K4 = K & ~3;
  MLG_PFOR(t, tm * tn) {
    const int i0 = (t / tn) << 2, j0 = (t % tn) << 2;
    float c[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q) c[r][q] = 0.f;
    const float* ap[4];
    const float* bp[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) ap[r] = A + (i0 + r < M ? i0 + r : i0) * rsA;
#pragma unroll
    for (int q = 0; q < 4; ++q) bp[q] = B + (j0 + q < N ? j0 + q : j0) * csB;
    for (int k = 0; k < K4; k += 4) {
      float a[4][4], bb[4][4];     // a[r][kk], bb[kk][q]
      if (AM == 1) {
#pragma unroll
        for (int r = 0; r < 4; ++r) { const F4 v = ldv<AS>(ap[r] + k); a[r][0] = v.x; a[r][1] = v.y; a[r][2] = v.z; a[r][3] = v.w; }
      } else if (AM == 2) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { const F4 v = ldv<AS>(A + (k + kk) * csA + i0); a[0][kk] = v.x; a[1][kk] = v.y; a[2][kk] = v.z; a[3][kk] = v.w; }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) a[r][kk] = ldf<AS>(ap[r] + (k + kk) * csA);
      }
      if (BM == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) { const F4 v = ldv<BS>(bp[q] + k); bb[0][q] = v.x; bb[1][q] = v.y; bb[2][q] = v.z; bb[3][q] = v.w; }
      } else if (BM == 2) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) { const F4 v = ldv<BS>(B + (k + kk) * rsB + j0); bb[kk][0] = v.x; bb[kk][1] = v.y; bb[kk][2] = v.z; bb[kk][3] = v.w; }
      } else {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int q = 0; q < 4; ++q) bb[kk][q] = ldf<BS>(bp[q] + (k + kk) * rsB);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) c[r][q] = fmaf(a[r][kk], bb[kk][q], c[r][q]);
    }
    for (int k = K4; k < K; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) a[r] = ldf<AS>(ap[r] + k * csA);
#pragma unroll
      for (int q = 0; q < 4; ++q) bb[q] = ldf<BS>(bp[q] + k * rsB);
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int q = 0; q < 4; ++q) c[r][q] = fmaf(a[r], bb[q], c[r][q]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (i0 + r < M && j0 + q < N) {
          float* p = C + (i0 + r) * ldc + j0 + q;
          *p = acc ? fmaf(alpha, c[r][q], *p) : alpha * c[r][q];
        }
  }
  MLG_SYNC();
}

// operand mode from its strides and alignment (uniform over the block).  Mode 2 reads up to 3 elements past the tile's own
// dimension: only allowed for shared-memory operands, whose rows are padded (`padded`), or when that dimension is a
// multiple of 4 anyway.
MLG_DEV int op_mode(const float* p, int s_k, int s_own, bool padded, int own) {
  if (s_k == 1 && vec_ok(p, s_own)) return 1;
  if (s_own == 1 && (padded || (own & 3) == 0) && vec_ok(p, s_k)) return 2;
  return 0;
}

// a_pad / b_pad: the operand lives in shared memory with P4-padded rows (mode 2 allowed, shared-memory loads)
template <bool AS, bool BS>
MLG_DEV void mm_modes(float* C, int ldc, const float* A, int rsA, int csA, const float* B, int rsB, int csB, int M, int N, int K,
                      float alpha, bool acc) {
  const int am = op_mode(A, csA, rsA, AS, M), bm = op_mode(B, rsB, csB, BS, N);
#define MLG_MM(AM_, BM_) mm_tile<AM_, BM_, AS, BS>(C, ldc, A, rsA, csA, B, rsB, csB, M, N, K, alpha, acc)
  if (am == 1) { if (bm == 1) MLG_MM(1, 1); else if (bm == 2) MLG_MM(1, 2); else MLG_MM(1, 0); }
  else if (am == 2) { if (bm == 1) MLG_MM(2, 1); else if (bm == 2) MLG_MM(2, 2); else MLG_MM(2, 0); }
  else { if (bm == 1) MLG_MM(0, 1); else if (bm == 2) MLG_MM(0, 2); else MLG_MM(0, 0); }
#undef MLG_MM
}

MLG_DEV_CALL void mm(float* C, int ldc, const float* A, int rsA, int csA, bool a_pad, const float* B, int rsB, int csB, bool b_pad,
                     int M, int N, int K, float alpha, bool acc) {
  if (a_pad) {
    if (b_pad) mm_modes<true, true>(C, ldc, A, rsA, csA, B, rsB, csB, M, N, K, alpha, acc);
    else mm_modes<true, false>(C, ldc, A, rsA, csA, B, rsB, csB, M, N, K, alpha, acc);
  } else {
    if (b_pad) mm_modes<false, true>(C, ldc, A, rsA, csA, B, rsB, csB, M, N, K, alpha, acc);
    else mm_modes<false, false>(C, ldc, A, rsA, csA, B, rsB, csB, M, N, K, alpha, acc);
  }
}

// sum of v[0..n) -> *dst in a fixed tree (fan-in 16 per level; the levels' partials go to the scratch right behind v, so
// v needs ~1.07 n + 8 floats).  After the closing barrier everyone may read *dst.
MLG_DEV void block_sum(float* v, int n, float* dst) {
  while (n > 16) {
    const int m = (n + 15) >> 4;
    float* w = v + P4(n);
    MLG_PFOR(t, m) {
      const int e = 16 * t + 16 < n ? 16 * t + 16 : n;
      float s = 0.f;
      for (int i = 16 * t; i < e; ++i) s += v[i];
      w[t] = s;
    }
    MLG_SYNC();
    v = w;
    n = m;
  }
  MLG_PFOR(t, 1) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += v[i];
    *dst = s;
  }
  MLG_SYNC();
}

// Row-per-work-item loops walk their row starting at column (row % len): concurrently running rows then touch different
// shared-memory banks although the leading dimensions are multiples of 4 (a plain j = 0.. walk is a 32-way conflict for
// ld = 32).  The order is fixed per row, so results stay deterministic.
#define MLG_ROT(j, jj, i, len) int j = (jj) + (i) % (len); if (j >= (len)) j -= (len)

}  // namespace bmm
