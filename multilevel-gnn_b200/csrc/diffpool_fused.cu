// DiffPool at the reference's size as ONE kernel per direction (sm_100a).
//
// Reference: DiffPool.forward (models/diff_pooling.py:116-133) = per layer DiffPoolLayer.forward (:59-65):
//   s = DenseSAGE_pool(x, adj), z = DenseSAGE_embed(x, adj)                (PyG DenseSAGEConv, normalize=True, :24-32)
//   dense_diff_pool(z, adj, s): S = softmax(s); Xp = S^T z; Ap = S^T adj S; link = ||adj - S S^T||_F / numel(adj);
//                               ent = mean_rows(sum_k -S log(S + 1e-15))
//   x = DenseSAGE_after(Xp, Ap)                                             (after_pooling_layer = 1, :28-31,45)
// called from VAE.predict_head (models/vae.py:238-243) with x [b, 146, 32] (b = B*3P = 576 for the lgg shape), ONE shared
// adj [146, 146], clusters 146 -> 37 -> 10, channels 32 -> 32 -> 64.  As library ops that is ~150 launches of
// 146 x 37 x 32-sized GEMMs / pointwise kernels per step (5.0 ms forward + backward on B200: launch-bound).
//
// Here one persistent CTA walks the samples; everything of one sample (A, M = A X / deg, S, Z, A S, pooled X / A of both
// layers) lives in shared memory (<= 227 KB), all contractions are register-tiled (4 x 4 per thread) loops over shared
// memory, and only x, the output [k_last, h_last], four scalars (link / entropy partial sums of both layers) and, in
// backward, dL/dx and a per-CTA partial of the 18 parameter gradients ever touch global memory.  The backward kernel
// recomputes the forward pass of its sample (cheaper than spilling ~200 KB of intermediates per sample) and then walks it
// in reverse; parameter-gradient partials are reduced over the CTAs in a fixed order by a second tiny kernel.
//
// Every phase is written as "parallel for over work items, then barrier" (MLG_PFOR / MLG_SYNC) without warp-level
// primitives, so the SAME source also compiles as plain host C++ (-DMLG_HOST_EMU, tests/ only) where the items of a phase
// run sequentially: the forward and backward algebra is then checked against the CPU oracle without a GPU.
#ifndef MLG_HOST_EMU
#include "common.cuh"
#include "../../include/mlg_b200.h"
#define MLG_DEV __device__ __forceinline__
#define MLG_PFOR(i, n) for (int i = threadIdx.x; i < (n); i += blockDim.x)
#define MLG_SYNC() __syncthreads()
#else
#include <math.h>
#include <stdint.h>
#include <string.h>
#define MLG_DEV static inline
#define MLG_PFOR(i, n) for (int i = 0; i < (n); ++i)
#define MLG_SYNC() ((void)0)
#define __restrict__
#endif

namespace dpf {

constexpr int kMaxLayers = 2;
constexpr float kNormEps = 1e-12f;   // F.normalize eps
constexpr float kEntEps = 1e-15f;    // dense_diff_pool EPS

struct SageW {        // DenseSAGEConv parameters (lin_rel has no bias, lin_root has one)
  const float* rel;   // [out, in]
  const float* root;  // [out, in]
  const float* bias;  // [out]
};

struct LayerDims { int n, c, k, h; };   // nodes, in-channels, clusters, embedding channels

struct Params {
  int layers;
  LayerDims d[kMaxLayers];
  SageW pool[kMaxLayers], embed[kMaxLayers], after[kMaxLayers];
  int grad_off[kMaxLayers][9];   // offsets of (pool rel, root, bias, embed rel, root, bias, after rel, root, bias) in a partial
  int grad_floats;
  const float* x;      // [b, n0, c0]
  const float* adj;    // [n0, n0] shared
  int b;
  float* out;          // [b, k_last, h_last]
  float* stats;        // [b, 2 * layers]: (F, E) per layer: ||A - S S^T||_F^2 and sum -S log(S + eps) of this sample
  // backward
  const float* g_out;  // [b, k_last, h_last]
  const float* coef;   // [2 * layers]: (c_link, c_ent) per layer: g_l / (sqrt(sum_b F) numel(adj)), g_e / (b n)
  float* g_x;          // [b, n0, c0]
  float* partial;      // [gridDim.x, grad_floats]
};

// shared-memory map of one layer (float offsets from the CTA's buffer)
struct LayerMem {
  int A, deg, M, S, Z, rp, re, lse, Xp, Ap, degp, Mp, Xn, ra;   // live from the layer's forward to its backward
  int T1, T2;          // two [n x max(k, h, c)] temporaries of this layer
  int recompute_M;     // layer 0: M = A X / deg sits in T2 during forward (dead once layer 1 reuses that space) and is
                       // recomputed into S's storage for the convolution backward
};
// Layer 0 owns [A | S | Z | small per-layer state]; its temporaries T1 | T2 are dead while layer 1 runs (between layer 0's
// forward and its backward), so ALL of layer 1 -- temporaries and state -- lives inside that region.
struct MemMap {
  LayerMem L[kMaxLayers];
  int small, total;    // scratch for reductions and the small [k x k] / [k x h] gradients
  int small_floats;
};

// ---------------------------------------------------------------------------------------------------------------------
// C[M x N] (ldc) = alpha * A[M x K] * B[K x N] (+ C when acc); A, B addressed by (row stride, column stride): a transposed
// operand is just swapped strides.  TM x TN register tile per work item.
template <int TM, int TN>
MLG_DEV void mm_tile(float* C, int ldc, const float* A, int rsA, int csA, const float* B, int rsB, int csB, int M, int N, int K,
                     float alpha, bool acc) {
  const int tm = (M + TM - 1) / TM, tn = (N + TN - 1) / TN;
  MLG_PFOR(t, tm * tn) {
    const int i0 = (t / tn) * TM, j0 = (t % tn) * TN;
    float c[TM][TN];
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int s = 0; s < TN; ++s) c[r][s] = 0.f;
    const float* ap[TM];
    const float* bp[TN];
#pragma unroll
    for (int r = 0; r < TM; ++r) ap[r] = A + (i0 + r < M ? i0 + r : i0) * rsA;
#pragma unroll
    for (int s = 0; s < TN; ++s) bp[s] = B + (j0 + s < N ? j0 + s : j0) * csB;
#pragma unroll 2
    for (int k = 0; k < K; ++k) {
      float a[TM], bb[TN];
#pragma unroll
      for (int r = 0; r < TM; ++r) a[r] = ap[r][k * csA];
#pragma unroll
      for (int s = 0; s < TN; ++s) bb[s] = bp[s][k * rsB];
#pragma unroll
      for (int r = 0; r < TM; ++r)
#pragma unroll
        for (int s = 0; s < TN; ++s) c[r][s] = fmaf(a[r], bb[s], c[r][s]);
    }
#pragma unroll
    for (int r = 0; r < TM; ++r)
#pragma unroll
      for (int s = 0; s < TN; ++s)
        if (i0 + r < M && j0 + s < N) {
          float* p = C + (i0 + r) * ldc + j0 + s;
          *p = acc ? fmaf(alpha, c[r][s], *p) : alpha * c[r][s];
        }
  }
  MLG_SYNC();
}

// 4 x 4 tiles throughout: smaller tiles for the small pooled products ([37 x 32] over K = 146 keeps 80 of 512 threads busy)
// were measured SLOWER on B200 (b = 576 forward + backward 5.1 ms vs 3.5 ms): the kernel is bound by its instruction count
// (r02 ncu: 229 M warp instructions forward, 41 % of them FFMA), not by idle threads.
MLG_DEV void mm(float* C, int ldc, const float* A, int rsA, int csA, const float* B, int rsB, int csB, int M, int N, int K,
                float alpha, bool acc) {
  mm_tile<4, 4>(C, ldc, A, rsA, csA, B, rsB, csB, M, N, K, alpha, acc);
}

// sum of v[0..n) in index order by one work item -> *dst (after the barrier everyone may read it)
MLG_DEV void ordered_sum(const float* v, int n, float* dst) {
  MLG_PFOR(t, 1) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += v[i];
    *dst = s;
  }
  MLG_SYNC();
}

// ---------------------------------------------------------------------------------------------------------------------
// DenseSAGEConv forward: U = ((A X) / deg) Wrel^T + X Wroot^T + b, then row-normalise.  Leaves M = A X / deg, the row
// norms r (clamped at eps) and Y = U / r.  X: [n x c] (ldx), A: [n x n] (lda), Y: [n x o] (ldy = o).
MLG_DEV void sage_fwd(const float* X, int ldx, const float* A, int lda, const float* deg, int n, int c, int o, const SageW& W,
                      float* M, float* Y, float* r, bool have_M) {
  if (!have_M) {
    mm(M, c, A, lda, 1, X, ldx, 1, n, c, n, 1.f, false);
    MLG_PFOR(t, n * c) M[t] /= deg[t / c];
    MLG_SYNC();
  }
  mm(Y, o, M, c, 1, W.rel, 1, c, n, o, c, 1.f, false);      // M Wrel^T: B(k, j) = Wrel[j][k]
  mm(Y, o, X, ldx, 1, W.root, 1, c, n, o, c, 1.f, true);
  MLG_PFOR(i, n) {
    float ss = 0.f;
    for (int j = 0; j < o; ++j) {
      const float u = Y[i * o + j] + W.bias[j];
      Y[i * o + j] = u;
      ss = fmaf(u, u, ss);
    }
    const float rr = fmaxf(sqrtf(ss), kNormEps);
    r[i] = rr;
    const float inv = 1.f / rr;
    for (int j = 0; j < o; ++j) Y[i * o + j] *= inv;
  }
  MLG_SYNC();
}

MLG_DEV void row_degrees(const float* A, int lda, int n, float* deg) {
  MLG_PFOR(i, n) {
    float s = 0.f;
    for (int j = 0; j < n; ++j) s += A[i * lda + j];
    deg[i] = fmaxf(s, 1.f);
  }
  MLG_SYNC();
}

// One DiffPool layer + its after-pool DenseSAGE, forward.  X [n x c] (ldx) -> Xn [k x h] (in shared memory, the next
// layer's input) ; the layer's pooled adjacency Ap [k x k] is the next layer's A.  stats: (F, E) of this sample.
MLG_DEV void layer_fwd(const Params& P, int l, float* sm, const MemMap& mp, const float* X, int ldx, float* stats) {
  const LayerDims d = P.d[l];
  const LayerMem& m = mp.L[l];
  const int n = d.n, c = d.c, k = d.k, h = d.h;
  float *A = sm + m.A, *deg = sm + m.deg, *M = sm + m.M, *S = sm + m.S, *Z = sm + m.Z;
  float *T1 = sm + m.T1, *small = sm + mp.small;
  sage_fwd(X, ldx, A, n, deg, n, c, k, P.pool[l], M, S, sm + m.rp, false);
  sage_fwd(X, ldx, A, n, deg, n, c, h, P.embed[l], M, Z, sm + m.re, true);
  // softmax over clusters + entropy, per row
  MLG_PFOR(i, n) {
    float mx = -INFINITY;
    for (int j = 0; j < k; ++j) mx = fmaxf(mx, S[i * k + j]);
    float den = 0.f;
    for (int j = 0; j < k; ++j) den += expf(S[i * k + j] - mx);
    const float lse = mx + logf(den);
    sm[m.lse + i] = lse;
    float ent = 0.f;
    for (int j = 0; j < k; ++j) {
      const float s = expf(S[i * k + j] - lse);
      S[i * k + j] = s;
      ent -= s * logf(s + kEntEps);
    }
    small[i] = ent;
  }
  MLG_SYNC();
  ordered_sum(small, n, stats + 1);
  // Xp = S^T Z ; Q = A S ; Ap = S^T Q
  mm(sm + m.Xp, h, S, 1, k, Z, h, 1, k, h, n, 1.f, false);
  mm(T1, k, A, n, 1, S, k, 1, n, k, n, 1.f, false);
  mm(sm + m.Ap, k, S, 1, k, T1, k, 1, k, k, n, 1.f, false);
  // F = ||A - S S^T||_F^2: 4 x 4 tiles of S S^T formed on the fly, one partial per row block, summed in order
  {
    const int tn = (n + 3) >> 2;
    MLG_PFOR(t, tn) small[t] = 0.f;
    MLG_SYNC();
    // each work item owns ONE row block (4 rows) and walks its column tiles: partial[t] has a single writer
    MLG_PFOR(t, tn) {
      const int i0 = t << 2;
      float f = 0.f;
      for (int j0 = 0; j0 < n; j0 += 4) {
        float cc[4][4];
        for (int r = 0; r < 4; ++r)
          for (int s = 0; s < 4; ++s) cc[r][s] = 0.f;
        for (int q = 0; q < k; ++q) {
          float a[4], bb[4];
          for (int r = 0; r < 4; ++r) a[r] = S[(i0 + r < n ? i0 + r : i0) * k + q];
          for (int s = 0; s < 4; ++s) bb[s] = S[(j0 + s < n ? j0 + s : j0) * k + q];
          for (int r = 0; r < 4; ++r)
            for (int s = 0; s < 4; ++s) cc[r][s] = fmaf(a[r], bb[s], cc[r][s]);
        }
        for (int r = 0; r < 4; ++r)
          for (int s = 0; s < 4; ++s)
            if (i0 + r < n && j0 + s < n) {
              const float dlt = A[(i0 + r) * n + j0 + s] - cc[r][s];
              f = fmaf(dlt, dlt, f);
            }
      }
      small[t] = f;
    }
    MLG_SYNC();
    ordered_sum(small, tn, stats + 0);
  }
  // after-pool DenseSAGE on (Xp, Ap)
  row_degrees(sm + m.Ap, k, k, sm + m.degp);
  sage_fwd(sm + m.Xp, h, sm + m.Ap, k, sm + m.degp, k, h, h, P.after[l], sm + m.Mp, sm + m.Xn, sm + m.ra, false);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward helpers
// dU_i = (dY_i - Y_i (Y_i . dY_i)) / r_i   (r_i = max(||U_i||, eps); at the clamp the norm is a constant: dU = dY / eps)
MLG_DEV void normalize_bwd_row(float* dY, const float* Y, float r, int o) {
  float dot = 0.f;
  if (r > kNormEps)
    for (int j = 0; j < o; ++j) dot = fmaf(Y[j], dY[j], dot);
  const float inv = 1.f / r;
  for (int j = 0; j < o; ++j) dY[j] = (dY[j] - Y[j] * dot) * inv;
}

// gW[o x c] += dU^T X ; partial accumulators in global memory (this CTA's slice): single writer per element
MLG_DEV void wgrad(float* gW, const float* dU, int o, const float* X, int ldx, int c, int n) {
  mm(gW, c, dU, 1, o, X, ldx, 1, o, c, n, 1.f, true);
}
MLG_DEV void bgrad(float* gb, const float* dU, int o, int n) {
  MLG_PFOR(j, o) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += dU[i * o + j];
    gb[j] += s;
  }
  MLG_SYNC();
}

// DenseSAGEConv backward given dU [n x o] (gradient at the pre-normalisation output).  Accumulates the parameter
// gradients, adds dL/dX into dX [n x c] (ldd) and, when dA != nullptr, dL/dA (A is itself a function of earlier layers).
// T: [n x c] scratch.  M = A X / deg (from forward), rowsum(A) > 1 <=> deg > 1 (deg == 1 may be the clamp: no gradient).
MLG_DEV void sage_bwd(const float* dU, int o, const float* X, int ldx, const float* A, int lda, const float* deg, const float* M,
                      int n, int c, const SageW& W, float* gWrel, float* gWroot, float* gb, float* T, bool first_into_T,
                      float* dX, int ldd, bool acc_dX, bool finish, float* dA) {
  wgrad(gWrel, dU, o, M, c, c, n);
  wgrad(gWroot, dU, o, X, ldx, c, n);
  bgrad(gb, dU, o, n);
  // T (+)= dU Wrel  (dL/dM, summed over the convolutions that share M) ; dX (+)= dU Wroot
  mm(T, c, dU, o, 1, W.rel, c, 1, n, c, o, 1.f, !first_into_T);
  mm(dX, ldd, dU, o, 1, W.root, c, 1, n, c, o, 1.f, acc_dX);
  if (!finish) return;
  if (dA) {   // through deg: d/d deg_i of (A X)_i / deg_i = -M_i / deg_i, only where the row sum exceeds the clamp
    MLG_PFOR(t, n * n) {
      const int i = t / n;
      if (deg[i] > 1.f) {
        float dot = 0.f;
        for (int q = 0; q < c; ++q) dot = fmaf(T[i * c + q], M[i * c + q], dot);
        dA[i * lda + (t - i * n)] -= dot / deg[i];
      }
    }
    MLG_SYNC();
  }
  MLG_PFOR(t, n * c) T[t] /= deg[t / c];
  MLG_SYNC();
  mm(dX, ldd, A, 1, lda, T, c, 1, n, c, n, 1.f, true);          // A^T T
  if (dA) mm(dA, lda, T, c, 1, X, 1, ldx, n, n, c, 1.f, true);   // T X^T
}

// One DiffPool layer backward.  In: dXn [k x h] = dL/d(after-pool output) (shared memory, overwritten), dAp_in [k x k] =
// dL/dAp from later layers (nullptr: none).  Out: dX [n x c] (ldd; global for layer 0) and, for l > 0, dA [n x n] added.
MLG_DEV void layer_bwd(const Params& P, int l, float* sm, const MemMap& mp, const float* X, int ldx, float* dXn, const float* dAp_in,
                       float* dX, int ldd, float* dA, float* gpart) {
  const LayerDims d = P.d[l];
  const LayerMem& m = mp.L[l];
  const int n = d.n, c = d.c, k = d.k, h = d.h;
  float *A = sm + m.A, *deg = sm + m.deg, *M = sm + m.M, *S = sm + m.S, *Z = sm + m.Z;
  float *T1 = sm + m.T1, *T2 = sm + m.T2, *small = sm + mp.small;
  const float c_link = P.coef[2 * l], c_ent = P.coef[2 * l + 1];
  const int* go = P.grad_off[l];
  // scratch inside `small`: dXp [k x h], dAp [k x k], G [k x k] (later S^T S), Tp [k x h]
  float* dXp = small;
  float* dAp = dXp + k * h;
  float* G = dAp + k * k;
  float* SS = G;
  float* Tp = G + k * k;
  // ---- after-pool DenseSAGE: Xn = normalize(U_a) ----
  MLG_PFOR(i, k) normalize_bwd_row(dXn + i * h, sm + m.Xn + i * h, sm[m.ra + i], h);
  MLG_SYNC();
  MLG_PFOR(t, k * k) dAp[t] = dAp_in ? dAp_in[t] : 0.f;
  MLG_SYNC();
  sage_bwd(dXn, h, sm + m.Xp, h, sm + m.Ap, k, sm + m.degp, sm + m.Mp, k, h, P.after[l], gpart + go[6], gpart + go[7],
           gpart + go[8], Tp, true, dXp, h, false, true, dAp);
  // ---- pooling: Xp = S^T Z, Ap = S^T A S, link, entropy ----
  // row pass: dS_i = Z_i dXp^T (into T2) ; dZ_i = S_i dXp -> dU_e_i (over Z)
  mm(T2, k, Z, h, 1, dXp, 1, h, n, k, h, 1.f, false);
  mm(T1, h, S, k, 1, dXp, h, 1, n, h, k, 1.f, false);     // dZ into T1 (as [n x h])
  MLG_PFOR(i, n) {
    normalize_bwd_row(T1 + i * h, Z + i * h, sm[m.re + i], h);
    for (int j = 0; j < h; ++j) Z[i * h + j] = T1[i * h + j];      // Z now holds dU_e
  }
  MLG_SYNC();
  // G1 = dAp^T - c_link I ; Q = A S (T1) ; dS += Q G1
  MLG_PFOR(t, k * k) G[t] = dAp[(t % k) * k + t / k] - ((t % k) == (t / k) ? c_link : 0.f);
  MLG_SYNC();
  mm(T1, k, A, n, 1, S, k, 1, n, k, n, 1.f, false);
  mm(T2, k, T1, k, 1, G, k, 1, n, k, k, 1.f, true);
  if (dA) {   // dA += S dAp S^T + c_link (A - S S^T)   (A of this layer is the previous layer's pooled adjacency)
    // R = S dAp (reuse T1 after Q is consumed), then dA += R S^T - c_link S S^T + c_link A = (R - c_link S) S^T + c_link A
    mm(T1, k, S, k, 1, dAp, k, 1, n, k, k, 1.f, false);
    MLG_PFOR(t, n * k) T1[t] -= c_link * S[t];
    MLG_SYNC();
    mm(dA, n, T1, k, 1, S, 1, k, n, n, k, 1.f, true);
    MLG_PFOR(t, n * n) dA[t] = fmaf(c_link, A[t], dA[t]);
    MLG_SYNC();
  }
  // G2 = dAp - c_link I ; Q2 = A^T S (T1) ; dS += Q2 G2
  MLG_PFOR(t, k * k) G[t] = dAp[t] - ((t % k) == (t / k) ? c_link : 0.f);
  MLG_SYNC();
  mm(T1, k, A, 1, n, S, k, 1, n, k, n, 1.f, false);
  mm(T2, k, T1, k, 1, G, k, 1, n, k, k, 1.f, true);
  // + 2 c_link S (S^T S)
  mm(SS, k, S, 1, k, S, k, 1, k, k, n, 1.f, false);
  mm(T2, k, S, k, 1, SS, k, 1, n, k, k, 2.f * c_link, true);
  // entropy, softmax backward, normalisation backward: T2 row i -> dU_p_i
  MLG_PFOR(i, n) {
    float* dS = T2 + i * k;
    const float* s = S + i * k;
    float dot = 0.f;
    for (int j = 0; j < k; ++j) {
      dS[j] -= c_ent * (logf(s[j] + kEntEps) + s[j] / (s[j] + kEntEps));
      dot = fmaf(s[j], dS[j], dot);
    }
    const float lse = sm[m.lse + i], r = sm[m.rp + i];
    // dR = S (dS - S.dS) ; S_raw = log S + lse (the normalised pre-softmax row) ; dU_p = normalize_bwd(dR, S_raw, r)
    float dn = 0.f;
    for (int j = 0; j < k; ++j) {
      const float dr = s[j] * (dS[j] - dot);
      dS[j] = dr;
      if (r > kNormEps) dn = fmaf(logf(s[j]) + lse, dr, dn);
    }
    const float inv = 1.f / r;
    for (int j = 0; j < k; ++j) dS[j] = (dS[j] - (logf(s[j]) + lse) * dn) * inv;
  }
  MLG_SYNC();
  // ---- the two DenseSAGE convolutions that produced S and Z (shared M) ----
  if (m.recompute_M) {   // S is dead from here on: its storage takes M = A X / deg again
    M = S;
    mm(M, c, A, n, 1, X, ldx, 1, n, c, n, 1.f, false);
    MLG_PFOR(t, n * c) M[t] /= deg[t / c];
    MLG_SYNC();
  }
  sage_bwd(T2, k, X, ldx, A, n, deg, M, n, c, P.pool[l], gpart + go[0], gpart + go[1], gpart + go[2], T1, true, dX, ldd, false,
           false, nullptr);
  sage_bwd(Z, h, X, ldx, A, n, deg, M, n, c, P.embed[l], gpart + go[3], gpart + go[4], gpart + go[5], T1, false, dX, ldd, true,
           true, dA);
}

// ---------------------------------------------------------------------------------------------------------------------
MLG_DEV void load_adj(const Params& P, float* sm, const MemMap& mp) {
  const int n = P.d[0].n;
  MLG_PFOR(t, n * n) sm[mp.L[0].A + t] = P.adj[t];
  MLG_SYNC();
  row_degrees(sm + mp.L[0].A, n, n, sm + mp.L[0].deg);
}

MLG_DEV void sample_fwd(const Params& P, float* sm, const MemMap& mp, int s, float* stats) {
  const float* X = P.x + (size_t)s * P.d[0].n * P.d[0].c;
  int ldx = P.d[0].c;
  for (int l = 0; l < P.layers; ++l) {
    if (l > 0) row_degrees(sm + mp.L[l].A, P.d[l].n, P.d[l].n, sm + mp.L[l].deg);
    layer_fwd(P, l, sm, mp, X, ldx, stats + 2 * l);
    X = sm + mp.L[l].Xn;
    ldx = P.d[l].h;
  }
}

MLG_DEV void forward_body(const Params& P, float* sm, const MemMap& mp, int cta, int nctas) {
  load_adj(P, sm, mp);
  const LayerDims dl = P.d[P.layers - 1];
  for (int s = cta; s < P.b; s += nctas) {
    sample_fwd(P, sm, mp, s, P.stats + (size_t)s * 2 * P.layers);
    float* o = P.out + (size_t)s * dl.k * dl.h;
    const float* xn = sm + mp.L[P.layers - 1].Xn;
    MLG_PFOR(t, dl.k * dl.h) o[t] = xn[t];
    MLG_SYNC();
  }
}

MLG_DEV void backward_body(const Params& P, float* sm, const MemMap& mp, int cta, int nctas) {
  load_adj(P, sm, mp);
  float* gpart = P.partial + (size_t)cta * P.grad_floats;
  MLG_PFOR(t, P.grad_floats) gpart[t] = 0.f;
  MLG_SYNC();
  const int Lz = P.layers;
  const LayerDims dl = P.d[Lz - 1];
  float* small = sm + mp.small;
  float stats_dummy[2 * kMaxLayers];
  for (int s = cta; s < P.b; s += nctas) {
#ifndef MLG_HOST_EMU
    float* st = small + mp.small_floats - 2 * kMaxLayers;   // forward statistics are not needed again: park them in scratch
#else
    float* st = stats_dummy;
#endif
    (void)stats_dummy;
    sample_fwd(P, sm, mp, s, st);
    // dL/d(output) into the last layer's Xn-shaped gradient buffer: reuse Mp of the last layer's after-pool? no: keep
    // a dedicated [k x h] slot at the end of `small`
    float* dXn = small + mp.small_floats - 2 * kMaxLayers - dl.k * dl.h;
    const float* go = P.g_out + (size_t)s * dl.k * dl.h;
    MLG_PFOR(t, dl.k * dl.h) dXn[t] = go[t];
    MLG_SYNC();
    if (Lz == 1) {
      layer_bwd(P, 0, sm, mp, P.x + (size_t)s * P.d[0].n * P.d[0].c, P.d[0].c, dXn, nullptr,
                P.g_x + (size_t)s * P.d[0].n * P.d[0].c, P.d[0].c, nullptr, gpart);
    } else {
      // layer 1: its input is layer 0's after-pool output Xn0 [k0 x h0], its adjacency layer 0's Ap0 [k0 x k0]
      const LayerDims d0 = P.d[0];
      float* dX1 = small + mp.small_floats - 2 * kMaxLayers - dl.k * dl.h - d0.k * d0.h;     // dL/dXn0
      float* dA1 = dX1 - d0.k * d0.k;                                                          // dL/dAp0
      MLG_PFOR(t, d0.k * d0.k) dA1[t] = 0.f;
      MLG_SYNC();
      layer_bwd(P, 1, sm, mp, sm + mp.L[0].Xn, d0.h, dXn, nullptr, dX1, d0.h, dA1, gpart);
      layer_bwd(P, 0, sm, mp, P.x + (size_t)s * d0.n * d0.c, d0.c, dX1, dA1, P.g_x + (size_t)s * d0.n * d0.c, d0.c, nullptr,
                gpart);
    }
  }
}

// host side: shared-memory map (float offsets).  Layer l > 0 aliases its adjacency onto layer l-1's pooled adjacency.
static inline int build_map(const Params& P, MemMap& mp) {
  int off = 0;
  auto take = [&](int nfl) { const int o = off; off += (nfl + 3) & ~3; return o; };
  auto tsize = [](const LayerDims& d) { const int kk = d.k > d.h ? d.k : d.h; return d.n * (kk > d.c ? kk : d.c); };
  auto state = [&](int l, bool own_M) {
    const LayerDims d = P.d[l];
    LayerMem& m = mp.L[l];
    m.deg = take(d.n);
    if (own_M) m.M = take(d.n * d.c);
    m.S = take(d.n * (own_M ? d.k : (d.k > d.c ? d.k : d.c)));
    m.Z = take(d.n * d.h);
    m.rp = take(d.n);
    m.re = take(d.n);
    m.lse = take(d.n);
    m.Xp = take(d.k * d.h);
    m.Ap = take(d.k * d.k);
    m.degp = take(d.k);
    m.Mp = take(d.k * d.h);
    m.Xn = take(d.k * d.h);
    m.ra = take(d.k);
  };
  int smax = 0;
  for (int l = 0; l < P.layers; ++l) {
    const LayerDims d = P.d[l];
    const int sl = d.n + 2 * d.k * d.h + 2 * d.k * d.k;   // per-row scratch (n), layer_bwd's dXp, dAp, G, Tp
    smax = sl > smax ? sl : smax;
  }
  // layer 0
  mp.L[0].A = take(P.d[0].n * P.d[0].n);
  state(0, false);
  mp.L[0].recompute_M = 1;
  const int region = off;
  mp.L[0].T1 = take(tsize(P.d[0]));
  mp.L[0].T2 = take(tsize(P.d[0]));
  mp.L[0].M = mp.L[0].T2;
  int region_end = off;
  if (P.layers > 1) {          // layer 1 inside layer 0's temporaries
    off = region;
    mp.L[1].A = mp.L[0].Ap;
    mp.L[1].T1 = take(tsize(P.d[1]));
    mp.L[1].T2 = take(tsize(P.d[1]));
    state(1, true);
    mp.L[1].recompute_M = 0;
    if (off > region_end) region_end = off;
  }
  off = region_end;
  // tail of `small`: statistics, dL/d(out), and for two layers dL/dXn0 + dL/dAp0
  int tail = 2 * kMaxLayers + P.d[P.layers - 1].k * P.d[P.layers - 1].h;
  if (P.layers > 1) tail += P.d[0].k * P.d[0].h + P.d[0].k * P.d[0].k;
  mp.small_floats = ((smax + tail) + 3) & ~3;
  mp.small = take(mp.small_floats);
  mp.total = off;
  return off;
}

}  // namespace dpf

#ifndef MLG_HOST_EMU
namespace {

constexpr int kThreadsDP = 512;
constexpr int kMaxSmemBytes = 227 * 1024;

__global__ void __launch_bounds__(kThreadsDP, 1) diffpool_fwd_kernel(const dpf::Params P, const dpf::MemMap mp) {
  extern __shared__ __align__(16) float dp_sm[];
  dpf::forward_body(P, dp_sm, mp, blockIdx.x, gridDim.x);
}

__global__ void __launch_bounds__(kThreadsDP, 1) diffpool_bwd_kernel(const dpf::Params P, const dpf::MemMap mp) {
  extern __shared__ __align__(16) float dp_sm[];
  dpf::backward_body(P, dp_sm, mp, blockIdx.x, gridDim.x);
}

__global__ void diffpool_grad_reduce_kernel(const float* __restrict__ partial, int nctas, int nfl, float* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nfl) return;
  float s = 0.f;
  for (int i = 0; i < nctas; ++i) s += partial[(size_t)i * nfl + t];
  out[t] = s;
}

int fill_params(dpf::Params& P, int64_t layers, const int64_t* dims, const float* const* weights, const float* x,
                const float* adj, int64_t b) {
  memset(&P, 0, sizeof(P));
  P.layers = (int)layers;
  int off = 0;
  for (int l = 0; l < layers; ++l) {
    P.d[l].n = (int)dims[4 * l];
    P.d[l].c = (int)dims[4 * l + 1];
    P.d[l].k = (int)dims[4 * l + 2];
    P.d[l].h = (int)dims[4 * l + 3];
    const float* const* w = weights + 9 * l;
    P.pool[l] = {w[0], w[1], w[2]};
    P.embed[l] = {w[3], w[4], w[5]};
    P.after[l] = {w[6], w[7], w[8]};
    const int n_c = P.d[l].c, k = P.d[l].k, h = P.d[l].h;
    const int sizes[9] = {k * n_c, k * n_c, k, h * n_c, h * n_c, h, h * h, h * h, h};
    for (int q = 0; q < 9; ++q) {
      P.grad_off[l][q] = off;
      off += sizes[q];
    }
  }
  P.grad_floats = off;
  P.x = x;
  P.adj = adj;
  P.b = (int)b;
  return off;
}

int check_dims(int64_t layers, const int64_t* dims) {
  MLG_CHECK_ARG(layers >= 1 && layers <= dpf::kMaxLayers, "mlg_diffpool: 1 or 2 pooling layers are supported (got %lld)", (long long)layers);
  for (int l = 0; l < layers; ++l) {
    MLG_CHECK_ARG(dims[4 * l] >= 1 && dims[4 * l + 1] >= 1 && dims[4 * l + 2] >= 1 && dims[4 * l + 3] >= 1, "mlg_diffpool: bad dims");
    if (l > 0)
      MLG_CHECK_ARG(dims[4 * l] == dims[4 * (l - 1) + 2] && dims[4 * l + 1] == dims[4 * (l - 1) + 3],
                    "mlg_diffpool: layer %d must take layer %d's clusters / channels", l, l - 1);
  }
  return MLG_OK;
}

}  // namespace

extern "C" int64_t mlg_diffpool_smem_bytes(int64_t layers, const int64_t* dims) {
  if (!dims || layers < 1 || layers > dpf::kMaxLayers) return -1;
  dpf::Params P;
  memset(&P, 0, sizeof(P));
  P.layers = (int)layers;
  for (int l = 0; l < layers; ++l) P.d[l] = {(int)dims[4 * l], (int)dims[4 * l + 1], (int)dims[4 * l + 2], (int)dims[4 * l + 3]};
  dpf::MemMap mp;
  return (int64_t)dpf::build_map(P, mp) * 4;
}

extern "C" int mlg_diffpool_supported(int64_t layers, const int64_t* dims) {
  const int64_t bytes = mlg_diffpool_smem_bytes(layers, dims);
  return bytes > 0 && bytes <= kMaxSmemBytes;
}

extern "C" int64_t mlg_diffpool_grad_floats(int64_t layers, const int64_t* dims) {
  int64_t tot = 0;
  for (int l = 0; l < layers; ++l) {
    const int64_t c = dims[4 * l + 1], k = dims[4 * l + 2], h = dims[4 * l + 3];
    tot += 2 * k * c + k + 2 * h * c + h + 2 * h * h + h;
  }
  return tot;
}

extern "C" int64_t mlg_diffpool_ctas(int64_t b) { return b < 148 ? (b < 1 ? 1 : b) : 148; }

extern "C" int mlg_diffpool_fwd(const float* x, const float* adj, const float* const* weights, int64_t layers,
                                const int64_t* dims, int64_t b, float* out, float* stats, void* stream) {
  MLG_CHECK_ARG(x && adj && weights && dims && out && stats && b >= 1, "mlg_diffpool_fwd: bad arguments");
  if (int rc = check_dims(layers, dims)) return rc;
  MLG_CHECK_ARG(mlg_diffpool_supported(layers, dims), "mlg_diffpool_fwd: %lld bytes of shared memory needed (limit %d)",
                (long long)mlg_diffpool_smem_bytes(layers, dims), kMaxSmemBytes);
  dpf::Params P;
  fill_params(P, layers, dims, weights, x, adj, b);
  P.out = out;
  P.stats = stats;
  dpf::MemMap mp;
  const size_t smem = (size_t)dpf::build_map(P, mp) * 4;
  MLG_CUDA(cudaFuncSetAttribute(diffpool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  diffpool_fwd_kernel<<<(unsigned)mlg_diffpool_ctas(b), kThreadsDP, smem, (cudaStream_t)stream>>>(P, mp);
  MLG_CHECK_LAUNCH("mlg_diffpool_fwd");
  return MLG_OK;
}

extern "C" int mlg_diffpool_bwd(const float* g_out, const float* coef, const float* x, const float* adj,
                                const float* const* weights, int64_t layers, const int64_t* dims, int64_t b, float* g_x,
                                float* g_weights, void* workspace, int64_t workspace_bytes, void* stream) {
  MLG_CHECK_ARG(g_out && coef && x && adj && weights && dims && g_x && g_weights && workspace && b >= 1, "mlg_diffpool_bwd: bad arguments");
  if (int rc = check_dims(layers, dims)) return rc;
  MLG_CHECK_ARG(mlg_diffpool_supported(layers, dims), "mlg_diffpool_bwd: shared-memory limit exceeded");
  dpf::Params P;
  const int nfl = fill_params(P, layers, dims, weights, x, adj, b);
  const int64_t ctas = mlg_diffpool_ctas(b);
  MLG_CHECK_ARG(workspace_bytes >= ctas * nfl * 4, "mlg_diffpool_bwd: workspace too small");
  P.g_out = g_out;
  P.coef = coef;
  P.g_x = g_x;
  P.partial = (float*)workspace;
  dpf::MemMap mp;
  const size_t smem = (size_t)dpf::build_map(P, mp) * 4;
  cudaStream_t st = (cudaStream_t)stream;
  MLG_CUDA(cudaFuncSetAttribute(diffpool_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  diffpool_bwd_kernel<<<(unsigned)ctas, kThreadsDP, smem, st>>>(P, mp);
  MLG_CHECK_LAUNCH("mlg_diffpool_bwd");
  diffpool_grad_reduce_kernel<<<mlg_ceil_div(nfl, 256), 256, 0, st>>>(P.partial, (int)ctas, nfl, g_weights);
  MLG_CHECK_LAUNCH("mlg_diffpool_bwd(reduce)");
  return MLG_OK;
}
#endif  // MLG_HOST_EMU
