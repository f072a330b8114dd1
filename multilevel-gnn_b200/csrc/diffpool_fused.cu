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
// backward, dL/dx and a per-CTA partial of the 18 parameter gradients ever touch global memory -- plus, when a backward
// pass will follow, the ~90 KB per sample of forward state backward reads again (S, Z, the pooled X / A, the row norms:
// two contiguous shared-memory ranges, written / read with 128-bit moves; 52 MB at b = 576 against 0.6 ms of recomputation).
// Without a state buffer the backward kernel recomputes the forward pass of its sample.  Parameter-gradient partials are
// reduced over the CTAs in a fixed order by a second tiny kernel.
//
// Every phase is written as "parallel for over work items, then barrier" (MLG_PFOR / MLG_SYNC) without warp-level
// primitives, so the SAME source also compiles as plain host C++ (-DMLG_HOST_EMU, tests/ only) where the items of a phase
// run sequentially: the forward and backward algebra is then checked against the CPU oracle without a GPU.
#include "block_mm.cuh"

namespace dpf {

using namespace bmm;

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
  float* state;        // [b, MemMap::state_floats] or nullptr: forward writes it (when given), backward reads it (when given;
                       // nullptr: backward recomputes the forward pass of every sample)
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
  // what backward needs from forward is one contiguous range of shared memory per layer (deg .. ra): the forward kernel can
  // write it out per sample (Params::state) and the backward kernel reads it back instead of recomputing the forward pass
  int state_off[kMaxLayers], state_len[kMaxLayers];   // floats, multiples of 4
  int state_floats;                                    // per sample
};

// ---------------------------------------------------------------------------------------------------------------------
// DenseSAGEConv forward: U = ((A X) / deg) Wrel^T + X Wroot^T + b, then row-normalise.  Leaves M = A X / deg [n x c] (ld
// P4(c)), the row norms r (clamped at eps) and Y = U / r [n x o] (ld P4(o)).  X: [n x c] (ldx; x_pad: shared memory),
// A: [n x n] (lda), both A and the outputs in shared memory.
MLG_DEV void sage_fwd(const float* X, int ldx, bool x_pad, const float* A, int lda, const float* deg, int n, int c, int o,
                      const SageW& W, float* M, float* Y, float* r, bool have_M) {
  const int cp = P4(c), op = P4(o);
  if (!have_M) {
    mm(M, cp, A, lda, 1, true, X, ldx, 1, x_pad, n, c, n, 1.f, false);
    MLG_PFOR(t, n * c) M[(t / c) * cp + t % c] /= deg[t / c];
    MLG_SYNC();
  }
  mm(Y, op, M, cp, 1, true, W.rel, 1, c, false, n, o, c, 1.f, false);      // M Wrel^T: B(k, j) = Wrel[j][k]
  mm(Y, op, X, ldx, 1, x_pad, W.root, 1, c, false, n, o, c, 1.f, true);
  MLG_PFOR(i, n) {
    float ss = 0.f;
    for (int jj = 0; jj < o; ++jj) {
      MLG_ROT(j, jj, i, o);
      const float u = Y[i * op + j] + W.bias[j];
      Y[i * op + j] = u;
      ss = fmaf(u, u, ss);
    }
    const float rr = fmaxf(sqrtf(ss), kNormEps);
    r[i] = rr;
    const float inv = 1.f / rr;
    for (int jj = 0; jj < o; ++jj) {
      MLG_ROT(j, jj, i, o);
      Y[i * op + j] *= inv;
    }
  }
  MLG_SYNC();
}

MLG_DEV void row_degrees(const float* A, int lda, int n, float* deg) {
  MLG_PFOR(i, n) {
    float s = 0.f;
    for (int jj = 0; jj < n; ++jj) {
      MLG_ROT(j, jj, i, n);
      s += A[i * lda + j];
    }
    deg[i] = fmaxf(s, 1.f);
  }
  MLG_SYNC();
}

// One DiffPool layer + its after-pool DenseSAGE, forward.  X [n x c] (ldx) -> Xn [k x h] (in shared memory, the next
// layer's input) ; the layer's pooled adjacency Ap [k x k] is the next layer's A.  stats: (F, E) of this sample.
MLG_DEV void layer_fwd(const Params& P, int l, float* sm, const MemMap& mp, const float* X, int ldx, bool x_pad, float* stats) {
  const LayerDims d = P.d[l];
  const LayerMem& m = mp.L[l];
  const int n = d.n, c = d.c, k = d.k, h = d.h;
  const int np = P4(n), kp = P4(k), hp = P4(h);
  float *A = sm + m.A, *deg = sm + m.deg, *M = sm + m.M, *S = sm + m.S, *Z = sm + m.Z;
  float *T1 = sm + m.T1, *small = sm + mp.small;
  if (!x_pad) {   // global input: one copy into T1 (free until Q = A S) instead of three latency-bound passes over global
    const int cp = P4(c);
    MLG_PFOR(t, n * c) T1[(t / c) * cp + t % c] = X[t / c * ldx + t % c];
    MLG_SYNC();
    X = T1;
    ldx = cp;
    x_pad = true;
  }
  sage_fwd(X, ldx, x_pad, A, np, deg, n, c, k, P.pool[l], M, S, sm + m.rp, false);
  sage_fwd(X, ldx, x_pad, A, np, deg, n, c, h, P.embed[l], M, Z, sm + m.re, true);
  // softmax over clusters + entropy, per row
  MLG_PFOR(i, n) {
    float* s_ = S + i * kp;
    float mx = -INFINITY;
    for (int jj = 0; jj < k; ++jj) {
      MLG_ROT(j, jj, i, k);
      mx = fmaxf(mx, s_[j]);
    }
    float den = 0.f;
    for (int jj = 0; jj < k; ++jj) {
      MLG_ROT(j, jj, i, k);
      den += expf(s_[j] - mx);
    }
    const float lse = mx + logf(den);
    sm[m.lse + i] = lse;
    float ent = 0.f;
    for (int jj = 0; jj < k; ++jj) {
      MLG_ROT(j, jj, i, k);
      const float s = expf(s_[j] - lse);
      s_[j] = s;
      ent -= s * logf(s + kEntEps);
    }
    small[i] = ent;
  }
  MLG_SYNC();
  block_sum(small, n, stats + 1);
  // Xp = S^T Z ; Q = A S ; Ap = S^T Q
  mm(sm + m.Xp, hp, S, 1, kp, true, Z, hp, 1, true, k, h, n, 1.f, false);
  mm(T1, kp, A, np, 1, true, S, kp, 1, true, n, k, n, 1.f, false);
  mm(sm + m.Ap, kp, S, 1, kp, true, T1, kp, 1, true, k, k, n, 1.f, false);
  // F = ||A - S S^T||_F^2.  S S^T is symmetric: one work item per unordered pair of 4-row blocks {ti, tj} forms the 4 x 4
  // tile once (rows of S: 128-bit loads along the clusters) and charges it against A[ti, tj] and, off the diagonal,
  // A[tj, ti].  Pairs are enumerated as (ti, (ti + d) % tn), d = 0 .. tn / 2 (each unordered pair exactly once); one partial
  // per item (single writer), summed in a fixed tree.
  {
    const int tn = (n + 3) >> 2, k4 = k & ~3, nd = tn / 2 + 1;
    MLG_PFOR(t, tn * nd) {
      const int ti = t / nd, dd = t - ti * nd;
      int tj = ti + dd;
      if (tj >= tn) tj -= tn;
      float f = 0.f;
      // even tn: the d = tn / 2 pairs would appear from both ends
      if (!((tn & 1) == 0 && dd == tn / 2 && ti >= tn / 2)) {
        const int i0 = ti << 2, j0 = tj << 2;
        const float* si[4];
        const float* sj[4];
        for (int r = 0; r < 4; ++r) si[r] = S + (i0 + r < n ? i0 + r : i0) * kp;
        for (int q = 0; q < 4; ++q) sj[q] = S + (j0 + q < n ? j0 + q : j0) * kp;
        float cc[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) cc[r][q] = 0.f;
        for (int q0 = 0; q0 < k4; q0 += 4) {
          F4 a[4], bq[4];
#pragma unroll
          for (int r = 0; r < 4; ++r) a[r] = ld4(si[r] + q0);
#pragma unroll
          for (int q = 0; q < 4; ++q) bq[q] = ld4(sj[q] + q0);
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q)
              cc[r][q] = fmaf(a[r].x, bq[q].x, fmaf(a[r].y, bq[q].y, fmaf(a[r].z, bq[q].z, fmaf(a[r].w, bq[q].w, cc[r][q]))));
        }
        for (int q0 = k4; q0 < k; ++q0)
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) cc[r][q] = fmaf(si[r][q0], sj[q][q0], cc[r][q]);
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (i0 + r < n && j0 + q < n) {
              const float d1 = A[(i0 + r) * np + j0 + q] - cc[r][q];
              f = fmaf(d1, d1, f);
              if (ti != tj) {
                const float d2 = A[(j0 + q) * np + i0 + r] - cc[r][q];
                f = fmaf(d2, d2, f);
              }
            }
      }
      small[t] = f;
    }
    MLG_SYNC();
    block_sum(small, tn * nd, stats + 0);
  }
  // after-pool DenseSAGE on (Xp, Ap)
  row_degrees(sm + m.Ap, kp, k, sm + m.degp);
  sage_fwd(sm + m.Xp, hp, true, sm + m.Ap, kp, sm + m.degp, k, h, h, P.after[l], sm + m.Mp, sm + m.Xn, sm + m.ra, false);
}

// ---------------------------------------------------------------------------------------------------------------------
// backward helpers
// dU_i = (dY_i - Y_i (Y_i . dY_i)) / r_i   (r_i = max(||U_i||, eps); at the clamp the norm is a constant: dU = dY / eps)
MLG_DEV void normalize_bwd_row(float* dY, const float* Y, float r, int o, int i) {
  float dot = 0.f;
  if (r > kNormEps)
    for (int jj = 0; jj < o; ++jj) {
      MLG_ROT(j, jj, i, o);
      dot = fmaf(Y[j], dY[j], dot);
    }
  const float inv = 1.f / r;
  for (int jj = 0; jj < o; ++jj) {
    MLG_ROT(j, jj, i, o);
    dY[j] = (dY[j] - Y[j] * dot) * inv;
  }
}

// gW[o x c] (global partial, ld c) += dU^T X ; single writer per element
MLG_DEV void wgrad(float* gW, const float* dU, int ldu, int o, const float* X, int ldx, bool x_pad, int c, int n) {
  mm(gW, c, dU, 1, ldu, true, X, ldx, 1, x_pad, o, c, n, 1.f, true);
}
MLG_DEV void bgrad(float* gb, const float* dU, int ldu, int o, int n) {
  MLG_PFOR(j, o) {
    float s = 0.f;
    for (int i = 0; i < n; ++i) s += dU[i * ldu + j];
    gb[j] += s;
  }
  MLG_SYNC();
}

// DenseSAGEConv backward given dU [n x o] (ld P4(o), shared memory; gradient at the pre-normalisation output).  Accumulates
// the parameter gradients, adds dL/dX into dX [n x c] (ldd) and, when dA != nullptr, dL/dA (A is itself a function of
// earlier layers).  T: [n x c] scratch (ld P4(c)).  M = A X / deg (from forward, ld P4(c)); rowsum(A) > 1 <=> deg > 1.
MLG_DEV void sage_bwd(const float* dU, int o, const float* X, int ldx, bool x_pad, const float* A, int lda, const float* deg,
                      const float* M, int n, int c, const SageW& W, float* gWrel, float* gWroot, float* gb, float* T,
                      bool first_into_T, float* dX, int ldd, bool acc_dX, bool finish, float* dA) {
  const int cp = P4(c), op = P4(o);
  wgrad(gWrel, dU, op, o, M, cp, true, c, n);
  wgrad(gWroot, dU, op, o, X, ldx, x_pad, c, n);
  bgrad(gb, dU, op, o, n);
  // T (+)= dU Wrel  (dL/dM, summed over the convolutions that share M) ; dX (+)= dU Wroot
  mm(T, cp, dU, op, 1, true, W.rel, c, 1, false, n, c, o, 1.f, !first_into_T);
  mm(dX, ldd, dU, op, 1, true, W.root, c, 1, false, n, c, o, 1.f, acc_dX);
  if (!finish) return;
  if (dA) {   // through deg: d/d deg_i of (A X)_i / deg_i = -M_i / deg_i, only where the row sum exceeds the clamp
    MLG_PFOR(t, n * n) {
      const int i = t / n;
      if (deg[i] > 1.f) {
        float dot = 0.f;
        for (int q = 0; q < c; ++q) dot = fmaf(T[i * cp + q], M[i * cp + q], dot);   // (same row for a whole warp: broadcast)
        dA[i * lda + (t - i * n)] -= dot / deg[i];
      }
    }
    MLG_SYNC();
  }
  MLG_PFOR(t, n * c) T[(t / c) * cp + t % c] /= deg[t / c];
  MLG_SYNC();
  mm(dX, ldd, A, 1, lda, true, T, cp, 1, true, n, c, n, 1.f, true);          // A^T T
  if (dA) mm(dA, lda, T, cp, 1, true, X, 1, ldx, x_pad, n, n, c, 1.f, true);   // T X^T
}

// One DiffPool layer backward.  In: dXn [k x h] (ld P4(h)) = dL/d(after-pool output) (shared memory, overwritten), dAp_in
// [k x k] (ld P4(k)) = dL/dAp from later layers (nullptr: none; overwritten).  Out: dX [n x c] (ldd; global for layer 0) and, for l > 0,
// dA [n x n] (ld P4(n)) added.
MLG_DEV void layer_bwd(const Params& P, int l, float* sm, const MemMap& mp, const float* X, int ldx, bool x_pad, float* dXn,
                       float* dAp_in, float* dX, int ldd, float* dA, float* gpart) {
  const LayerDims d = P.d[l];
  const LayerMem& m = mp.L[l];
  const int n = d.n, c = d.c, k = d.k, h = d.h;
  const int np = P4(n), cp = P4(c), kp = P4(k), hp = P4(h);
  float *A = sm + m.A, *deg = sm + m.deg, *M = sm + m.M, *S = sm + m.S, *Z = sm + m.Z;
  float *T1 = sm + m.T1, *T2 = sm + m.T2, *small = sm + mp.small;
  const float c_link = P.coef[2 * l], c_ent = P.coef[2 * l + 1];
  const int* go = P.grad_off[l];
  // scratch inside `small`: dXp [k x hp], dAp [k x kp], G [k x kp] (later S^T S), Tp [k x hp]
  // (dAp: the caller's dAp_in is consumed in place when there is one -- only the last layer needs its own)
  float* dXp = small;
  float* G = dXp + k * hp;
  float* SS = G;
  float* Tp = G + k * kp;
  float* dAp = dAp_in ? dAp_in : Tp + k * hp;
  // ---- after-pool DenseSAGE: Xn = normalize(U_a) ----
  MLG_PFOR(i, k) normalize_bwd_row(dXn + i * hp, sm + m.Xn + i * hp, sm[m.ra + i], h, i);
  MLG_SYNC();
  if (!dAp_in) {
    MLG_PFOR(t, k * kp) dAp[t] = 0.f;
  }
  MLG_SYNC();
  sage_bwd(dXn, h, sm + m.Xp, hp, true, sm + m.Ap, kp, sm + m.degp, sm + m.Mp, k, h, P.after[l], gpart + go[6], gpart + go[7],
           gpart + go[8], Tp, true, dXp, hp, false, true, dAp);
  // ---- pooling: Xp = S^T Z, Ap = S^T A S, link, entropy ----
  // dS = Z dXp^T (into T2) ; dZ = S dXp (T1, as [n x hp]) -> dU_e (over Z)
  mm(T2, kp, Z, hp, 1, true, dXp, 1, hp, true, n, k, h, 1.f, false);
  mm(T1, hp, S, kp, 1, true, dXp, hp, 1, true, n, h, k, 1.f, false);
  MLG_PFOR(i, n) {
    normalize_bwd_row(T1 + i * hp, Z + i * hp, sm[m.re + i], h, i);
    for (int jj = 0; jj < h; ++jj) {
      MLG_ROT(j, jj, i, h);
      Z[i * hp + j] = T1[i * hp + j];      // Z now holds dU_e
    }
  }
  MLG_SYNC();
  // G1 = dAp^T - c_link I ; Q = A S (T1) ; dS += Q G1
  MLG_PFOR(t, k * k) G[(t / k) * kp + t % k] = dAp[(t % k) * kp + t / k] - ((t % k) == (t / k) ? c_link : 0.f);
  MLG_SYNC();
  mm(T1, kp, A, np, 1, true, S, kp, 1, true, n, k, n, 1.f, false);
  mm(T2, kp, T1, kp, 1, true, G, kp, 1, true, n, k, k, 1.f, true);
  if (dA) {   // dA += S dAp S^T + c_link (A - S S^T) = (S dAp - c_link S) S^T + c_link A   (this layer's A = previous Ap)
    mm(T1, kp, S, kp, 1, true, dAp, kp, 1, true, n, k, k, 1.f, false);
    MLG_PFOR(t, n * k) T1[(t / k) * kp + t % k] -= c_link * S[(t / k) * kp + t % k];
    MLG_SYNC();
    mm(dA, np, T1, kp, 1, true, S, 1, kp, true, n, n, k, 1.f, true);
    MLG_PFOR(t, n * n) dA[(t / n) * np + t % n] = fmaf(c_link, A[(t / n) * np + t % n], dA[(t / n) * np + t % n]);
    MLG_SYNC();
  }
  // G2 = dAp - c_link I ; Q2 = A^T S (T1) ; dS += Q2 G2
  MLG_PFOR(t, k * k) G[(t / k) * kp + t % k] = dAp[(t / k) * kp + t % k] - ((t % k) == (t / k) ? c_link : 0.f);
  MLG_SYNC();
  mm(T1, kp, A, 1, np, true, S, kp, 1, true, n, k, n, 1.f, false);
  mm(T2, kp, T1, kp, 1, true, G, kp, 1, true, n, k, k, 1.f, true);
  // + 2 c_link S (S^T S)
  mm(SS, kp, S, 1, kp, true, S, kp, 1, true, k, k, n, 1.f, false);
  mm(T2, kp, S, kp, 1, true, SS, kp, 1, true, n, k, k, 2.f * c_link, true);
  // entropy, softmax backward, normalisation backward: T2 row i -> dU_p_i
  MLG_PFOR(i, n) {
    float* dS = T2 + i * kp;
    const float* s = S + i * kp;
    float dot = 0.f;
    for (int jj = 0; jj < k; ++jj) {
      MLG_ROT(j, jj, i, k);
      dS[j] -= c_ent * (logf(s[j] + kEntEps) + s[j] / (s[j] + kEntEps));
      dot = fmaf(s[j], dS[j], dot);
    }
    const float lse = sm[m.lse + i], r = sm[m.rp + i];
    // dR = S (dS - S.dS) ; S_raw = log S + lse (the normalised pre-softmax row) ; dU_p = normalize_bwd(dR, S_raw, r)
    float dn = 0.f;
    for (int jj = 0; jj < k; ++jj) {
      MLG_ROT(j, jj, i, k);
      const float dr = s[j] * (dS[j] - dot);
      dS[j] = dr;
      if (r > kNormEps) dn = fmaf(logf(s[j]) + lse, dr, dn);
    }
    const float inv = 1.f / r;
    for (int jj = 0; jj < k; ++jj) {
      MLG_ROT(j, jj, i, k);
      dS[j] = (dS[j] - (logf(s[j]) + lse) * dn) * inv;
    }
  }
  MLG_SYNC();
  // ---- the two DenseSAGE convolutions that produced S and Z (shared M) ----
  if (m.recompute_M) {   // S is dead from here on: its storage takes M = A X / deg again
    M = S;
    mm(M, cp, A, np, 1, true, X, ldx, 1, x_pad, n, c, n, 1.f, false);
    MLG_PFOR(t, n * c) M[(t / c) * cp + t % c] /= deg[t / c];
    MLG_SYNC();
  }
  sage_bwd(T2, k, X, ldx, x_pad, A, np, deg, M, n, c, P.pool[l], gpart + go[0], gpart + go[1], gpart + go[2], T1, true, dX, ldd,
           false, false, nullptr);
  sage_bwd(Z, h, X, ldx, x_pad, A, np, deg, M, n, c, P.embed[l], gpart + go[3], gpart + go[4], gpart + go[5], T1, false, dX, ldd,
           true, true, dA);
}

// ---------------------------------------------------------------------------------------------------------------------
MLG_DEV void load_adj(const Params& P, float* sm, const MemMap& mp) {
  const int n = P.d[0].n, np = P4(n);
  MLG_PFOR(t, n * n) sm[mp.L[0].A + (t / n) * np + t % n] = P.adj[t];
  MLG_SYNC();
  row_degrees(sm + mp.L[0].A, np, n, sm + mp.L[0].deg);
}

MLG_DEV void sample_fwd(const Params& P, float* sm, const MemMap& mp, int s, float* stats) {
  const float* X = P.x + (size_t)s * P.d[0].n * P.d[0].c;
  int ldx = P.d[0].c;
  bool x_pad = false;
  for (int l = 0; l < P.layers; ++l) {
    if (l > 0) row_degrees(sm + mp.L[l].A, P4(P.d[l].n), P.d[l].n, sm + mp.L[l].deg);
    layer_fwd(P, l, sm, mp, X, ldx, x_pad, stats + 2 * l);
    X = sm + mp.L[l].Xn;
    ldx = P4(P.d[l].h);
    x_pad = true;
  }
}

// copy the per-layer state ranges of the current sample to / from global memory (128-bit moves)
MLG_DEV void state_io(const Params& P, float* sm, const MemMap& mp, int s, bool save) {
  float* g = P.state + (size_t)s * mp.state_floats;
  for (int l = 0; l < P.layers; ++l) {
    F4* a = reinterpret_cast<F4*>(sm + mp.state_off[l]);
    F4* d = reinterpret_cast<F4*>(g);
    const int n4 = mp.state_len[l] >> 2;
    if (save) {
      MLG_PFOR(t, n4) d[t] = a[t];
    } else {
      MLG_PFOR(t, n4) a[t] = d[t];
    }
    g += mp.state_len[l];
  }
  MLG_SYNC();
}

MLG_DEV void forward_body(const Params& P, float* sm, const MemMap& mp, int cta, int nctas) {
  load_adj(P, sm, mp);
  const LayerDims dl = P.d[P.layers - 1];
  const int hp = P4(dl.h);
  for (int s = cta; s < P.b; s += nctas) {
    sample_fwd(P, sm, mp, s, P.stats + (size_t)s * 2 * P.layers);
    if (P.state) state_io(P, sm, mp, s, true);
    float* o = P.out + (size_t)s * dl.k * dl.h;
    const float* xn = sm + mp.L[P.layers - 1].Xn;
    MLG_PFOR(t, dl.k * dl.h) o[t] = xn[(t / dl.h) * hp + t % dl.h];
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
  const int hpl = P4(dl.h);
  float* small = sm + mp.small;
  // tail of `small` (see build_map): forward statistics (not needed again), dL/d(out) [k_last x hp], and for two layers
  // dL/dXn0 [k0 x hp0] + dL/dAp0 [k0 x kp0]
  float* st = small + mp.small_floats - 2 * kMaxLayers;
  float* dXn = st - dl.k * hpl;
  for (int s = cta; s < P.b; s += nctas) {
    if (P.state) state_io(P, sm, mp, s, false);
    else sample_fwd(P, sm, mp, s, st);
    const float* go = P.g_out + (size_t)s * dl.k * dl.h;
    MLG_PFOR(t, dl.k * dl.h) dXn[(t / dl.h) * hpl + t % dl.h] = go[t];
    MLG_SYNC();
    if (Lz == 1) {
      layer_bwd(P, 0, sm, mp, P.x + (size_t)s * P.d[0].n * P.d[0].c, P.d[0].c, false, dXn, nullptr,
                P.g_x + (size_t)s * P.d[0].n * P.d[0].c, P.d[0].c, nullptr, gpart);
    } else {
      // layer 1: its input is layer 0's after-pool output Xn0 [k0 x h0], its adjacency layer 0's Ap0 [k0 x k0]
      const LayerDims d0 = P.d[0];
      const int hp0 = P4(d0.h), kp0 = P4(d0.k);
      float* dX1 = dXn - d0.k * hp0;     // dL/dXn0
      float* dA1 = dX1 - d0.k * kp0;     // dL/dAp0
      MLG_PFOR(t, d0.k * kp0) dA1[t] = 0.f;
      MLG_SYNC();
      layer_bwd(P, 1, sm, mp, sm + mp.L[0].Xn, hp0, true, dXn, nullptr, dX1, hp0, dA1, gpart);
      layer_bwd(P, 0, sm, mp, P.x + (size_t)s * d0.n * d0.c, d0.c, false, dX1, dA1, P.g_x + (size_t)s * d0.n * d0.c, d0.c,
                nullptr, gpart);
    }
  }
}

// host side: shared-memory map (float offsets, every matrix with P4-padded rows at a 16-byte aligned offset).  Layer l > 0
// aliases its adjacency onto layer l-1's pooled adjacency.
static inline int build_map(const Params& P, MemMap& mp) {
  int off = 0;
  auto take = [&](int nfl) { const int o = off; off += (nfl + 3) & ~3; return o; };
  auto p4 = [](int v) { return (v + 3) & ~3; };
  auto tsize = [&](const LayerDims& d) { const int kk = d.k > d.h ? d.k : d.h; return d.n * p4(kk > d.c ? kk : d.c); };
  auto state = [&](int l, bool own_M) {
    const LayerDims d = P.d[l];
    LayerMem& m = mp.L[l];
    mp.state_off[l] = off;
    m.deg = take(d.n);
    if (own_M) m.M = take(d.n * p4(d.c));
    m.S = take(d.n * p4(own_M ? d.k : (d.k > d.c ? d.k : d.c)));
    m.Z = take(d.n * p4(d.h));
    m.rp = take(d.n);
    m.re = take(d.n);
    m.lse = take(d.n);
    m.Xp = take(d.k * p4(d.h));
    m.Ap = take(d.k * p4(d.k));
    m.degp = take(d.k);
    m.Mp = take(d.k * p4(d.h));
    m.Xn = take(d.k * p4(d.h));
    m.ra = take(d.k);
    mp.state_len[l] = off - mp.state_off[l];
  };
  int smax = 0;
  mp.state_floats = 0;
  for (int l = 0; l < kMaxLayers; ++l) mp.state_off[l] = mp.state_len[l] = 0;
  for (int l = 0; l < P.layers; ++l) {
    const LayerDims d = P.d[l];
    // per-row scratch (n) ; layer_bwd's dXp, Tp, G and -- last layer only -- dAp
    const int sl = p4(d.n) + 2 * d.k * p4(d.h) + (l == P.layers - 1 ? 2 : 1) * d.k * p4(d.k);
    smax = sl > smax ? sl : smax;
    // block_sum over the F partials (one per pair of 4-row blocks) and its tree levels
    const int tn = (d.n + 3) / 4, items = tn * (tn / 2 + 1), sf = items + items / 8 + 64;
    smax = sf > smax ? sf : smax;
  }
  // layer 0
  mp.L[0].A = take(P.d[0].n * p4(P.d[0].n));
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
  // tail of `small`: statistics, dL/d(out), and for two layers dL/dXn0 + dL/dAp0 (all P4-padded rows)
  const LayerDims dl = P.d[P.layers - 1];
  int tail = 2 * kMaxLayers + dl.k * p4(dl.h);
  if (P.layers > 1) tail += P.d[0].k * p4(P.d[0].h) + P.d[0].k * p4(P.d[0].k);
  smax = (smax + 3) & ~3;
  mp.small_floats = smax + ((tail + 3) & ~3) + 4;
  mp.small = take(mp.small_floats);
  mp.total = off;
  for (int l = 0; l < P.layers; ++l) mp.state_floats += mp.state_len[l];
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

extern "C" int64_t mlg_diffpool_state_floats(int64_t layers, const int64_t* dims) {
  if (!dims || layers < 1 || layers > dpf::kMaxLayers) return -1;
  dpf::Params P;
  memset(&P, 0, sizeof(P));
  P.layers = (int)layers;
  for (int l = 0; l < layers; ++l) P.d[l] = {(int)dims[4 * l], (int)dims[4 * l + 1], (int)dims[4 * l + 2], (int)dims[4 * l + 3]};
  dpf::MemMap mp;
  dpf::build_map(P, mp);
  return mp.state_floats;
}

extern "C" int64_t mlg_diffpool_ctas(int64_t b) { return b < 148 ? (b < 1 ? 1 : b) : 148; }

extern "C" int mlg_diffpool_fwd(const float* x, const float* adj, const float* const* weights, int64_t layers,
                                const int64_t* dims, int64_t b, float* out, float* stats, float* state, void* stream) {
  MLG_CHECK_ARG(x && adj && weights && dims && out && stats && b >= 1, "mlg_diffpool_fwd: bad arguments");
  if (int rc = check_dims(layers, dims)) return rc;
  MLG_CHECK_ARG(mlg_diffpool_supported(layers, dims), "mlg_diffpool_fwd: %lld bytes of shared memory needed (limit %d)",
                (long long)mlg_diffpool_smem_bytes(layers, dims), kMaxSmemBytes);
  dpf::Params P;
  fill_params(P, layers, dims, weights, x, adj, b);
  P.out = out;
  P.stats = stats;
  P.state = state;
  MLG_CHECK_ARG(((uintptr_t)state & 15) == 0, "mlg_diffpool_fwd: state must be 16-byte aligned");
  dpf::MemMap mp;
  const size_t smem = (size_t)dpf::build_map(P, mp) * 4;
  MLG_CUDA(cudaFuncSetAttribute(diffpool_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  diffpool_fwd_kernel<<<(unsigned)mlg_diffpool_ctas(b), kThreadsDP, smem, (cudaStream_t)stream>>>(P, mp);
  MLG_CHECK_LAUNCH("mlg_diffpool_fwd");
  return MLG_OK;
}

extern "C" int mlg_diffpool_bwd(const float* g_out, const float* coef, const float* x, const float* adj,
                                const float* const* weights, int64_t layers, const int64_t* dims, int64_t b, float* g_x,
                                float* g_weights, const float* state, void* workspace, int64_t workspace_bytes, void* stream) {
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
  P.state = const_cast<float*>(state);
  MLG_CHECK_ARG(((uintptr_t)state & 15) == 0, "mlg_diffpool_bwd: state must be 16-byte aligned");
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
