// MultilevelGNN's classification head and loss as FOUR kernels (two forward, two backward), sm_100a.
//
// Reference (models/multilevel_gnn.py:262-290, train.py:60,118): on the pooled pathway tensor [B, 32, 146, 3P]
//     Conv2d(32->32, 1x1) + ReLU, Conv2d(32->64, 1x1) + ReLU, MaxPool2d((kh, kw)), Dropout(0.25), flatten, cat(age),
//     Linear(F+1 -> D) + ReLU + Dropout(0.5), Linear(D -> 2), Softmax, BCELoss(weight [B,2]).
// As library ops that is ~60 launches per training step (each under 12 us: conv-as-GEMM x2, bias/ReLU, max-pool, two
// dropouts, cat, two Linears, softmax, BCE and all their backward ops) -- a fifth of the gbm step.  Here:
//   head_conv_pool_fwd   one warp per pooling window: both 1x1 convs of every pixel of the window as 32-lane mat-vecs
//                        (weights in shared memory, activations exchanged by shuffles), running max, dropout, and the
//                        result scattered straight into the flattened [B, F+1] matrix (age in the last column).
//   head_mlp_fwd         split-K batched GEMV over the [D, F+1] weight (the only real traffic: 7 MB for gbm), then one
//                        block per sample: bias + ReLU + dropout, Linear(D -> 2), softmax, weighted BCE, mean.
//   head_mlp_bwd         one pass over the [D, F+1] index space produces BOTH the weight gradient (written once) and
//                        the input gradient (the weight read once); the tiny softmax/BCE/Linear(D->2) backward is
//                        recomputed by every block.
//   head_conv_pool_bwd   recomputes the two convs per window (cheaper than saving 3 activations), routes the gradient to
//                        the first maximum (ATen's tie rule), back through both convs to the pooled features, and
//                        accumulates the four parameter gradients in registers; fixed-order two-level reduction.
// Dropout masks come from caller-provided random int32 words (torch's generator: graph-safe, torch.manual_seed applies);
// element kept iff bits >= p * 2^31.  Everything is fp32 FMA with fixed summation orders (deterministic).
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int CIN = 32, C1 = 32, C2 = 64;
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kGradFloats = C1 * CIN + C1 + C2 * C1 + C2;   // gW1, gb1, gW2, gb2 = 3168

__device__ __forceinline__ bool drop_keep(const int* bits, size_t i, int thr) { return bits == nullptr || __ldg(bits + i) >= thr; }

struct HeadW {   // shared-memory copies: *t = transposed, so that lane = output channel reads are conflict-free
  float w1[C1 * CIN];    // [j][i]
  float w1t[CIN * C1];   // [i][j]
  float w2[C2 * C1];     // [c][i]
  float w2t[C1 * C2];    // [i][c]
  float b1[C1];
  float b2[C2];
};

__device__ __forceinline__ void load_weights(HeadW& S, const float* W1, const float* b1, const float* W2, const float* b2) {
  for (int t = threadIdx.x; t < C1 * CIN; t += blockDim.x) {
    const float v = __ldg(W1 + t);
    S.w1[t] = v;
    S.w1t[(t % CIN) * C1 + t / CIN] = v;
  }
  for (int t = threadIdx.x; t < C2 * C1; t += blockDim.x) {
    const float v = __ldg(W2 + t);
    S.w2[t] = v;
    S.w2t[(t % C1) * C2 + t / C1] = v;
  }
  for (int t = threadIdx.x; t < C1; t += blockDim.x) S.b1[t] = __ldg(b1 + t);
  for (int t = threadIdx.x; t < C2; t += blockDim.x) S.b2[t] = __ldg(b2 + t);
  __syncthreads();
}

// both 1x1 convs of one pixel: lane holds x[lane] -> h1[lane] (post-ReLU), z2[lane], z2[lane+32] (pre-ReLU)
__device__ __forceinline__ void pixel_forward(const HeadW& S, float x, int lane, float& h1, float& za, float& zb) {
  float a = S.b1[lane];
#pragma unroll
  for (int i = 0; i < CIN; ++i) a = fmaf(S.w1t[i * C1 + lane], __shfl_sync(0xffffffffu, x, i), a);
  h1 = fmaxf(a, 0.f);
  za = S.b2[lane];
  zb = S.b2[lane + 32];
#pragma unroll
  for (int i = 0; i < C1; ++i) {
    const float hv = __shfl_sync(0xffffffffu, h1, i);
    za = fmaf(S.w2t[i * C2 + lane], hv, za);
    zb = fmaf(S.w2t[i * C2 + lane + 32], hv, zb);
  }
}

struct ConvPoolP {
  const float* x;   // [B, H, W, CIN] channel-last
  const float *W1, *b1, *W2, *b2;
  const float* age;   // [B] or NULL
  const int* bits;    // [B * F] or NULL
  int thr;            // keep iff bits >= thr
  float keep_scale;   // 1 / (1 - p)
  int B, H, W, kh, kw, Ho, Wo;
  long long ld;       // row pitch of the flattened matrix
  float* a0;          // fwd out [B, ld]
  // backward
  const float* g_a0;  // [B, ld]
  float* g_x;         // [B, H, W, CIN]
  float* partial;     // [blocks, kGradFloats]
};

__global__ void __launch_bounds__(kThreads) head_conv_pool_fwd_kernel(const ConvPoolP P) {
  __shared__ HeadW S;
  load_weights(S, P.W1, P.b1, P.W2, P.b2);
  const int lane = threadIdx.x & 31;
  const long long nwin = (long long)P.B * P.Ho * P.Wo;
  const long long HoWo = (long long)P.Ho * P.Wo;
  const long long F = C2 * HoWo;
  for (long long o = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5; o < nwin; o += (long long)gridDim.x * kWarps) {
    const int wo = (int)(o % P.Wo);
    const int ho = (int)((o / P.Wo) % P.Ho);
    const int b = (int)(o / HoWo);
    float best_a = -INFINITY, best_b = -INFINITY;
    for (int dh = 0; dh < P.kh; ++dh)
      for (int dw = 0; dw < P.kw; ++dw) {
        const size_t pix = ((size_t)b * P.H + (size_t)ho * P.kh + dh) * P.W + (size_t)wo * P.kw + dw;
        const float x = __ldg(P.x + pix * CIN + lane);
        float h1, za, zb;
        pixel_forward(S, x, lane, h1, za, zb);
        best_a = fmaxf(best_a, fmaxf(za, 0.f));
        best_b = fmaxf(best_b, fmaxf(zb, 0.f));
      }
    const size_t col_a = (size_t)lane * HoWo + (size_t)ho * P.Wo + wo, col_b = col_a + (size_t)32 * HoWo;
    float* row = P.a0 + (size_t)b * P.ld;
    row[col_a] = drop_keep(P.bits, (size_t)b * F + col_a, P.thr) ? best_a * P.keep_scale : 0.f;
    row[col_b] = drop_keep(P.bits, (size_t)b * F + col_b, P.thr) ? best_b * P.keep_scale : 0.f;
    if (P.age && ho == 0 && wo == 0 && lane == 0) row[F] = __ldg(P.age + b);
  }
}

__global__ void __launch_bounds__(kThreads) head_conv_pool_bwd_kernel(const ConvPoolP P) {
  __shared__ HeadW S;
  __shared__ float red[kGradFloats];
  load_weights(S, P.W1, P.b1, P.W2, P.b2);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long HoWo = (long long)P.Ho * P.Wo;
  const long long F = C2 * HoWo;
  const long long nwin = (long long)P.B * HoWo;
  // pixels the floor-mode pool never reads (tail rows / columns) get a zero gradient: extra work items after the windows
  const int tail_rows = P.H - P.Ho * P.kh, tail_cols = P.W - P.Wo * P.kw;
  const long long tail_per_b = (long long)tail_rows * P.W + (long long)(P.H - tail_rows) * tail_cols;
  const long long nitems = nwin + (long long)P.B * tail_per_b;
  float gw1[CIN], gw2a[C1], gw2b[C1], gb1 = 0.f, gb2a = 0.f, gb2b = 0.f;
#pragma unroll
  for (int i = 0; i < CIN; ++i) gw1[i] = 0.f;
#pragma unroll
  for (int i = 0; i < C1; ++i) gw2a[i] = gw2b[i] = 0.f;
  for (long long o = ((long long)blockIdx.x * kThreads + threadIdx.x) >> 5; o < nitems; o += (long long)gridDim.x * kWarps) {
    if (o >= nwin) {   // a dropped pixel
      const long long t = o - nwin;
      const int b = (int)(t / tail_per_b);
      long long r = t % tail_per_b;
      int h, w;
      if (r < (long long)tail_rows * P.W) {
        h = P.Ho * P.kh + (int)(r / P.W);
        w = (int)(r % P.W);
      } else {
        r -= (long long)tail_rows * P.W;
        h = (int)(r / tail_cols);
        w = P.Wo * P.kw + (int)(r % tail_cols);
      }
      P.g_x[(((size_t)b * P.H + h) * P.W + w) * CIN + lane] = 0.f;
      continue;
    }
    const int wo = (int)(o % P.Wo);
    const int ho = (int)((o / P.Wo) % P.Ho);
    const int b = (int)(o / HoWo);
    // pass 1: window maximum and its FIRST position (ATen: strict >) per channel
    float best_a = -INFINITY, best_b = -INFINITY;
    int arg_a = 0, arg_b = 0;
    for (int dh = 0; dh < P.kh; ++dh)
      for (int dw = 0; dw < P.kw; ++dw) {
        const size_t pix = ((size_t)b * P.H + (size_t)ho * P.kh + dh) * P.W + (size_t)wo * P.kw + dw;
        const float x = __ldg(P.x + pix * CIN + lane);
        float h1, za, zb;
        pixel_forward(S, x, lane, h1, za, zb);
        const float ra = fmaxf(za, 0.f), rb = fmaxf(zb, 0.f);
        if (ra > best_a) { best_a = ra; arg_a = dh * P.kw + dw; }
        if (rb > best_b) { best_b = rb; arg_b = dh * P.kw + dw; }
      }
    const size_t col_a = (size_t)lane * HoWo + (size_t)ho * P.Wo + wo, col_b = col_a + (size_t)32 * HoWo;
    const float* grow = P.g_a0 + (size_t)b * P.ld;
    // gradient reaching the maximum: dropout scale, and ReLU'(z) = 0 where the maximum is the clamped 0
    float go_a = drop_keep(P.bits, (size_t)b * F + col_a, P.thr) ? __ldg(grow + col_a) * P.keep_scale : 0.f;
    float go_b = drop_keep(P.bits, (size_t)b * F + col_b, P.thr) ? __ldg(grow + col_b) * P.keep_scale : 0.f;
    if (!(best_a > 0.f)) go_a = 0.f;
    if (!(best_b > 0.f)) go_b = 0.f;
    // pass 2: per pixel, back through conv2 / ReLU / conv1
    for (int dh = 0; dh < P.kh; ++dh)
      for (int dw = 0; dw < P.kw; ++dw) {
        const int pos = dh * P.kw + dw;
        const size_t pix = ((size_t)b * P.H + (size_t)ho * P.kh + dh) * P.W + (size_t)wo * P.kw + dw;
        const float gza = arg_a == pos ? go_a : 0.f, gzb = arg_b == pos ? go_b : 0.f;
        float gx = 0.f;
        if (__any_sync(0xffffffffu, gza != 0.f || gzb != 0.f)) {
          const float x = __ldg(P.x + pix * CIN + lane);
          float h1, za, zb;
          pixel_forward(S, x, lane, h1, za, zb);
          gb2a += gza;
          gb2b += gzb;
          float gh = 0.f;   // dL/dh1[lane]
#pragma unroll
          for (int i = 0; i < C1; ++i) {
            const float hv = __shfl_sync(0xffffffffu, h1, i);
            gw2a[i] = fmaf(gza, hv, gw2a[i]);
            gw2b[i] = fmaf(gzb, hv, gw2b[i]);
            gh = fmaf(S.w2[i * C1 + lane], __shfl_sync(0xffffffffu, gza, i), gh);
            gh = fmaf(S.w2[(i + 32) * C1 + lane], __shfl_sync(0xffffffffu, gzb, i), gh);
          }
          const float gz1 = h1 > 0.f ? gh : 0.f;
          gb1 += gz1;
#pragma unroll
          for (int i = 0; i < CIN; ++i) {
            gw1[i] = fmaf(gz1, __shfl_sync(0xffffffffu, x, i), gw1[i]);
            gx = fmaf(S.w1[i * CIN + lane], __shfl_sync(0xffffffffu, gz1, i), gx);
          }
        }
        P.g_x[pix * CIN + lane] = gx;
      }
  }
  // block partial: warps add their register accumulators in turn (fixed order), then one row of the workspace
  for (int t = threadIdx.x; t < kGradFloats; t += kThreads) red[t] = 0.f;
  __syncthreads();
  for (int w = 0; w < kWarps; ++w) {
    if (warp == w) {
#pragma unroll
      for (int i = 0; i < CIN; ++i) red[lane * CIN + i] += gw1[i];
      red[C1 * CIN + lane] += gb1;
      float* r2 = red + C1 * CIN + C1;
#pragma unroll
      for (int i = 0; i < C1; ++i) {
        r2[lane * C1 + i] += gw2a[i];
        r2[(lane + 32) * C1 + i] += gw2b[i];
      }
      r2[C2 * C1 + lane] += gb2a;
      r2[C2 * C1 + lane + 32] += gb2b;
    }
    __syncthreads();
  }
  for (int t = threadIdx.x; t < kGradFloats; t += kThreads) P.partial[(size_t)blockIdx.x * kGradFloats + t] = red[t];
}

__global__ void head_conv_reduce_kernel(const float* __restrict__ partial, int nblocks, float* __restrict__ gW1,
                                        float* __restrict__ gb1, float* __restrict__ gW2, float* __restrict__ gb2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= kGradFloats) return;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += partial[(size_t)b * kGradFloats + t];
  if (t < C1 * CIN) gW1[t] = s;
  else if (t < C1 * CIN + C1) gb1[t - C1 * CIN] = s;
  else if (t < C1 * CIN + C1 + C2 * C1) gW2[t - C1 * CIN - C1] = s;
  else gb2[t - C1 * CIN - C1 - C2 * C1] = s;
}

inline int conv_pool_blocks(long long items) {
  long long b = (items + kWarps - 1) / kWarps;
  const long long cap = 148 * 4;
  return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------------------------------------
// MLP + loss
// ------------------------------------------------------------------------------------------------------------
struct MlpP {
  const float* a0;   // [R, ld_a] (K columns used)
  const float* W0;   // [D, K] row pitch K
  const float* b0;   // [D]
  const float* W3;   // [2, D]
  const float* b3;   // [2]
  const int* bits;   // [R * D] or NULL
  int thr;
  float keep_scale;
  const float* y;        // [R, 2] or NULL (no loss)
  const float* weight;   // [R, 2] or NULL
  int R, D, K, slices, kps;
  long long ld_a;
  float* partial;   // [slices, R, D]
  float* a1;        // [R, D] post ReLU + dropout
  float* pred;      // [R, 2]
  float* rowloss;   // [R]
  float* loss;      // [1]
  unsigned* counter;
};

template <int RT>
__global__ void __launch_bounds__(kThreads) head_mlp_partial_kernel(const MlpP P) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * kWarps + warp;
  const int k_lo = blockIdx.y * P.kps, k_hi = min(P.K, k_lo + P.kps);
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) *P.counter = 0u;   // the finish kernel's arrival counter
  float acc[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = 0.f;
  if (n < P.D) {
    const float* wrow = P.W0 + (size_t)n * P.K;
    for (int k = k_lo + lane; k < k_hi; k += 32) {
      const float w = __ldg(wrow + k);
#pragma unroll
      for (int r = 0; r < RT; ++r) {
        const float xv = (r < P.R) ? __ldg(P.a0 + (size_t)r * P.ld_a + k) : 0.f;
        acc[r] = fmaf(w, xv, acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RT; ++r) acc[r] = warp_sum(acc[r]);
  if (n < P.D && lane == 0) {
    float* p = P.partial + ((size_t)blockIdx.y * P.R) * P.D + n;
#pragma unroll
    for (int r = 0; r < RT; ++r)
      if (r < P.R) p[(size_t)r * P.D] = acc[r];
  }
}

// one block per sample r: slice reduction + bias + ReLU + dropout -> a1[r, :]; logits, softmax, weighted BCE; the last
// block to finish sums the per-sample losses in index order
__global__ void __launch_bounds__(kThreads) head_mlp_finish_kernel(const MlpP P) {
  __shared__ float red[2][kWarps];
  __shared__ int is_last;
  const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float z0 = 0.f, z1 = 0.f;
  for (int n = threadIdx.x; n < P.D; n += kThreads) {
    float s = 0.f;
    for (int g = 0; g < P.slices; ++g) s += P.partial[((size_t)g * P.R + r) * P.D + n];
    s += __ldg(P.b0 + n);
    s = fmaxf(s, 0.f);
    s = drop_keep(P.bits, (size_t)r * P.D + n, P.thr) ? s * P.keep_scale : 0.f;
    P.a1[(size_t)r * P.D + n] = s;
    z0 = fmaf(s, __ldg(P.W3 + n), z0);
    z1 = fmaf(s, __ldg(P.W3 + P.D + n), z1);
  }
  z0 = warp_sum(z0);
  z1 = warp_sum(z1);
  if (lane == 0) { red[0][warp] = z0; red[1][warp] = z1; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = __ldg(P.b3), b = __ldg(P.b3 + 1);
    for (int w = 0; w < kWarps; ++w) { a += red[0][w]; b += red[1][w]; }
    const float m = fmaxf(a, b);
    const float ea = expf(a - m), eb = expf(b - m);
    const float inv = 1.f / (ea + eb);
    const float p0 = ea * inv, p1 = eb * inv;
    P.pred[2 * r] = p0;
    P.pred[2 * r + 1] = p1;
    if (P.y) {
      // BCELoss: -(y log p + (1 - y) log(1 - p)), logs clamped at -100 (ATen), times weight
      const float y0 = __ldg(P.y + 2 * r), y1 = __ldg(P.y + 2 * r + 1);
      const float w0 = P.weight ? __ldg(P.weight + 2 * r) : 1.f, w1 = P.weight ? __ldg(P.weight + 2 * r + 1) : 1.f;
      const float l0 = -(y0 * fmaxf(logf(p0), -100.f) + (1.f - y0) * fmaxf(logf(1.f - p0), -100.f)) * w0;
      const float l1 = -(y1 * fmaxf(logf(p1), -100.f) + (1.f - y1) * fmaxf(logf(1.f - p1), -100.f)) * w1;
      P.rowloss[r] = l0 + l1;
      __threadfence();
      const unsigned c = atomicAdd(P.counter, 1u);
      is_last = (c == gridDim.x - 1);
    } else {
      is_last = 0;
    }
  }
  __syncthreads();
  if (is_last && threadIdx.x == 0) {
    __threadfence();
    float s = 0.f;
    for (int i = 0; i < P.R; ++i) s += *reinterpret_cast<volatile float*>(P.rowloss + i);
    P.loss[0] = s / (2.f * (float)P.R);
    *P.counter = 0u;
  }
}

struct MlpBwdP {
  const float* g_pred;   // [R, 2] or NULL
  const float* g_loss;   // [1] or NULL
  const float *pred, *y, *weight;
  const float *a0, *a1, *W0, *W3;
  int R, D, K;
  long long ld_a, ld_g;
  float keep_scale;
  float *g_a0, *g_W0, *g_b0, *g_W3, *g_b3;
};

// Block = one 32-column chunk of K; warp w = output rows n in [32w, 32w + 32) of that chunk (D / 32 warps).
// shared: gz1 [R][D] (dL/dz1, recomputed by every block); the same storage then holds the per-warp partials
// [warps][R][32] of the input gradient (warps * 32 == D).
template <int RT>
__global__ void __launch_bounds__(512) head_mlp_bwd_kernel(const MlpBwdP P) {
  extern __shared__ float sm[];
  float* gz1 = sm;                                  // [R][D]
  float* accs = sm;                                 // [nwarps][R][32], after the main loop
  float* gz2 = sm + (size_t)P.R * P.D;              // [R][2]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  // ---- phase A: softmax + BCE backward, Linear(D -> 2) backward, ReLU / dropout mask ----
  for (int r = threadIdx.x; r < P.R; r += blockDim.x) {
    const float p0 = __ldg(P.pred + 2 * r), p1 = __ldg(P.pred + 2 * r + 1);
    float g0 = P.g_pred ? __ldg(P.g_pred + 2 * r) : 0.f, g1 = P.g_pred ? __ldg(P.g_pred + 2 * r + 1) : 0.f;
    if (P.g_loss && P.y) {
      const float gl = __ldg(P.g_loss) / (2.f * (float)P.R);
      const float y0 = __ldg(P.y + 2 * r), y1 = __ldg(P.y + 2 * r + 1);
      const float w0 = P.weight ? __ldg(P.weight + 2 * r) : 1.f, w1 = P.weight ? __ldg(P.weight + 2 * r + 1) : 1.f;
      // ATen binary_cross_entropy_backward: (p - y) / max((1 - p) p, 1e-12) * weight
      g0 += gl * w0 * (p0 - y0) / fmaxf((1.f - p0) * p0, 1e-12f);
      g1 += gl * w1 * (p1 - y1) / fmaxf((1.f - p1) * p1, 1e-12f);
    }
    const float dot = g0 * p0 + g1 * p1;
    gz2[2 * r] = p0 * (g0 - dot);
    gz2[2 * r + 1] = p1 * (g1 - dot);
  }
  __syncthreads();
  for (int t = threadIdx.x; t < P.R * P.D; t += blockDim.x) {
    const int r = t / P.D, n = t % P.D;
    const float ga1 = gz2[2 * r] * __ldg(P.W3 + n) + gz2[2 * r + 1] * __ldg(P.W3 + P.D + n);
    gz1[t] = __ldg(P.a1 + t) > 0.f ? ga1 * P.keep_scale : 0.f;
  }
  __syncthreads();
  if (blockIdx.x == 0) {   // the small parameter gradients, fixed order
    for (int n = threadIdx.x; n < P.D; n += blockDim.x) {
      float s = 0.f, u0 = 0.f, u1 = 0.f;
      for (int r = 0; r < P.R; ++r) {
        s += gz1[(size_t)r * P.D + n];
        const float a = __ldg(P.a1 + (size_t)r * P.D + n);
        u0 = fmaf(gz2[2 * r], a, u0);
        u1 = fmaf(gz2[2 * r + 1], a, u1);
      }
      P.g_b0[n] = s;
      P.g_W3[n] = u0;
      P.g_W3[P.D + n] = u1;
    }
    if (threadIdx.x < 2) {
      float s = 0.f;
      for (int r = 0; r < P.R; ++r) s += gz2[2 * r + threadIdx.x];
      P.g_b3[threadIdx.x] = s;
    }
  }
  // ---- phase B: this block's 32 columns of K ----
  const int k = blockIdx.x * 32 + lane;
  const bool kok = k < P.K;
  float a0r[RT], accA[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    a0r[r] = (kok && r < P.R) ? __ldg(P.a0 + (size_t)r * P.ld_a + k) : 0.f;
    accA[r] = 0.f;
  }
#pragma unroll 4
  for (int n = warp * 32; n < warp * 32 + 32; ++n) {
    const float w = kok ? __ldg(P.W0 + (size_t)n * P.K + k) : 0.f;
    float accW = 0.f;
#pragma unroll
    for (int r = 0; r < RT; ++r) {
      const float g = r < P.R ? gz1[(size_t)r * P.D + n] : 0.f;
      accA[r] = fmaf(g, w, accA[r]);
      accW = fmaf(g, a0r[r], accW);
    }
    if (kok) P.g_W0[(size_t)n * P.K + k] = accW;
  }
  __syncthreads();   // every warp is done with gz1: its storage becomes the partials buffer
#pragma unroll
  for (int r = 0; r < RT; ++r)
    if (r < P.R) accs[((size_t)warp * P.R + r) * 32 + lane] = accA[r];
  __syncthreads();
  if (P.g_a0) {
    for (int t = threadIdx.x; t < P.R * 32; t += blockDim.x) {
      const int r = t >> 5, l = t & 31;
      const int kk = blockIdx.x * 32 + l;
      if (kk < P.K) {
        float s = 0.f;
        for (int w = 0; w < nwarps; ++w) s += accs[((size_t)w * P.R + r) * 32 + l];
        P.g_a0[(size_t)r * P.ld_g + kk] = s;
      }
    }
  }
}

inline int mlp_slices(int64_t D, int64_t K) {
  const int64_t col_blocks = (D + kWarps - 1) / kWarps;
  int64_t s = (148 * 8 + col_blocks - 1) / col_blocks;
  const int64_t max_s = (K + 255) / 256;
  if (s > max_s) s = max_s;
  return (int)(s < 1 ? 1 : s);
}

inline int drop_threshold(float p) {
  if (!(p > 0.f)) return 0;
  double t = (double)p * 2147483648.0;
  if (t > 2147483647.0) t = 2147483647.0;
  return (int)t;
}

}  // namespace

extern "C" int mlg_head_conv_pool_supported(int64_t cin, int64_t c1, int64_t c2) { return cin == CIN && c1 == C1 && c2 == C2; }

extern "C" int64_t mlg_head_conv_pool_bwd_workspace_bytes(int64_t B, int64_t H, int64_t W, int64_t kh, int64_t kw) {
  if (kh < 1 || kw < 1) return 0;
  return (int64_t)conv_pool_blocks(B * H * W) * kGradFloats * 4;
}

extern "C" int mlg_head_conv_pool_fwd(const float* x_cl, const float* W1, const float* b1, const float* W2, const float* b2,
                                      const float* age, const int32_t* drop_bits, float drop_p, int64_t B, int64_t H,
                                      int64_t W, int64_t kh, int64_t kw, float* a0, int64_t ld, void* stream) {
  MLG_CHECK_ARG(x_cl && W1 && b1 && W2 && b2 && a0, "mlg_head_conv_pool_fwd: null pointer");
  MLG_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && kh >= 1 && kw >= 1 && kh <= H && kw <= W, "mlg_head_conv_pool_fwd: bad sizes");
  MLG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f, "mlg_head_conv_pool_fwd: dropout p must be in [0, 1)");
  ConvPoolP P;
  memset(&P, 0, sizeof(P));
  P.x = x_cl; P.W1 = W1; P.b1 = b1; P.W2 = W2; P.b2 = b2; P.age = age;
  P.bits = drop_p > 0.f ? drop_bits : nullptr;
  MLG_CHECK_ARG(drop_p == 0.f || drop_bits, "mlg_head_conv_pool_fwd: dropout needs drop_bits");
  P.thr = drop_threshold(drop_p); P.keep_scale = 1.f / (1.f - drop_p);
  P.B = (int)B; P.H = (int)H; P.W = (int)W; P.kh = (int)kh; P.kw = (int)kw; P.Ho = (int)(H / kh); P.Wo = (int)(W / kw);
  const long long F = (long long)C2 * P.Ho * P.Wo;
  MLG_CHECK_ARG(ld >= F + (age ? 1 : 0), "mlg_head_conv_pool_fwd: ld too small");
  P.ld = ld; P.a0 = a0;
  head_conv_pool_fwd_kernel<<<conv_pool_blocks((long long)B * P.Ho * P.Wo), kThreads, 0, (cudaStream_t)stream>>>(P);
  MLG_CHECK_LAUNCH("mlg_head_conv_pool_fwd");
  return MLG_OK;
}

extern "C" int mlg_head_conv_pool_bwd(const float* g_a0, int64_t ld, const float* x_cl, const float* W1, const float* b1,
                                      const float* W2, const float* b2, const int32_t* drop_bits, float drop_p, int64_t B,
                                      int64_t H, int64_t W, int64_t kh, int64_t kw, float* g_x_cl, float* g_W1, float* g_b1,
                                      float* g_W2, float* g_b2, void* workspace, int64_t workspace_bytes, void* stream) {
  MLG_CHECK_ARG(g_a0 && x_cl && W1 && b1 && W2 && b2 && g_x_cl && g_W1 && g_b1 && g_W2 && g_b2 && workspace,
                "mlg_head_conv_pool_bwd: null pointer");
  MLG_CHECK_ARG(B >= 1 && H >= 1 && W >= 1 && kh >= 1 && kw >= 1 && kh <= H && kw <= W, "mlg_head_conv_pool_bwd: bad sizes");
  MLG_CHECK_ARG(workspace_bytes >= mlg_head_conv_pool_bwd_workspace_bytes(B, H, W, kh, kw), "mlg_head_conv_pool_bwd: workspace too small");
  MLG_CHECK_ARG(drop_p == 0.f || drop_bits, "mlg_head_conv_pool_bwd: dropout needs drop_bits");
  ConvPoolP P;
  memset(&P, 0, sizeof(P));
  P.x = x_cl; P.W1 = W1; P.b1 = b1; P.W2 = W2; P.b2 = b2;
  P.bits = drop_p > 0.f ? drop_bits : nullptr;
  P.thr = drop_threshold(drop_p); P.keep_scale = 1.f / (1.f - drop_p);
  P.B = (int)B; P.H = (int)H; P.W = (int)W; P.kh = (int)kh; P.kw = (int)kw; P.Ho = (int)(H / kh); P.Wo = (int)(W / kw);
  P.ld = ld; P.g_a0 = g_a0; P.g_x = g_x_cl; P.partial = (float*)workspace;
  const int blocks = conv_pool_blocks(B * H * W);
  head_conv_pool_bwd_kernel<<<blocks, kThreads, 0, (cudaStream_t)stream>>>(P);
  MLG_CHECK_LAUNCH("mlg_head_conv_pool_bwd");
  head_conv_reduce_kernel<<<mlg_ceil_div(kGradFloats, 128), 128, 0, (cudaStream_t)stream>>>(P.partial, blocks, g_W1, g_b1, g_W2, g_b2);
  MLG_CHECK_LAUNCH("mlg_head_conv_pool_bwd(reduce)");
  return MLG_OK;
}

extern "C" int64_t mlg_head_mlp_workspace_bytes(int64_t R, int64_t D, int64_t K) {
  return ((int64_t)mlp_slices(D, K) * R * D + R + 4) * 4;
}

extern "C" int mlg_head_mlp_fwd(const float* a0, int64_t ld_a, const float* W0, const float* b0, const float* W3,
                                const float* b3, const int32_t* drop_bits, float drop_p, const float* y, const float* weight,
                                int64_t R, int64_t D, int64_t K, float* a1, float* pred, float* loss, void* workspace,
                                int64_t workspace_bytes, void* stream) {
  MLG_CHECK_ARG(a0 && W0 && b0 && W3 && b3 && a1 && pred && workspace, "mlg_head_mlp_fwd: null pointer");
  MLG_CHECK_ARG(R >= 1 && R <= 64 && D >= 1 && K >= 1 && ld_a >= K, "mlg_head_mlp_fwd: need 1 <= rows <= 64 (got %lld)", (long long)R);
  MLG_CHECK_ARG(!y || loss, "mlg_head_mlp_fwd: y given without a loss output");
  MLG_CHECK_ARG(drop_p >= 0.f && drop_p < 1.f && (drop_p == 0.f || drop_bits), "mlg_head_mlp_fwd: bad dropout arguments");
  MLG_CHECK_ARG(workspace_bytes >= mlg_head_mlp_workspace_bytes(R, D, K), "mlg_head_mlp_fwd: workspace too small");
  MlpP P;
  memset(&P, 0, sizeof(P));
  P.a0 = a0; P.W0 = W0; P.b0 = b0; P.W3 = W3; P.b3 = b3;
  P.bits = drop_p > 0.f ? drop_bits : nullptr;
  P.thr = drop_threshold(drop_p); P.keep_scale = 1.f / (1.f - drop_p);
  P.y = y; P.weight = weight;
  P.R = (int)R; P.D = (int)D; P.K = (int)K; P.ld_a = ld_a;
  P.slices = mlp_slices(D, K);
  int kps = (int)((K + P.slices - 1) / P.slices);
  P.kps = ((kps + 31) / 32) * 32;
  P.partial = (float*)workspace;
  P.rowloss = P.partial + (size_t)P.slices * R * D;
  P.counter = (unsigned*)(P.rowloss + R);
  P.a1 = a1; P.pred = pred; P.loss = loss;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((D + kWarps - 1) / kWarps), (unsigned)P.slices);
  if (R <= 8) head_mlp_partial_kernel<8><<<grid, kThreads, 0, st>>>(P);
  else if (R <= 16) head_mlp_partial_kernel<16><<<grid, kThreads, 0, st>>>(P);
  else if (R <= 32) head_mlp_partial_kernel<32><<<grid, kThreads, 0, st>>>(P);
  else head_mlp_partial_kernel<64><<<grid, kThreads, 0, st>>>(P);
  MLG_CHECK_LAUNCH("mlg_head_mlp_fwd(partial)");
  head_mlp_finish_kernel<<<(unsigned)R, kThreads, 0, st>>>(P);
  MLG_CHECK_LAUNCH("mlg_head_mlp_fwd(finish)");
  return MLG_OK;
}

extern "C" int mlg_head_mlp_bwd(const float* g_pred, const float* g_loss, const float* pred, const float* y,
                                const float* weight, const float* a0, int64_t ld_a, const float* a1, const float* W0,
                                const float* W3, float drop_p, int64_t R, int64_t D, int64_t K, float* g_a0, int64_t ld_g,
                                float* g_W0, float* g_b0, float* g_W3, float* g_b3, void* stream) {
  MLG_CHECK_ARG(pred && a0 && a1 && W0 && W3 && g_W0 && g_b0 && g_W3 && g_b3, "mlg_head_mlp_bwd: null pointer");
  MLG_CHECK_ARG(g_pred || g_loss, "mlg_head_mlp_bwd: neither g_pred nor g_loss given");
  MLG_CHECK_ARG(!g_loss || y, "mlg_head_mlp_bwd: g_loss needs y");
  MLG_CHECK_ARG(R >= 1 && R <= 64 && D >= 32 && D % 32 == 0 && D <= 512 && K >= 1, "mlg_head_mlp_bwd: need rows <= 64, D a multiple of 32 up to 512");
  MLG_CHECK_ARG(ld_a >= K && (!g_a0 || ld_g >= K), "mlg_head_mlp_bwd: leading dimension too small");
  MlpBwdP P;
  memset(&P, 0, sizeof(P));
  P.g_pred = g_pred; P.g_loss = g_loss; P.pred = pred; P.y = y; P.weight = weight;
  P.a0 = a0; P.a1 = a1; P.W0 = W0; P.W3 = W3;
  P.R = (int)R; P.D = (int)D; P.K = (int)K; P.ld_a = ld_a; P.ld_g = ld_g;
  P.keep_scale = 1.f / (1.f - drop_p);
  P.g_a0 = g_a0; P.g_W0 = g_W0; P.g_b0 = g_b0; P.g_W3 = g_W3; P.g_b3 = g_b3;
  const int nwarps = (int)(D / 32);
  const size_t smem = ((size_t)R * D + 2 * R) * 4;
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned blocks = (unsigned)((K + 31) / 32);
  if (R <= 32) {
    MLG_CUDA(cudaFuncSetAttribute(head_mlp_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_mlp_bwd_kernel<32><<<blocks, nwarps * 32, smem, st>>>(P);
  } else {
    MLG_CUDA(cudaFuncSetAttribute(head_mlp_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    head_mlp_bwd_kernel<64><<<blocks, nwarps * 32, smem, st>>>(P);
  }
  MLG_CHECK_LAUNCH("mlg_head_mlp_bwd");
  return MLG_OK;
}
