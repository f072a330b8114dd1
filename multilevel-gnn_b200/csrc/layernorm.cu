// LayerNorm over the channel axis of a tall [rows, C] activation, forward + backward (sm_100a).
//
// DeeperGCN applies nn.LayerNorm to every node row twice per layer (the res+ block's norm, models/deepergcn.py:262-275,
// and the norm inside GENConv's MLP, gcn_lib/sparse/torch_nn.py:55-73 via norm_layer('layer', ...)), rows = 100 146 nodes,
// C = 128 / 256.  The library kernels take 160 us (forward), 81 us (input gradient) and 268 us (gamma/beta gradient) per
// call at that shape against a ~35 / 50 us HBM floor (profiles/r01_launches_deepergcn4.csv).  Here one warp owns one row
// (C/32 values per lane in registers, two-pass mean / variance in registers), and the backward produces the input
// gradient AND the per-block partial gamma / beta gradients in the same pass; partials are summed in fixed order.
//   y = (x - mean) * rstd * gamma + beta,   rstd = 1 / sqrt(var_biased + eps)
//   gx = rstd * (dy - mean_c(dy) - xhat * mean_c(dy * xhat)),   dy = g * gamma
// HBM-bound: forward 8*rows*C bytes, backward 12*rows*C.
// RELU variants (mlg_layernorm_relu_fwd / _bwd): the ReLU that follows both norms (deepergcn.py:268-270 `F.relu(h1)`,
// torch_nn.py MLP: Lin -> norm -> act) is applied in the same pass; backward rebuilds the mask from
// xhat * gamma + beta > 0 (= output > 0, ATen's rule) instead of reading the activation: two elementwise passes over
// [rows, C] per norm less (a ReLU forward and a threshold backward, 17-45 us each at 100 k rows).
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int kWarps = 8;

template <int NV, bool RELU>   // C == 128 * NV: lane holds NV float4 at columns v*128 + lane*4
__global__ void __launch_bounds__(kWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta, long long rows,
              float eps, float* __restrict__ y, float* __restrict__ mean_out, float* __restrict__ rstd_out) {
  constexpr int C = 128 * NV;
  const int lane = threadIdx.x & 31;
  const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * C + lane * 4;
  float4 v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = ld_stream4(xr + i * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) * (1.f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / C) + eps);
  float* yr = y + (size_t)row * C + lane * 4;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float4 g = gamma ? ld_gather4(gamma + i * 128 + lane * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
    const float4 b = beta ? ld_gather4(beta + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 o = make_float4(fmaf((v[i].x - mean) * rstd, g.x, b.x), fmaf((v[i].y - mean) * rstd, g.y, b.y),
                           fmaf((v[i].z - mean) * rstd, g.z, b.z), fmaf((v[i].w - mean) * rstd, g.w, b.w));
    if (RELU) o = make_float4(fmaxf(o.x, 0.f), fmaxf(o.y, 0.f), fmaxf(o.z, 0.f), fmaxf(o.w, 0.f));
    st4(yr + i * 128, o);
  }
  if (lane == 0) {
    mean_out[row] = mean;
    rstd_out[row] = rstd;
  }
}

template <int NV, bool RELU>
__global__ void __launch_bounds__(kWarps * 32)
ln_bwd_kernel(const float* __restrict__ x, const float* __restrict__ g, const float* __restrict__ gamma,
              const float* __restrict__ beta, const float* __restrict__ mean_in, const float* __restrict__ rstd_in, long long rows,
              float* __restrict__ gx, float* __restrict__ partial /* [gridDim.x][2][C] */) {
  constexpr int C = 128 * NV;
  extern __shared__ float stage[];   // [kWarps][2][C]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long warp0 = (long long)blockIdx.x * kWarps + wib;
  const long long stride = (long long)gridDim.x * kWarps;
  float4 gam[NV], bet[RELU ? NV : 1], dg[NV], db[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    gam[i] = gamma ? ld_gather4(gamma + i * 128 + lane * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
    if (RELU) bet[i] = beta ? ld_gather4(beta + i * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    dg[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    db[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (long long row = warp0; row < rows; row += stride) {
    const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
    const float* xr = x + (size_t)row * C + lane * 4;
    const float* gr = g + (size_t)row * C + lane * 4;
    float4 xh[NV], dy[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xh[i] = ld_stream4(xr + i * 128);
      dy[i] = ld_stream4(gr + i * 128);
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      xh[i] = make_float4((xh[i].x - mean) * rstd, (xh[i].y - mean) * rstd, (xh[i].z - mean) * rstd, (xh[i].w - mean) * rstd);
      if (RELU) {   // gradient reaches the norm only where its output was positive (the forward's own expression)
        if (!(fmaf(xh[i].x, gam[i].x, bet[i].x) > 0.f)) dy[i].x = 0.f;
        if (!(fmaf(xh[i].y, gam[i].y, bet[i].y) > 0.f)) dy[i].y = 0.f;
        if (!(fmaf(xh[i].z, gam[i].z, bet[i].z) > 0.f)) dy[i].z = 0.f;
        if (!(fmaf(xh[i].w, gam[i].w, bet[i].w) > 0.f)) dy[i].w = 0.f;
      }
      dg[i].x = fmaf(dy[i].x, xh[i].x, dg[i].x); dg[i].y = fmaf(dy[i].y, xh[i].y, dg[i].y);
      dg[i].z = fmaf(dy[i].z, xh[i].z, dg[i].z); dg[i].w = fmaf(dy[i].w, xh[i].w, dg[i].w);
      db[i].x += dy[i].x; db[i].y += dy[i].y; db[i].z += dy[i].z; db[i].w += dy[i].w;
      dy[i] = make_float4(dy[i].x * gam[i].x, dy[i].y * gam[i].y, dy[i].z * gam[i].z, dy[i].w * gam[i].w);
      s1 += (dy[i].x + dy[i].y) + (dy[i].z + dy[i].w);
      s2 += (dy[i].x * xh[i].x + dy[i].y * xh[i].y) + (dy[i].z * xh[i].z + dy[i].w * xh[i].w);
    }
    const float c1 = warp_sum(s1) * (1.f / C), c2 = warp_sum(s2) * (1.f / C);
    float* o = gx + (size_t)row * C + lane * 4;
#pragma unroll
    for (int i = 0; i < NV; ++i)
      st4(o + i * 128, make_float4((dy[i].x - c1 - xh[i].x * c2) * rstd, (dy[i].y - c1 - xh[i].y * c2) * rstd,
                                   (dy[i].z - c1 - xh[i].z * c2) * rstd, (dy[i].w - c1 - xh[i].w * c2) * rstd));
  }
  // block partials: warps write their column sums to shared memory, warp-order sum (deterministic)
  float* mine = stage + (size_t)wib * 2 * C;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    *reinterpret_cast<float4*>(mine + i * 128 + lane * 4) = dg[i];
    *reinterpret_cast<float4*>(mine + C + i * 128 + lane * 4) = db[i];
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * C; j += kWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += stage[(size_t)w * 2 * C + j];
    partial[(size_t)blockIdx.x * 2 * C + j] = s;
  }
}

// one block per 32 columns of the [2C] partial rows: lane = column, the 8 warps split the partials (coalesced 128-byte
// reads, 4 in flight), then a warp-ordered sum through shared memory (fixed order: deterministic)
__global__ void __launch_bounds__(256)
ln_param_reduce_kernel(const float* __restrict__ partial, int n_part, int C2, float* __restrict__ dgamma,
                       float* __restrict__ dbeta) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (j < C2) {
    int p = w;
    for (; p + 24 < n_part; p += 32) {
      s0 += partial[(size_t)p * C2 + j];
      s1 += partial[(size_t)(p + 8) * C2 + j];
      s2 += partial[(size_t)(p + 16) * C2 + j];
      s3 += partial[(size_t)(p + 24) * C2 + j];
    }
    for (; p < n_part; p += 8) s0 += partial[(size_t)p * C2 + j];
  }
  sm[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && j < C2) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][lane];
    const int C = C2 / 2;
    if (j < C) { if (dgamma) dgamma[j] = s; }
    else if (dbeta) dbeta[j - C] = s;
  }
}

inline int bwd_blocks(long long rows) {
  long long b = (rows + kWarps - 1) / kWarps;
  const long long cap = 148 * 4;
  return (int)(b < cap ? b : cap);
}

}  // namespace

extern "C" int mlg_layernorm_supported(int64_t C) { return C == 128 || C == 256 || C == 384 || C == 512; }

extern "C" int64_t mlg_layernorm_bwd_workspace_bytes(int64_t rows, int64_t C) {
  return (int64_t)bwd_blocks(rows) * 2 * C * 4;
}

static int layernorm_fwd_impl(const float* x, const float* gamma, const float* beta, int64_t rows, int64_t C, float eps,
                              float* y, float* mean, float* rstd, bool relu, void* stream) {
  MLG_CHECK_ARG(x && y && mean && rstd, "mlg_layernorm_fwd: null pointer");
  MLG_CHECK_ARG(mlg_layernorm_supported(C) && rows >= 0, "mlg_layernorm_fwd: C=%lld not supported (128/256/384/512)", (long long)C);
  MLG_CHECK_ARG(((uintptr_t)x | (uintptr_t)y) % 16 == 0 && (!gamma || (uintptr_t)gamma % 16 == 0) &&
                    (!beta || (uintptr_t)beta % 16 == 0), "mlg_layernorm_fwd: operands must be 16-byte aligned");
  if (rows == 0) return MLG_OK;
  const unsigned grid = (unsigned)mlg_ceil_div(rows, kWarps);
  cudaStream_t st = (cudaStream_t)stream;
#define MLG_LN_F(NV)                                                                                         \
  if (relu) ln_fwd_kernel<NV, true><<<grid, kWarps * 32, 0, st>>>(x, gamma, beta, rows, eps, y, mean, rstd); \
  else ln_fwd_kernel<NV, false><<<grid, kWarps * 32, 0, st>>>(x, gamma, beta, rows, eps, y, mean, rstd);
  switch (C / 128) {
    case 1: MLG_LN_F(1) break;
    case 2: MLG_LN_F(2) break;
    case 3: MLG_LN_F(3) break;
    default: MLG_LN_F(4) break;
  }
#undef MLG_LN_F
  MLG_CHECK_LAUNCH("mlg_layernorm_fwd");
  return MLG_OK;
}

static int layernorm_bwd_impl(const float* x, const float* g, const float* gamma, const float* beta, const float* mean,
                              const float* rstd, int64_t rows, int64_t C, float* gx, float* dgamma, float* dbeta,
                              void* workspace, int64_t workspace_bytes, bool relu, void* stream) {
  MLG_CHECK_ARG(x && g && mean && rstd && gx && workspace, "mlg_layernorm_bwd: null pointer");
  MLG_CHECK_ARG(mlg_layernorm_supported(C) && rows >= 0, "mlg_layernorm_bwd: C=%lld not supported (128/256/384/512)", (long long)C);
  MLG_CHECK_ARG(((uintptr_t)x | (uintptr_t)g | (uintptr_t)gx) % 16 == 0 && (!gamma || (uintptr_t)gamma % 16 == 0) &&
                    (!beta || (uintptr_t)beta % 16 == 0),
                "mlg_layernorm_bwd: operands must be 16-byte aligned");
  MLG_CHECK_ARG(workspace_bytes >= mlg_layernorm_bwd_workspace_bytes(rows, C), "mlg_layernorm_bwd: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  if (rows == 0) {
    if (dgamma) MLG_CUDA(cudaMemsetAsync(dgamma, 0, C * 4, st));
    if (dbeta) MLG_CUDA(cudaMemsetAsync(dbeta, 0, C * 4, st));
    return MLG_OK;
  }
  const int grid = bwd_blocks(rows);
  const size_t smem = (size_t)kWarps * 2 * C * 4;
  float* part = (float*)workspace;
#define MLG_LN_B(NV)                                                                                                 \
  if (relu) ln_bwd_kernel<NV, true><<<grid, kWarps * 32, smem, st>>>(x, g, gamma, beta, mean, rstd, rows, gx, part); \
  else ln_bwd_kernel<NV, false><<<grid, kWarps * 32, smem, st>>>(x, g, gamma, beta, mean, rstd, rows, gx, part);
  switch (C / 128) {
    case 1: MLG_LN_B(1) break;
    case 2: MLG_LN_B(2) break;
    case 3: MLG_LN_B(3) break;
    default: MLG_LN_B(4) break;
  }
#undef MLG_LN_B
  MLG_CHECK_LAUNCH("mlg_layernorm_bwd");
  if (dgamma || dbeta) {
    ln_param_reduce_kernel<<<(unsigned)mlg_ceil_div(2 * C, 32), 256, 0, st>>>(part, grid, (int)(2 * C), dgamma, dbeta);
    MLG_CHECK_LAUNCH("mlg_layernorm_bwd(reduce)");
  }
  return MLG_OK;
}

extern "C" int mlg_layernorm_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int64_t C, float eps,
                                 float* y, float* mean, float* rstd, void* stream) {
  return layernorm_fwd_impl(x, gamma, beta, rows, C, eps, y, mean, rstd, false, stream);
}
extern "C" int mlg_layernorm_bwd(const float* x, const float* g, const float* gamma, const float* mean, const float* rstd,
                                 int64_t rows, int64_t C, float* gx, float* dgamma, float* dbeta, void* workspace,
                                 int64_t workspace_bytes, void* stream) {
  return layernorm_bwd_impl(x, g, gamma, nullptr, mean, rstd, rows, C, gx, dgamma, dbeta, workspace, workspace_bytes, false, stream);
}
extern "C" int mlg_layernorm_relu_fwd(const float* x, const float* gamma, const float* beta, int64_t rows, int64_t C,
                                      float eps, float* y, float* mean, float* rstd, void* stream) {
  return layernorm_fwd_impl(x, gamma, beta, rows, C, eps, y, mean, rstd, true, stream);
}
extern "C" int mlg_layernorm_relu_bwd(const float* x, const float* g, const float* gamma, const float* beta,
                                      const float* mean, const float* rstd, int64_t rows, int64_t C, float* gx,
                                      float* dgamma, float* dbeta, void* workspace, int64_t workspace_bytes, void* stream) {
  return layernorm_bwd_impl(x, g, gamma, beta, mean, rstd, rows, C, gx, dgamma, dbeta, workspace, workspace_bytes, true, stream);
}
