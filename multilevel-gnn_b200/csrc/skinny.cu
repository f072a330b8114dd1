// Skinny Linear forward  out[R, N] = act(x[R, K] . W[N, K]^T + bias)  for R <= 32 rows and a long reduction (K >> N),
// fp32 FMA, deterministic, sm_100a.
//
// MultilevelGNN's classification head starts with Linear(6913 -> 256) on a batch of 32 graphs
// (models/multilevel_gnn.py:104-110 of the reference): the 7 MB weight is the only real traffic, the product is a
// batched GEMV.  cuBLAS picks sgemm_largek_lds64 for it (58 us on B200, profiles/r01_launches_trainstep_v3_xtytc.csv);
// here one warp owns one output column n and a K slice, streams W[n, slice] once with 128-bit loads, and multiplies
// it against all R rows of x (served by L1: the 8 warps of a block share the slice).  Per-slice partials are summed
// in a fixed order by a second tiny kernel that also applies bias + (Leaky)ReLU.
// HBM-bound: algorithmic bytes = 4*N*K (+ 4*R*K once).
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxR = 32;

template <int R>
__global__ void __launch_bounds__(kWarps * 32) skinny_partial_kernel(const float* __restrict__ x, unsigned ld_x,
                                                                    const float* __restrict__ W, unsigned ld_w, int N,
                                                                    int K, int k_per_slice, int rows,
                                                                    float* __restrict__ partial) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int n = blockIdx.x * kWarps + warp;
  const int slice = blockIdx.y;
  const int k_lo = slice * k_per_slice, k_hi = min(K, k_lo + k_per_slice);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  if (n < N) {
    const float* wrow = W + (size_t)n * ld_w;
    for (int k = k_lo + lane; k < k_hi; k += 32) {   // scalar, coalesced: K (6913) and the row pitch are odd
      const float w = __ldg(wrow + k);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float xv = (r < rows) ? __ldg(x + (size_t)r * ld_x + k) : 0.f;
        acc[r] = fmaf(w, xv, acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = warp_sum(acc[r]);
  if (n < N && lane == 0) {
    float* p = partial + ((size_t)slice * kMaxR) * N + n;
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (r < rows) p[(size_t)r * N] = acc[r];
  }
}

__global__ void skinny_finish_kernel(const float* __restrict__ partial, int n_slices, int rows, int N,
                                     const float* __restrict__ bias, int act, float slope, float* __restrict__ out,
                                     unsigned ld_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * N) return;
  const int r = i / N, n = i % N;
  float s = 0.f;
  for (int g = 0; g < n_slices; ++g) s += partial[((size_t)g * kMaxR + r) * N + n];
  if (bias) s += __ldg(bias + n);
  if (act) s = s > 0.f ? s : s * slope;
  out[(size_t)r * ld_out + n] = s;
}

inline int pick_slices(int64_t N, int64_t K) {
  const int64_t col_blocks = (N + kWarps - 1) / kWarps;
  int64_t s = (148 * 8 + col_blocks - 1) / col_blocks;        // ~8 blocks per SM
  const int64_t max_s = (K + 255) / 256;                      // at least 256 k per slice
  if (s > max_s) s = max_s;
  return (int)(s < 1 ? 1 : s);
}

}  // namespace

extern "C" int64_t mlg_skinny_linear_workspace_bytes(int64_t N, int64_t K) {
  return (int64_t)pick_slices(N, K) * kMaxR * N * 4;
}

extern "C" int mlg_skinny_linear(const float* x, int64_t ld_x, const float* W, int64_t ld_w, const float* bias,
                                 int64_t rows, int64_t N, int64_t K, int act, float slope, float* out, int64_t ld_out,
                                 void* workspace, int64_t workspace_bytes, void* stream) {
  MLG_CHECK_ARG(x && W && out && workspace, "mlg_skinny_linear: null pointer");
  MLG_CHECK_ARG(rows >= 1 && rows <= kMaxR && N >= 1 && K >= 1 && N < (1 << 24) && K < (1ll << 31),
                "mlg_skinny_linear: need 1 <= rows <= 32 (got %lld), N, K >= 1", (long long)rows);
  MLG_CHECK_ARG(ld_x >= K && ld_w >= K && ld_out >= N, "mlg_skinny_linear: leading dimension too small");
  MLG_CHECK_ARG(workspace_bytes >= mlg_skinny_linear_workspace_bytes(N, K), "mlg_skinny_linear: workspace too small");
  const int slices = pick_slices(N, K);
  int kps = (int)((K + slices - 1) / slices);
  kps = ((kps + 31) / 32) * 32;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((unsigned)((N + kWarps - 1) / kWarps), (unsigned)slices);
  float* part = (float*)workspace;
#define MLG_SK(RR)                                                                                             \
  skinny_partial_kernel<RR><<<grid, kWarps * 32, 0, st>>>(x, (unsigned)ld_x, W, (unsigned)ld_w, (int)N, (int)K, kps, \
                                                         (int)rows, part)
  if (rows <= 8) MLG_SK(8);
  else if (rows <= 16) MLG_SK(16);
  else MLG_SK(32);
#undef MLG_SK
  MLG_CHECK_LAUNCH("mlg_skinny_linear");
  skinny_finish_kernel<<<(unsigned)mlg_ceil_div(rows * N, 256), 256, 0, st>>>(part, slices, (int)rows, (int)N, bias, act,
                                                                             slope, out, (unsigned)ld_out);
  MLG_CHECK_LAUNCH("mlg_skinny_linear(finish)");
  return MLG_OK;
}
