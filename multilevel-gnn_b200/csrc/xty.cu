// Tall-skinny transposed product  out[M,K] = A[rows,M]^T * X[rows,K]  (+ column sums of A), fp32, sm_100a.
//
// This is the weight/bias gradient of every Linear on the hot path (SAGEConv.update's MLP and lin_r,
// models/gcn_lib/sparse/torch_vertex.py:281-291): rows = B*N = 492,960 nodes, M,K <= 128.  A library
// GEMM sees an [M x rows] x [rows x K] problem with a tiny output and spends ~385 us per call on the
// gbm shape (profiles/r01_launches_trainstep_v0.csv); here the row range is split over the grid, every
// block keeps a 64x128 fp32 accumulator tile in registers (4x8 per thread), rows are staged through a
// 2-stage cp.async ring in shared memory, and the per-block partials are reduced in a fixed order
// (deterministic, no atomics).  fp32 FMA only: the parity bar is rtol 1e-4 on gradients.
// Compute-bound on the fp32 pipe: flops = 2*rows*M*K; compulsory bytes = 4*rows*(M+K).
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int TM = 64, TK = 128, TR = 32, kThreads = 256, kStages = 2;
constexpr int kSmemBytes = kStages * TR * (TM + TK) * 4;

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem, int src_bytes) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool ALIGNED>
__global__ void __launch_bounds__(kThreads, 2)
xty_kernel(const float* __restrict__ A, unsigned ld_a, const float* __restrict__ X, unsigned ld_x, long long rows,
           int M, int K, int rows_per_block, float* __restrict__ partial, int want_colsum,
           float* __restrict__ colsum_direct) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                          // [kStages][TR][TM]
  float* Xs = smem + kStages * TR * TM;      // [kStages][TR][TK]
  const int tid = threadIdx.x;
  const int tiles_k = (K + TK - 1) / TK;
  const int m0 = (blockIdx.y / tiles_k) * TM, k0 = (blockIdx.y % tiles_k) * TK;
  const long long r_begin = (long long)blockIdx.x * rows_per_block;
  const long long r_end = min(rows, r_begin + rows_per_block);
  const int n_steps = (int)((r_end - r_begin + TR - 1) / TR);

  auto load_stage = [&](int step, int stage) {
    const long long r0 = r_begin + (long long)step * TR;
    if (ALIGNED) {
      // A tile: TR x TM floats = 512 float4, X tile: TR x TK = 1024 float4
#pragma unroll
      for (int i = 0; i < (TR * TM / 4) / kThreads; ++i) {
        const int f = tid + i * kThreads;
        const int r = f / (TM / 4), c4 = f % (TM / 4);
        const long long gr = r0 + r;
        const int gm = m0 + c4 * 4;
        const bool ok = gr < r_end && gm < M;
        cp_async16(As + (stage * TR + r) * TM + c4 * 4, ok ? (const void*)(A + (size_t)gr * ld_a + gm) : (const void*)A,
                   ok ? 16 : 0);
      }
#pragma unroll
      for (int i = 0; i < (TR * TK / 4) / kThreads; ++i) {
        const int f = tid + i * kThreads;
        const int r = f / (TK / 4), c4 = f % (TK / 4);
        const long long gr = r0 + r;
        const int gk = k0 + c4 * 4;
        const bool ok = gr < r_end && gk < K;
        cp_async16(Xs + (stage * TR + r) * TK + c4 * 4, ok ? (const void*)(X + (size_t)gr * ld_x + gk) : (const void*)X,
                   ok ? 16 : 0);
      }
    } else {
      // unaligned / odd shapes: 4-byte copies
      for (int f = tid; f < TR * TM; f += kThreads) {
        const int r = f / TM, cc = f % TM;
        const long long gr = r0 + r;
        const bool ok = gr < r_end && m0 + cc < M;
        cp_async4(As + (stage * TR + r) * TM + cc, ok ? (const void*)(A + (size_t)gr * ld_a + m0 + cc) : (const void*)A,
                  ok ? 4 : 0);
      }
      for (int f = tid; f < TR * TK; f += kThreads) {
        const int r = f / TK, cc = f % TK;
        const long long gr = r0 + r;
        const bool ok = gr < r_end && k0 + cc < K;
        cp_async4(Xs + (stage * TR + r) * TK + cc, ok ? (const void*)(X + (size_t)gr * ld_x + k0 + cc) : (const void*)X,
                  ok ? 4 : 0);
      }
    }
    cp_async_commit();
  };

  const int tm = tid / 16, tk = tid % 16;  // 16 x 16 threads; thread tile: m = tm*4..+3, k = {tk*4..+3, 64+tk*4..+3}
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};

  if (n_steps > 0) load_stage(0, 0);
  for (int s = 0; s < n_steps; ++s) {
    const int stage = s & 1;
    if (s + 1 < n_steps) {
      load_stage(s + 1, stage ^ 1);
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* as = As + stage * TR * TM + tm * 4;
    const float* xs = Xs + stage * TR * TK + tk * 4;
#pragma unroll 8
    for (int r = 0; r < TR; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(as + r * TM);
      const float4 x0 = *reinterpret_cast<const float4*>(xs + r * TK);
      const float4 x1 = *reinterpret_cast<const float4*>(xs + r * TK + 64);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], xv[j], acc[i][j]);
        cs[i] += av[i];
      }
    }
    __syncthreads();
  }
  // partial[blockIdx.x][m][k] (+ a trailing [M] block of column sums per row-chunk)
  // (with a single row chunk the host passes `out` as partial and `colsum` as colsum_direct: no reduce pass)
  const size_t stride = (size_t)M * K + (want_colsum ? M : 0);
  float* pb = partial + (size_t)blockIdx.x * stride;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + tm * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + h * 64 + tk * 4 + j;
        if (k < K) pb[(size_t)m * K + k] = acc[i][h * 4 + j];
      }
    }
    if (want_colsum && tk == 0 && k0 == 0) (colsum_direct ? colsum_direct : pb + (size_t)M * K)[m] = cs[i];
  }
}

// 8 lanes per output element sum interleaved partials, then a fixed-order shuffle tree
__global__ void xty_reduce_kernel(const float* __restrict__ partial, int n_part, long long stride, long long mk,
                                  int M, float* __restrict__ out, float* __restrict__ colsum) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long i = t >> 3;
  const int sub = (int)(t & 7);
  float s = 0.f;
  if (i < stride)
    for (int g = sub; g < n_part; g += 8) s += partial[(size_t)g * stride + i];
  s += __shfl_down_sync(0xffffffffu, s, 4, 8);
  s += __shfl_down_sync(0xffffffffu, s, 2, 8);
  s += __shfl_down_sync(0xffffffffu, s, 1, 8);
  if (sub != 0 || i >= stride) return;
  if (i < mk) out[i] = s;
  else if (colsum) colsum[i - mk] = s;
}

inline int pick_parts(long long rows, int tiles) {
  long long want = (2 * 148 + tiles - 1) / tiles;            // ~2 blocks per SM overall
  long long max_parts = (rows + 4 * TR - 1) / (4 * TR);      // at least 4 row-steps per block
  long long p = want < max_parts ? want : max_parts;
  return (int)(p < 1 ? 1 : p);
}

}  // namespace

extern "C" int64_t mlg_xty_workspace_bytes(int64_t rows, int64_t M, int64_t K) {
  const int tiles = (int)(((M + TM - 1) / TM) * ((K + TK - 1) / TK));
  return (int64_t)pick_parts(rows, tiles) * (M * K + M) * 4;
}

extern "C" int mlg_xty(const float* A, int64_t ld_a, const float* X, int64_t ld_x, int64_t rows, int64_t M,
                       int64_t K, float* out, float* colsum, void* workspace, int64_t workspace_bytes,
                       void* stream) {
  MLG_CHECK_ARG(A && X && out && workspace, "mlg_xty: null pointer");
  MLG_CHECK_ARG(rows >= 0 && M > 0 && K > 0 && M <= (1 << 20) && K <= (1 << 20) && M * K < (1ll << 31), "mlg_xty: bad sizes");
  const bool aligned = ld_a % 4 == 0 && ld_x % 4 == 0 && M % 4 == 0 && K % 4 == 0 && (uintptr_t)A % 16 == 0 &&
                       (uintptr_t)X % 16 == 0;
  MLG_CHECK_ARG(workspace_bytes >= mlg_xty_workspace_bytes(rows, M, K), "mlg_xty: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int tiles = (int)(((M + TM - 1) / TM) * ((K + TK - 1) / TK));
  const int parts = pick_parts(rows, tiles);
  long long rpb = (rows + parts - 1) / parts;
  rpb = ((rpb + TR - 1) / TR) * TR;
  if (rpb < TR) rpb = TR;
  dim3 grid(parts, tiles);
  float* dst = parts == 1 ? out : (float*)workspace;
  float* cs_direct = parts == 1 ? colsum : nullptr;
  static bool attr_done = false;   // 48 KB is exactly the default limit; set once, outside any stream capture
  if (!attr_done) {
    MLG_CUDA(cudaFuncSetAttribute(xty_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    MLG_CUDA(cudaFuncSetAttribute(xty_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    attr_done = true;
  }
  if (aligned) {
    xty_kernel<true><<<grid, kThreads, kSmemBytes, st>>>(A, (unsigned)ld_a, X, (unsigned)ld_x, rows, (int)M, (int)K,
                                                        (int)rpb, dst, colsum ? 1 : 0, cs_direct);
  } else {
    xty_kernel<false><<<grid, kThreads, kSmemBytes, st>>>(A, (unsigned)ld_a, X, (unsigned)ld_x, rows, (int)M, (int)K,
                                                         (int)rpb, dst, colsum ? 1 : 0, cs_direct);
  }
  MLG_CHECK_LAUNCH("mlg_xty");
  const long long stride = M * K + (colsum ? M : 0);
  if (parts == 1) return MLG_OK;   // the kernel wrote out / colsum directly
  xty_reduce_kernel<<<mlg_ceil_div(stride * 8, 256), 256, 0, st>>>((const float*)workspace, parts, stride, M * K, (int)M,
                                                              out, colsum);
  MLG_CHECK_LAUNCH("mlg_xty(reduce)");
  return MLG_OK;
}
