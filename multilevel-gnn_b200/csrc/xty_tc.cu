// Tall-skinny transposed product on the tensor cores, fp32-accurate (3xTF32 split), sm_100a:
//
//   out[N, 128] = G[rows, N]^T . X[rows, 128]     (+ colsum[N] = sum_r G[r, :])
//
// The weight / bias gradient of the Linears whose input is 128 wide: the SAGE update GEMM over [x | agg_x]
// (models/gcn_lib/sparse/torch_vertex.py:281-291) and GENConv's per-layer edge encoder (torch_vertex.py:76-77).
// The reduction runs over the rows, i.e. both operands arrive "MN-major" (row-major [rows, cols]); instead of relying
// on MN-major tensor-core descriptors the splitter warpgroup TRANSPOSES each 32-row chunk while it splits it:
// TMA drops the raw row-major chunk into shared memory, thread t owns column t, reads its 32 values (conflict-free
// column reads), and writes hi / lo as K-major rows of 32 fp32 with the 128-byte swizzle applied by hand
// (chunk index ^ (row & 7)), which is exactly the canonical layout tcgen05.mma consumes.  D[128 x N] = X^T.G stays
// in tensor memory across ALL chunks of a CTA (persistent, strided chunk ownership); per-CTA partials are reduced in a
// fixed order (deterministic).  Same accuracy contract as mlg_gemm_tf32x3 (relative error ~2^-20).
#include <cuda.h>

#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int KX = 128;                 // columns of X = rows (M) of the accumulator
constexpr int CH = 32;                  // rows per chunk = K of one stage (one 128 B swizzle row of fp32)
constexpr int UK = 8;
constexpr int kThreads = 512;           // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue, 8-11 / 12-15 splitters

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MLGX_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MLGX_DONE_%=;\n"
      "bra MLGX_WAIT_%=;\n"
      "MLGX_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_tf32(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major tile, 128-byte swizzle, 8-row groups 1024 B apart
__device__ __forceinline__ unsigned long long make_smem_desc(unsigned addr) {
  unsigned long long d = 0;
  d |= (unsigned long long)((addr & 0x3FFFF) >> 4);
  d |= (unsigned long long)1 << 16;
  d |= (unsigned long long)(1024 >> 4) << 32;
  d |= (unsigned long long)1 << 46;
  d |= (unsigned long long)2 << 61;
  return d;
}
// kind::tf32: D = f32 (1 << 4), A = B = tf32 (format code 2), K-major, M = 128, N
__device__ __forceinline__ float lds32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(unsigned addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__host__ __device__ constexpr unsigned make_idesc_tf32_xty(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(KX >> 4) << 24);
}

constexpr int kMaxRaw = 6, kSplit = 2;

struct CtlX {
  unsigned long long full_raw[kMaxRaw], empty_raw[kMaxRaw], full_split[kSplit], empty_split[kSplit], tmem_full;
  unsigned tmem_base;
  float csum[2][KX];
  float csum_x[2][KX];
};

// shared memory: raw ring   n_raw x [X chunk 32 x 128 fp32 = 16 KB][G chunk 32 x N fp32]      (TMA -> splitter)
//                split ring kSplit x [X_hi 16 KB][X_lo 16 KB][G_hi N*128 B][G_lo N*128 B]      (splitter -> MMA)
// The two rings are decoupled: a raw slot is refilled as soon as the splitter has read it, so HBM latency is covered
// by n_raw chunks in flight while only two (larger) split slots exist.  Two splitter warpgroups alternate chunks.
__global__ void __launch_bounds__(kThreads, 1)
xty_tc_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_g, long long rows, int N,
              int n_raw, int raw_bytes, int split_bytes, int tmem_cols, float* __restrict__ partial, int want_colsum) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  unsigned char* split_base = base + (size_t)n_raw * raw_bytes;
  CtlX& S = *reinterpret_cast<CtlX*>(split_base + (size_t)kSplit * split_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_chunks = (int)((rows + CH - 1) / CH);
  const unsigned g_raw_bytes = (unsigned)N * CH * 4, g_tile = (unsigned)N * 128u;
  const unsigned off_graw = 16384, off_xlo = 16384, off_ghi = 32768, off_glo = off_ghi + g_tile;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kMaxRaw; ++s) {
      mbar_init(&S.full_raw[s], 1);
      mbar_init(&S.empty_raw[s], 128);
    }
    for (int s = 0; s < kSplit; ++s) {
      mbar_init(&S.full_split[s], 128);
      mbar_init(&S.empty_split[s], 1);
    }
    mbar_init(&S.tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "r"(tmem_cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = S.tmem_base;
  const size_t stride_p = (size_t)KX * N + N + KX;   // per-CTA partial: D [128 x N], colsum(A) [N], colsum(X) [128]
  const int my_chunks = blockIdx.x < n_chunks ? (n_chunks - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: raw row-major chunks =====
      for (int i = 0; i < my_chunks; ++i) {
        const int chunk = blockIdx.x + i * gridDim.x;
        const int s = i % n_raw;
        const unsigned ph = (i / n_raw) & 1;
        mbar_wait(&S.empty_raw[s], ph ^ 1);
        mbar_expect_tx(&S.full_raw[s], 16384u + g_raw_bytes);
        unsigned char* st = base + (size_t)s * raw_bytes;
        tma_load_2d(st, &map_x, &S.full_raw[s], 0, chunk * CH);
        tma_load_2d(st + off_graw, &map_g, &S.full_raw[s], 0, chunk * CH);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer: D[128 x N] += X_chunk^T . G_chunk (3 tf32 products) =====
      const unsigned idesc = make_idesc_tf32_xty(N);
      for (int i = 0; i < my_chunks; ++i) {
        const int s = i % kSplit;
        const unsigned ph = (i / kSplit) & 1;
        mbar_wait(&S.full_split[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned st = smem_u32(split_base + (size_t)s * split_bytes);
        const unsigned long long dxh = make_smem_desc(st), dxl = make_smem_desc(st + off_xlo);
        const unsigned long long dgh = make_smem_desc(st + off_ghi), dgl = make_smem_desc(st + off_glo);
#pragma unroll
        for (int k = 0; k < CH / UK; ++k) {
          const unsigned long long o = (unsigned long long)(k * 2);
          umma_tf32(tmem, dxl + o, dgh + o, idesc, (i | k) != 0 ? 1u : 0u);
          umma_tf32(tmem, dxh + o, dgl + o, idesc, 1u);
          umma_tf32(tmem, dxh + o, dgh + o, idesc, 1u);
        }
        tcgen05_commit(&S.empty_split[s]);
      }
      tcgen05_commit(&S.tmem_full);
    }
  } else if (warp >= 4 && warp < 8) {  // ===== epilogue: this CTA's partial D -> workspace =====
    const int q = warp & 3;
    float* prow = partial + (size_t)blockIdx.x * stride_p + (size_t)(q * 32 + lane) * N;
    if (my_chunks > 0) {
      mbar_wait(&S.tmem_full, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int c0 = 0; c0 < N; c0 += 16) {
        unsigned r[16];
        const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; j += 4)
          st4(prow + c0 + j, make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                                         __uint_as_float(r[j + 3])));
      }
    } else {
      for (int c0 = 0; c0 < N; c0 += 4) st4(prow + c0, make_float4(0.f, 0.f, 0.f, 0.f));
    }
  } else if (warp >= 8) {  // ===== two splitter warpgroups: transpose + hi/lo split, hand-applied 128 B swizzle =====
    const int wg = (warp - 8) >> 2;
    const int t = (threadIdx.x - 256) & 127;   // column of X; column of G when t < N
    const unsigned swz = (unsigned)(t & 7);
    float csum = 0.f, csum_x = 0.f;
    for (int i = wg; i < my_chunks; i += 2) {
      const int s = i % n_raw, q = i % kSplit;
      mbar_wait(&S.full_raw[s], (i / n_raw) & 1);
      // explicit shared-space accesses (generic pointers made these LD.E / ST.E)
      const unsigned st = smem_u32(base + (size_t)s * raw_bytes);
      const unsigned xr = st + (unsigned)t * 4u;                 // [32][128] fp32, column t
      const unsigned gr = st + off_graw + (unsigned)t * 4u;      // [32][N]  fp32, column t
      float xv[CH], gv[CH];
#pragma unroll
      for (int r = 0; r < CH; ++r) xv[r] = lds32(xr + (unsigned)(r * KX * 4));
      if (want_colsum & 2) {
#pragma unroll
        for (int r = 0; r < CH; r += 4) csum_x += (xv[r] + xv[r + 1]) + (xv[r + 2] + xv[r + 3]);
      }
      if (t < N) {
#pragma unroll
        for (int r = 0; r < CH; ++r) gv[r] = lds32(gr + (unsigned)(r * N * 4));
      }
      mbar_wait(&S.empty_split[q], ((i / kSplit) & 1) ^ 1);
      const unsigned sp = smem_u32(split_base + (size_t)q * split_bytes);
#pragma unroll
      for (int j = 0; j < CH / 4; ++j) {   // 4 consecutive rows (= K elements) -> one 16-byte chunk of K-major row t
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(xv[4 * j + 0]) & 0xFFFFE000u); l.x = xv[4 * j + 0] - h.x;
        h.y = __uint_as_float(__float_as_uint(xv[4 * j + 1]) & 0xFFFFE000u); l.y = xv[4 * j + 1] - h.y;
        h.z = __uint_as_float(__float_as_uint(xv[4 * j + 2]) & 0xFFFFE000u); l.z = xv[4 * j + 2] - h.z;
        h.w = __uint_as_float(__float_as_uint(xv[4 * j + 3]) & 0xFFFFE000u); l.w = xv[4 * j + 3] - h.w;
        const unsigned o = (unsigned)t * 128u + (((unsigned)j ^ swz) << 4);
        sts128(sp + o, h);
        sts128(sp + off_xlo + o, l);
      }
      if (t < N) {
#pragma unroll
        for (int j = 0; j < CH / 4; ++j) {
          float4 h, l;
          csum += (gv[4 * j] + gv[4 * j + 1]) + (gv[4 * j + 2] + gv[4 * j + 3]);
          h.x = __uint_as_float(__float_as_uint(gv[4 * j + 0]) & 0xFFFFE000u); l.x = gv[4 * j + 0] - h.x;
          h.y = __uint_as_float(__float_as_uint(gv[4 * j + 1]) & 0xFFFFE000u); l.y = gv[4 * j + 1] - h.y;
          h.z = __uint_as_float(__float_as_uint(gv[4 * j + 2]) & 0xFFFFE000u); l.z = gv[4 * j + 2] - h.z;
          h.w = __uint_as_float(__float_as_uint(gv[4 * j + 3]) & 0xFFFFE000u); l.w = gv[4 * j + 3] - h.w;
          const unsigned o = (unsigned)t * 128u + (((unsigned)j ^ swz) << 4);
          sts128(sp + off_ghi + o, h);
          sts128(sp + off_glo + o, l);
        }
      }
      mbar_arrive(&S.empty_raw[s]);     // raw values are in registers: the slot can be refilled
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&S.full_split[q]);
    }
    S.csum[wg][t] = csum;
    S.csum_x[wg][t] = csum_x;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if ((want_colsum & 1) && (int)threadIdx.x < N)
    partial[(size_t)blockIdx.x * stride_p + (size_t)KX * N + threadIdx.x] = S.csum[0][threadIdx.x] + S.csum[1][threadIdx.x];
  if ((want_colsum & 2) && (int)threadIdx.x >= 128 && (int)threadIdx.x < 128 + KX)
    partial[(size_t)blockIdx.x * stride_p + (size_t)KX * N + N + (threadIdx.x - 128)] =
        S.csum_x[0][threadIdx.x - 128] + S.csum_x[1][threadIdx.x - 128];
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

// out[n][m] = sum_c partial[c][m*N + n]  (transposing), colsum[n] = sum_c partial[c][128*N + n],
// colsum_x[m] = sum_c partial[c][128*N + N + m].  A block owns 32 consecutive columns of the partials' own layout (every
// load is one coalesced 128-byte line; the transposition is on the 64 KB of stores): warp w adds the partials
// c = w, w + 8, ... with four independent chains, the eight warp sums are added in warp order (fixed: deterministic).
// (The first version gave 8 lanes to an OUTPUT element: 4-byte loads at stride N, 13.5 us for 9.7 MB of partials.)
__global__ void __launch_bounds__(256)
xty_tc_reduce_kernel(const float* __restrict__ partial, int n_part, int N, float* __restrict__ out,
                     float* __restrict__ colsum, float* __restrict__ colsum_x) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const long long total = (long long)KX * N + N + KX;
  const long long e = (long long)blockIdx.x * 32 + lane;
  const bool wanted =
      e < total && (e < (long long)KX * N || (e < (long long)KX * N + N ? colsum != nullptr : colsum_x != nullptr));
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (wanted) {
    const float* p = partial + e;
    int g = w;
    for (; g + 24 < n_part; g += 32) {
      s0 += p[(size_t)g * total];
      s1 += p[(size_t)(g + 8) * total];
      s2 += p[(size_t)(g + 16) * total];
      s3 += p[(size_t)(g + 24) * total];
    }
    for (; g < n_part; g += 8) s0 += p[(size_t)g * total];
  }
  sm[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w != 0 || !wanted) return;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) s += sm[k][lane];
  if (e < (long long)KX * N) {
    const int m = (int)(e / N), n = (int)(e % N);
    out[(size_t)n * KX + m] = s;
  } else if (e < (long long)KX * N + N) {
    colsum[e - (long long)KX * N] = s;
  } else {
    colsum_x[e - (long long)KX * N - N] = s;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

int make_map_rowmajor(CUtensorMap* map, const void* basep, long long cols, long long rows, long long ld) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    mlg_set_error("mlg_xty_tc: cuTensorMapEncodeTiled entry point not available");
    return MLG_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)cols, CH};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(basep), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mlg_set_error("mlg_xty_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MLG_ERR_CUDA;
  }
  return MLG_OK;
}

int plan(int N, int* n_raw, int* raw_bytes, int* split_bytes) {
  *raw_bytes = 16384 + ((N * CH * 4 + 1023) / 1024) * 1024;
  *split_bytes = 32768 + ((2 * N * 128 + 1023) / 1024) * 1024;
  int r = (220 * 1024 - kSplit * *split_bytes) / *raw_bytes;
  if (r > kMaxRaw) r = kMaxRaw;
  *n_raw = r;
  return r >= 2;
}

}  // namespace

extern "C" int mlg_xty_tc_supported(int64_t rows, int64_t M, int64_t K) {
  int nr, rb, sb;
  return rows >= 1 && K == KX && M >= 16 && M <= 128 && M % 16 == 0 && plan((int)M, &nr, &rb, &sb);
}

extern "C" int64_t mlg_xty_tc_workspace_bytes(int64_t M) { return (int64_t)148 * 2 * (KX * M + M + KX) * 4; }

// out[M, 128] = A[rows, M]^T . X[rows, 128], colsum[M] = column sums of A, colsum_x[128] = column sums of X (NULL ok)
extern "C" int mlg_xty_tc(const float* A, int64_t ld_a, const float* X, int64_t ld_x, int64_t rows, int64_t M, int64_t K,
                          float* out, float* colsum, float* colsum_x, void* workspace, int64_t workspace_bytes,
                          void* stream) {
  MLG_CHECK_ARG(A && X && out && workspace, "mlg_xty_tc: null pointer");
  MLG_CHECK_ARG(mlg_xty_tc_supported(rows, M, K), "mlg_xty_tc: unsupported shape (need K == 128, 16 <= M <= 128, M %% 16 == 0)");
  MLG_CHECK_ARG(ld_a % 4 == 0 && ld_x % 4 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)X % 16 == 0 && rows < (1ll << 31),
                "mlg_xty_tc: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_chunks = (int)((rows + CH - 1) / CH);
  const int grid = n_chunks < sms ? n_chunks : sms;
  MLG_CHECK_ARG(workspace_bytes >= (int64_t)grid * (KX * M + M + KX) * 4, "mlg_xty_tc: workspace too small");
  CUtensorMap mx, mg;
  int rc = make_map_rowmajor(&mx, X, KX, rows, ld_x);
  if (rc) return rc;
  rc = make_map_rowmajor(&mg, A, M, rows, ld_a);
  if (rc) return rc;
  int n_raw, raw_bytes, split_bytes;
  plan((int)M, &n_raw, &raw_bytes, &split_bytes);
  int tmem_cols = 32;
  while (tmem_cols < M) tmem_cols *= 2;
  const int smem = n_raw * raw_bytes + kSplit * split_bytes + (int)sizeof(CtlX) + 1024;
  static int attr_smem = 0;
  if (smem > attr_smem) {
    MLG_CUDA(cudaFuncSetAttribute(xty_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr_smem = smem;
  }
  cudaStream_t st = (cudaStream_t)stream;
  xty_tc_kernel<<<grid, kThreads, smem, st>>>(mx, mg, rows, (int)M, n_raw, raw_bytes, split_bytes, tmem_cols, (float*)workspace,
                                             (colsum ? 1 : 0) | (colsum_x ? 2 : 0));
  MLG_CHECK_LAUNCH("mlg_xty_tc");
  const long long total = (long long)KX * M + M + KX;
  xty_tc_reduce_kernel<<<mlg_ceil_div(total, 32), 256, 0, st>>>((const float*)workspace, grid, (int)M, out, colsum,
                                                                     colsum_x);
  MLG_CHECK_LAUNCH("mlg_xty_tc(reduce)");
  return MLG_OK;
}
