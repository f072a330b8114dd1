// fp32-accurate tall GEMM on the tcgen05 tensor cores (sm_100a), 3xTF32 error-compensated:
//
//   C[M,N] = act( A[M,K] . B[N,K]^T + bias )        A fp32 row-major, M huge (nodes / edges), K <= 256, N <= 256
//
// The Linear layers around the aggregations (SAGEConv.update's MLP + lin_r folded into one weight,
// models/gcn_lib/sparse/torch_vertex.py:281-291; GENConv's per-layer edge encoder, torch_vertex.py:76-77) are
// [rows x K] x [K x N] products with rows ~ 5e5..2e6 and K, N <= 128.  In fp32 on the CUDA cores they are
// FMA-bound (8 GFLOP -> 160-200 us at the gbm shape); plain TF32 would break the rtol-1e-4 parity bar.  Here every
// fp32 operand is split into hi = top 19 bits and lo = x - hi (exact), and the tensor core computes
//   hi_A.hi_B + hi_A.lo_B + lo_A.hi_B          (fp32 accumulation in tensor memory, error ~2^-21 relative)
// so the product becomes HBM-bound.  B (the small weight) is split on the host side once and stays resident in
// shared memory; A tiles arrive by TMA (128B-swizzled rows of 32 fp32) and a splitter warpgroup rewrites each tile
// in place as `hi` and writes `lo` to a twin tile (layout-agnostic: same byte offsets), then hands the stage to the
// single MMA-issuing thread through an mbarrier (with a generic->async proxy fence).  Persistent CTAs (one per SM)
// walk 128-row tiles; the accumulator is double-buffered in tensor memory so the epilogue (bias + LeakyReLU, 128-bit
// stores) of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int BM = 128, BKF = 32, UK = 8;          // tile rows, fp32 per 128 B swizzle row, K per tf32 MMA
constexpr int kSplitThreads = 256;                 // splitter threads (ncu r01: with 128 the splitter was the bottleneck stage)
constexpr int kThreads = 256 + kSplitThreads;      // warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 epilogue, 8-15 splitter
constexpr int kTileBytes = BM * BKF * 4;           // 16 KB

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
__host__ __device__ constexpr unsigned make_idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
}

struct Ctl {   // barriers and bookkeeping, placed after the tiles
  unsigned long long full_raw[4], full_split[4], empty[4], b_full, tmem_full[2], tmem_empty[2];
  unsigned tmem_base;
  alignas(16) float bias_s[256];
};

static_assert(sizeof(Ctl) <= 2048, "Ctl must fit the 2 KB reserved after the tiles");
constexpr int kStageOutBytes = 4 * 2 * 4096;   // epilogue staging: 4 warps x 2 buffers x [32 rows x 128 B]

__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts128(unsigned addr, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// MLG_TF32X3_STORE_HINT (A/B switch, default 0 = no hint): 1 = L2 evict_first, 2 = L2 evict_last on the output tiles.  The
// kernel that reads a GEMM output right away was measured 3x slower than behind an elementwise pass (DESIGN.md section 6
// item 1); the hint decides whether the 126 MB of output lines should stream through L2 or stay in it.
#ifndef MLG_TF32X3_STORE_HINT
#define MLG_TF32X3_STORE_HINT 0
#endif
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, unsigned src, int c0, int c1) {
#if MLG_TF32X3_STORE_HINT == 0
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
#else
  unsigned long long policy;
#if MLG_TF32X3_STORE_HINT == 1
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
#else
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(policy));
#endif
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(map),
               "r"(src), "r"(c0), "r"(c1), "l"(policy)
               : "memory");
#endif
}

// kNN candidate selection as the epilogue (csrc/knn.cu, tensor-core path): the N columns of this launch are candidate
// points, bias = their squared norms (+inf past `valid`), and instead of storing C every accumulator row keeps the kc
// smallest keys  |x_j|^2 - 2 x_i.x_j  (= distance minus the row constant |x_i|^2) seen so far, with their column ids, in a
// per-row list that lives in global memory between launches and in shared memory (the output staging area) during one.
constexpr int kKnnKC = 24;   // candidates kept per row (= csrc/knn.cu kTcCand)
constexpr int kKnnLd = 28;   // shared-memory row stride of a list, words
struct KnnEpi {
  int kc;          // kKnnKC, or 0: plain GEMM epilogue
  float* d;        // [M, kc] keys, unordered
  int* i;          // [M, kc] column ids
  int col0;        // id of this launch's first column
  int valid;       // columns < valid exist
};
__device__ __forceinline__ float lds32(unsigned addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ int lds32i(unsigned addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(unsigned addr, float v) { asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory"); }
__device__ __forceinline__ void sts32i(unsigned addr, int v) { asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory"); }

// one admission: the key replaces the list's largest entry (slot pmax); the new largest and its slot come from one vector
// scan (six 128-bit loads, a max tree, a match mask)
// (returns (new threshold, bit pattern of the new slot) in registers: reference arguments of an out-of-line routine live
// in local memory)
__device__ __noinline__ float2 knn_admit(unsigned ld_u32, unsigned li_u32, float key, int id, int pmax) {
  sts32(ld_u32 + (unsigned)pmax * 4u, key);
  sts32i(li_u32 + (unsigned)pmax * 4u, id);
  float4 v[kKnnKC / 4];
#pragma unroll
  for (int e = 0; e < kKnnKC / 4; ++e) v[e] = lds128(ld_u32 + (unsigned)e * 16u);
  float m = -INFINITY;
#pragma unroll
  for (int e = 0; e < kKnnKC / 4; ++e) m = fmaxf(m, fmaxf(fmaxf(v[e].x, v[e].y), fmaxf(v[e].z, v[e].w)));
  unsigned hit = 0;
#pragma unroll
  for (int e = 0; e < kKnnKC / 4; ++e)
    hit |= ((v[e].x == m ? 1u : 0u) | (v[e].y == m ? 2u : 0u) | (v[e].z == m ? 4u : 0u) | (v[e].w == m ? 8u : 0u)) << (4 * e);
  return make_float2(m, __int_as_float(__ffs(hit) - 1));
}

// shared memory: [B_hi: kb][N x 128 B] [B_lo: kb][...] [stages][A_hi 16 KB | A_lo 16 KB] [out staging 32 KB] Ctl
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32x3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_bh,
                   const __grid_constant__ CUtensorMap map_bl, const __grid_constant__ CUtensorMap map_c,
                   const float* __restrict__ bias, float* __restrict__ C, long long ldc, int M, int N, int K,
                   int n_stages, int tmem_cols, int act, float slope, int tma_out, int skip_hi, const KnnEpi knn) {
  extern __shared__ unsigned char smem_raw[];
  unsigned char* base = reinterpret_cast<unsigned char*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int k_blocks = K / BKF;
  const unsigned b_tile = (unsigned)N * 128u;                       // bytes of one [N x 32 fp32] tile
  unsigned char* b_hi = base;
  unsigned char* b_lo = base + (size_t)k_blocks * b_tile;
  unsigned char* a_st = base + (size_t)2 * k_blocks * b_tile;       // stage s: a_st + s * 2 * kTileBytes
  unsigned char* out_st = a_st + (size_t)n_stages * 2 * kTileBytes;
  Ctl& S = *reinterpret_cast<Ctl*>(out_st + kStageOutBytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (M + BM - 1) / BM;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 4; ++s) {
      mbar_init(&S.full_raw[s], 1);
      mbar_init(&S.full_split[s], kSplitThreads);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(&S.b_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&S.tmem_full[i], 1);
      mbar_init(&S.tmem_empty[i], 128);
    }
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

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer: resident weight tiles once, then the A tiles of this CTA's row tiles =====
      mbar_expect_tx(&S.b_full, 2u * (unsigned)k_blocks * b_tile);
      for (int kb = 0; kb < k_blocks; ++kb) {
        tma_load_2d(b_hi + (size_t)kb * b_tile, &map_bh, &S.b_full, kb * BKF, 0);
        tma_load_2d(b_lo + (size_t)kb * b_tile, &map_bl, &S.b_full, kb * BKF, 0);
      }
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % n_stages;
          const unsigned ph = (it / n_stages) & 1;
          mbar_wait(&S.empty[s], ph ^ 1);
          mbar_expect_tx(&S.full_raw[s], kTileBytes);
          tma_load_2d(a_st + (size_t)s * 2 * kTileBytes, &map_a, &S.full_raw[s], kb * BKF, tile * BM);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      const unsigned idesc = make_idesc_tf32(N);
      mbar_wait(&S.b_full, 0);
      int it = 0, tl = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const int buf = tl & 1;
        mbar_wait(&S.tmem_empty[buf], ((tl >> 1) & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned d_tmem = tmem + (unsigned)(buf * N);
        for (int kb = 0; kb < k_blocks; ++kb, ++it) {
          const int s = it % n_stages;
          const unsigned ph = (it / n_stages) & 1;
          mbar_wait(&S.full_split[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned a_hi = smem_u32(a_st + (size_t)s * 2 * kTileBytes), a_lo = a_hi + kTileBytes;
          const unsigned long long dah = make_smem_desc(a_hi), dal = make_smem_desc(a_lo);
          const unsigned long long dbh = make_smem_desc(smem_u32(b_hi + (size_t)kb * b_tile));
          const unsigned long long dbl = make_smem_desc(smem_u32(b_lo + (size_t)kb * b_tile));
#pragma unroll
          for (int k = 0; k < BKF / UK; ++k) {
            const unsigned long long o = (unsigned long long)(k * 2);   // 8 fp32 = 32 B along K: +2 (16-byte units)
            umma_tf32(d_tmem, dal + o, dbh + o, idesc, (kb | k) != 0 ? 1u : 0u);   // small terms first
            umma_tf32(d_tmem, dah + o, dbl + o, idesc, 1u);
            umma_tf32(d_tmem, dah + o, dbh + o, idesc, 1u);
          }
          tcgen05_commit(&S.empty[s]);
        }
        tcgen05_commit(&S.tmem_full[buf]);
      }
    }
  } else if (warp >= 4 && warp < 8) {  // ===== epilogue =====
    const int q = warp & 3;
    for (int i = threadIdx.x - 128; i < N; i += 128)
      S.bias_s[i] = knn.kc ? (i < knn.valid ? __ldg(bias + i) : INFINITY) : (bias ? __ldg(bias + i) : 0.f);
    asm volatile("bar.sync 1, 128;" ::: "memory");
    int tl = 0;
    if (knn.kc) {
      // thread = accumulator row.  List of row r: kKnnKC keys (UNORDERED; the largest one is the admission threshold and
      // the slot the next admitted key replaces -- the refinement pass orders the survivors by their exact distances
      // anyway) then kKnnKC ids; row stride kKnnLd words (16-byte aligned rows, conflict-free 128-bit loads).
      // Cost model: a warp pays for every admission of any of its 32 rows, ~768 / m of them in the m-th column chunk, so
      // the admission itself must be cheap: two stores, six 128-bit loads, a max tree and a match mask (~100 cycles; a
      // sorted list with dependent shifts was ~800 cycles and made a 128 x 256 tile cost 29 us against 1.6 us of MMA).
      const unsigned rl = (unsigned)(q * 32 + lane);
      const unsigned ld_u32 = smem_u32(out_st) + rl * (unsigned)kKnnLd * 4u;
      const unsigned li_u32 = ld_u32 + 128u * (unsigned)kKnnLd * 4u;
      const unsigned bias_u32 = smem_u32(&S.bias_s[0]);
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const int buf = tl & 1;
        const long long row = (long long)tile * BM + rl;
        const bool ok = row < M;
        float thr = -INFINITY;   // rows past M never admit anything
        int pmax = 0;
        if (ok) {
          const float4* gd = reinterpret_cast<const float4*>(knn.d + row * kKnnKC);
          const int4* gi = reinterpret_cast<const int4*>(knn.i + row * kKnnKC);
#pragma unroll
          for (int e = 0; e < kKnnKC / 4; ++e) {
            const float4 dv = gd[e];
            const int4 iv = gi[e];
            sts128(ld_u32 + (unsigned)e * 16u, dv);
            sts128(li_u32 + (unsigned)e * 16u, make_float4(__int_as_float(iv.x), __int_as_float(iv.y), __int_as_float(iv.z),
                                                           __int_as_float(iv.w)));
            const float dd[4] = {dv.x, dv.y, dv.z, dv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (dd[u] > thr) { thr = dd[u]; pmax = 4 * e + u; }
          }
        }
        mbar_wait(&S.tmem_full[buf], (tl >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < N; c0 += 32) {
          unsigned r[32];
          const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)(buf * N + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          // all 32 keys and their admission mask first (independent instructions, ONE branch per 32 columns on the common
          // path); admitted keys are then handled nibble by nibble through an out-of-line routine.  Testing and admitting
          // column by column serialised FFMA -> FSETP -> BRA on the single epilogue warp of each scheduler and unrolled
          // the admission code 32 times (53 KB of SASS, 39 % of the stalls were instruction fetch).
          unsigned mask = 0;
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 sv = lds128(bias_u32 + (unsigned)((c0 + j4 * 4) * 4));
            const float k0 = fmaf(-2.f, __uint_as_float(r[4 * j4]), sv.x), k1 = fmaf(-2.f, __uint_as_float(r[4 * j4 + 1]), sv.y);
            const float k2 = fmaf(-2.f, __uint_as_float(r[4 * j4 + 2]), sv.z), k3 = fmaf(-2.f, __uint_as_float(r[4 * j4 + 3]), sv.w);
            r[4 * j4] = __float_as_uint(k0); r[4 * j4 + 1] = __float_as_uint(k1);
            r[4 * j4 + 2] = __float_as_uint(k2); r[4 * j4 + 3] = __float_as_uint(k3);
            mask |= ((k0 < thr ? 1u : 0u) | (k1 < thr ? 2u : 0u) | (k2 < thr ? 4u : 0u) | (k3 < thr ? 8u : 0u)) << (4 * j4);
          }
          if (mask) {
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              if ((mask >> (4 * j4)) & 15u) {
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                  const float key = __uint_as_float(r[4 * j4 + u]);
                  if (((mask >> (4 * j4 + u)) & 1u) && key < thr) {   // the threshold may have tightened since the mask
                    const float2 t = knn_admit(ld_u32, li_u32, key, knn.col0 + c0 + j4 * 4 + u, pmax);
                    thr = t.x;
                    pmax = __float_as_int(t.y);
                  }
                }
              }
            }
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&S.tmem_empty[buf]);
        if (ok) {
          float4* gd = reinterpret_cast<float4*>(knn.d + row * kKnnKC);
          float4* gi = reinterpret_cast<float4*>(knn.i + row * kKnnKC);
#pragma unroll
          for (int e = 0; e < kKnnKC / 4; ++e) {
            gd[e] = lds128(ld_u32 + (unsigned)e * 16u);
            gi[e] = lds128(li_u32 + (unsigned)e * 16u);
          }
        }
      }
    } else if (tma_out) {
      // accumulator rows (lane = row) -> bias/act -> this warp's private [32 rows x 32 cols] staging tile (128 B rows,
      // hand-applied 128 B swizzle: conflict-free STS.128) -> one TMA store per 32-column group.  Row-per-lane global
      // stores wrote 32 half-filled sectors per instruction and kept these warps 100 % busy (ncu r01).
      const unsigned st0 = smem_u32(out_st) + (unsigned)q * 8192u;
      const unsigned my_row = (unsigned)lane * 128u, swz = (unsigned)(lane & 7);
      const unsigned bias_u32 = smem_u32(&S.bias_s[0]);
      int ob = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
        const int buf = tl & 1;
        mbar_wait(&S.tmem_full[buf], (tl >> 1) & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        for (int c0 = 0; c0 < N; c0 += 32, ob ^= 1) {
          unsigned r[32];
          const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)(buf * N + c0);
          asm volatile(
              "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
              "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
              "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
              : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
              : "r"(taddr));
          // the staging buffer we are about to overwrite: its previous TMA store (two groups ago) must have read it
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          __syncwarp();
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          const unsigned sb = st0 + (unsigned)ob * 4096u + my_row;
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 bv = lds128(bias_u32 + (unsigned)((c0 + j * 4) * 4));
            float4 v = make_float4(__uint_as_float(r[4 * j]) + bv.x, __uint_as_float(r[4 * j + 1]) + bv.y,
                                   __uint_as_float(r[4 * j + 2]) + bv.z, __uint_as_float(r[4 * j + 3]) + bv.w);
            if (act) {
              v.x = v.x > 0.f ? v.x : v.x * slope;
              v.y = v.y > 0.f ? v.y : v.y * slope;
              v.z = v.z > 0.f ? v.z : v.z * slope;
              v.w = v.w > 0.f ? v.w : v.w * slope;
            }
            sts128(sb + (((unsigned)j ^ swz) << 4), v);
          }
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            tma_store_2d(&map_c, st0 + (unsigned)ob * 4096u, c0, tile * BM + q * 32);   // rows past M are clipped
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        mbar_arrive(&S.tmem_empty[buf]);
      }
      if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // stores complete before exit
    } else {
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tl) {
      const int buf = tl & 1;
      mbar_wait(&S.tmem_full[buf], (tl >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = tile * BM + q * 32 + lane;
      float* crow = C + (size_t)row * ldc;
      for (int c0 = 0; c0 < N; c0 += 16) {
        unsigned r[16];
        const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)(buf * N + c0);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (row < M) {
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float t = __uint_as_float(r[j]) + S.bias_s[c0 + j];
            if (act) t = t > 0.f ? t : t * slope;
            v[j] = t;
          }
#pragma unroll
          for (int j = 0; j < 16; j += 4) st4(crow + c0 + j, make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      mbar_arrive(&S.tmem_empty[buf]);
    }
    }
  } else if (warp >= 8) {  // ===== splitter: raw fp32 tile -> hi (in place) + lo (twin tile) =====
    const int t = threadIdx.x - 256;   // 0..kSplitThreads-1
    constexpr int kPer = kTileBytes / 16 / kSplitThreads;   // float4 per thread per stage
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < k_blocks; ++kb, ++it) {
        const int s = it % n_stages;
        const unsigned ph = (it / n_stages) & 1;
        mbar_wait(&S.full_raw[s], ph);
        // explicit shared-space accesses (the generic pointer made these LD.E / ST.E before)
        const unsigned hi = smem_u32(a_st + (size_t)s * 2 * kTileBytes) + (unsigned)t * 16u;
        const unsigned lo = hi + kTileBytes;
        float4 x[kPer];
#pragma unroll
        for (int i = 0; i < kPer; ++i) x[i] = lds128(hi + (unsigned)(i * kSplitThreads * 16));
#pragma unroll
        for (int i = 0; i < kPer; ++i) {
          float4 h, l;
          h.x = __uint_as_float(__float_as_uint(x[i].x) & 0xFFFFE000u);
          h.y = __uint_as_float(__float_as_uint(x[i].y) & 0xFFFFE000u);
          h.z = __uint_as_float(__float_as_uint(x[i].z) & 0xFFFFE000u);
          h.w = __uint_as_float(__float_as_uint(x[i].w) & 0xFFFFE000u);
          l.x = x[i].x - h.x; l.y = x[i].y - h.y; l.z = x[i].z - h.z; l.w = x[i].w - h.w;
          if (!skip_hi) sts128(hi + (unsigned)(i * kSplitThreads * 16), h);
          sts128(lo + (unsigned)(i * kSplitThreads * 16), l);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to tcgen05.mma
        mbar_arrive(&S.full_split[s]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

__global__ void split_tf32_kernel(const float* __restrict__ w, long long n, float* __restrict__ hi, float* __restrict__ lo) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = w[i];
  const float h = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  hi[i] = h;
  lo[i] = x - h;
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
// 2-D map over a K-major fp32 matrix [rows, K] (leading dimension ld): box (32, box_rows), 128 B swizzle
int make_map_f32(CUtensorMap* map, const void* basep, long long K, long long rows, long long ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    mlg_set_error("mlg_gemm_tf32x3: cuTensorMapEncodeTiled entry point not available");
    return MLG_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {BKF, (cuuint32_t)box_rows};
  cuuint32_t elem[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(basep), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mlg_set_error("mlg_gemm_tf32x3: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MLG_ERR_CUDA;
  }
  return MLG_OK;
}

int g_attr_smem = 0;

}  // namespace

extern "C" int mlg_split_tf32(const float* w, int64_t n, float* hi, float* lo, void* stream) {
  MLG_CHECK_ARG(w && hi && lo && n >= 0, "mlg_split_tf32: bad arguments");
  if (n == 0) return MLG_OK;
  split_tf32_kernel<<<mlg_ceil_div(n, 256), 256, 0, (cudaStream_t)stream>>>(w, n, hi, lo);
  MLG_CHECK_LAUNCH("mlg_split_tf32");
  return MLG_OK;
}

extern "C" int mlg_gemm_tf32x3_supported(int64_t M, int64_t N, int64_t K) {
  if (M < 1 || N < 16 || N > 256 || N % 16 || K < 32 || K > 256 || K % 32) return 0;
  const long long bbytes = 2ll * (K / 32) * N * 128;
  return (224 * 1024 - kStageOutBytes - bbytes) / (2 * kTileBytes) >= 2 ? 1 : 0;
}

extern "C" int mlg_gemm_tf32x3(const float* A, int64_t lda, const float* B_hi, const float* B_lo, const float* bias,
                               float* C, int64_t ldc, int64_t M, int64_t N, int64_t K, int act, float slope,
                               void* stream) {
  MLG_CHECK_ARG(A && B_hi && B_lo && C, "mlg_gemm_tf32x3: null pointer");
  MLG_CHECK_ARG(mlg_gemm_tf32x3_supported(M, N, K), "mlg_gemm_tf32x3: unsupported shape M=%lld N=%lld K=%lld "
                "(need N%%16==0, 16<=N<=256, K%%32==0, K<=256 and the split weights must fit in shared memory)",
                (long long)M, (long long)N, (long long)K);
  MLG_CHECK_ARG(lda % 4 == 0 && lda >= K && ldc % 4 == 0 && ldc >= N && (uintptr_t)A % 16 == 0 && (uintptr_t)C % 16 == 0 &&
                    (uintptr_t)B_hi % 16 == 0 && (uintptr_t)B_lo % 16 == 0,
                "mlg_gemm_tf32x3: operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  MLG_CHECK_ARG(M < (1ll << 31), "mlg_gemm_tf32x3: M too large");
  CUtensorMap ma, mbh, mbl;
  int rc = make_map_f32(&ma, A, K, M, lda, BM);
  if (rc) return rc;
  rc = make_map_f32(&mbh, B_hi, K, N, K, (int)N);
  if (rc) return rc;
  rc = make_map_f32(&mbl, B_lo, K, N, K, (int)N);
  if (rc) return rc;
  const long long bbytes = 2ll * (K / 32) * N * 128;
  int stages = (int)((224 * 1024 - kStageOutBytes - bbytes) / (2 * kTileBytes));
  if (stages > 4) stages = 4;
  const int tma_out = N % 32 == 0;
  // kind::tf32 ignores the low 13 mantissa bits of its fp32 operands (truncation: verified by
  // tests/test_gpu_gemm.py::test_gemm_tf32x3_is_fp32_accurate, which fails at ~1e-3 if the hardware rounded), so the raw
  // tile IS the hi operand and the splitter only has to write lo.  MLG_TF32_WRITE_HI=1 restores the explicit hi write.
  static const int skip_hi = getenv("MLG_TF32_WRITE_HI") ? 0 : 1;
  CUtensorMap mc = ma;
  if (tma_out) {
    rc = make_map_f32(&mc, C, N, M, ldc, 32);   // box: 32 columns (128 B, swizzled) x 32 rows
    if (rc) return rc;
  }
  int tmem_cols = 32;
  while (tmem_cols < 2 * N) tmem_cols *= 2;
  const int smem = (int)(bbytes + (long long)stages * 2 * kTileBytes + kStageOutBytes + 2048 + 1024);
  if (smem > g_attr_smem) {   // one high-water mark for both launchers of this kernel
    MLG_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    g_attr_smem = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = (int)((M + BM - 1) / BM);
  const int grid = n_tiles < sms ? n_tiles : sms;
  KnnEpi none;
  memset(&none, 0, sizeof(none));
  gemm_tf32x3_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma, mbh, mbl, mc, bias, C, ldc, (int)M, (int)N,
                                                                    (int)K, stages, tmem_cols, act, slope, tma_out, skip_hi,
                                                                    none);
  MLG_CHECK_LAUNCH("mlg_gemm_tf32x3");
  return MLG_OK;
}

// One column chunk of the tensor-core kNN candidate search (csrc/knn.cu): rows = the M points of a graph (A [M, K], K a
// multiple of 32, zero padded), columns = `valid` <= Nc candidate points starting at id col0 (B_hi / B_lo [valid, K],
// pre-split), sq_cols = their squared norms.  Updates the per-row candidate lists (see KnnEpi).
int mlg_tf32x3_knn_chunk(const float* A, const float* B_hi, const float* B_lo, const float* sq_cols, int64_t M, int64_t Nc,
                         int64_t valid, int64_t K, int kc, float* list_d, int* list_i, int col0, void* stream) {
  MLG_CHECK_ARG(mlg_gemm_tf32x3_supported(M, Nc, K) && valid >= 1 && valid <= Nc && kc == kKnnKC && 128 * kKnnLd * 8 <= kStageOutBytes,
                "mlg_tf32x3_knn_chunk: unsupported shape");
  CUtensorMap ma, mbh, mbl;
  int rc = make_map_f32(&ma, A, K, M, K, BM);
  if (rc) return rc;
  rc = make_map_f32(&mbh, B_hi, K, valid, K, (int)Nc);    // rows past `valid` are zero-filled by TMA
  if (rc) return rc;
  rc = make_map_f32(&mbl, B_lo, K, valid, K, (int)Nc);
  if (rc) return rc;
  const long long bbytes = 2ll * (K / 32) * Nc * 128;
  int stages = (int)((224 * 1024 - kStageOutBytes - bbytes) / (2 * kTileBytes));
  if (stages > 4) stages = 4;
  static const int skip_hi = getenv("MLG_TF32_WRITE_HI") ? 0 : 1;
  int tmem_cols = 32;
  while (tmem_cols < 2 * Nc) tmem_cols *= 2;
  const int smem = (int)(bbytes + (long long)stages * 2 * kTileBytes + kStageOutBytes + 2048 + 1024);
  if (smem > g_attr_smem) {   // one high-water mark for both launchers of this kernel
    MLG_CUDA(cudaFuncSetAttribute(gemm_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    g_attr_smem = smem;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_tiles = (int)((M + BM - 1) / BM);
  const int grid = n_tiles < sms ? n_tiles : sms;
  KnnEpi e;
  e.kc = kc; e.d = list_d; e.i = list_i; e.col0 = col0; e.valid = (int)valid;
  gemm_tf32x3_kernel<<<grid, kThreads, smem, (cudaStream_t)stream>>>(ma, mbh, mbl, ma, sq_cols, nullptr, 0, (int)M, (int)Nc,
                                                                    (int)K, stages, tmem_cols, 0, 0.f, 0, skip_hi, e);
  MLG_CHECK_LAUNCH("mlg_tf32x3_knn_chunk");
  return MLG_OK;
}
