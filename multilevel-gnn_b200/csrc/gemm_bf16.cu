// bf16 x bf16 -> fp32 GEMM on the 5th-generation tensor cores (sm_100a): TMA-fed shared-memory tiles,
// tcgen05.mma issued by one thread, accumulator in tensor memory, tcgen05.ld epilogue.
//
//   C[b][M,N] = alpha * A[b][M,K] . B[b][N,K]^T          ("TN": both operands K-major, bf16; C fp32 row-major)
//
// This is the contraction engine of DiffPool at sizes where the assignment products really are dense
// GEMMs (models/diff_pooling.py:61-64 -> PyG dense_diff_pool / DenseSAGEConv):  S^T.X, S^T.A, (S^T.A).S,
// A.X and S.S^T; dense_ops.py casts / transposes the fp32 operands into K-major bf16 once per product.
//
// Structure (one 128 x BN output tile per CTA, BN = 256 or 128):
//   warp 0   TMA producer: cp.async.bulk.tensor (3-D maps: k, row, batch) into a kStages-deep ring of
//            128B-swizzled tiles, completion on `full` mbarriers
//   warp 1   MMA issuer: one elected lane issues BLOCK_K/16 tcgen05.mma (M=128, N=BN, K=16) per stage,
//            tcgen05.commit releases the stage (`empty`) and finally signals `tmem_full`
//   warp 2   allocates / frees BN tensor-memory columns
//   warps 4-7 epilogue: tcgen05.ld 32x32b (lane quarter = warp % 4) -> registers -> 128-bit global stores
// Tensor-bound: flops = 2*M*N*K per batch entry.
#include <cuda.h>
#include <cuda_bf16.h>
#include <stdlib.h>

#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int kThreads = 256;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "MLG_WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra MLG_DONE_%=;\n"
      "bra MLG_WAIT_%=;\n"
      "MLG_DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(unsigned long long* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc]
__device__ __forceinline__ void umma_bf16(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                          unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// shared-memory matrix descriptor: K-major tile, 128-byte swizzle, rows of 64 bf16 (128 B), 8-row groups 1024 B apart
__device__ __forceinline__ unsigned long long make_smem_desc(const void* tile) {
  unsigned long long d = 0;
  d |= (unsigned long long)((smem_u32(tile) & 0x3FFFF) >> 4);   // start address, 16-byte units
  d |= (unsigned long long)1 << 16;                             // leading byte offset (unused for swizzled K-major)
  d |= (unsigned long long)(1024 >> 4) << 32;                   // stride byte offset between 8-row groups
  d |= (unsigned long long)1 << 46;                             // descriptor version (sm_100)
  d |= (unsigned long long)2 << 61;                             // layout: SWIZZLE_128B
  return d;
}
// instruction descriptor, kind::f16: D=f32, A=B=bf16, both K-major, M=128, N=BN
__host__ __device__ constexpr unsigned make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(BM >> 4) << 24);
}

template <int BN, int kStages>
struct Smem {
  alignas(1024) __nv_bfloat16 a[kStages][BM * BK];
  alignas(1024) __nv_bfloat16 b[kStages][BN * BK];
  unsigned long long full[kStages], empty[kStages], tmem_full;
  unsigned tmem_base;
};

template <int BN, int kStages>
__global__ void __launch_bounds__(kThreads, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 float* __restrict__ C, long long ldc, long long stride_c, int M, int N, int K, float alpha) {
  extern __shared__ unsigned char smem_raw[];
  Smem<BN, kStages>& S = *reinterpret_cast<Smem<BN, kStages>*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, batch = blockIdx.z;
  const int k_blocks = (K + BK - 1) / BK;
  constexpr unsigned kStageBytes = (BM + BN) * BK * 2;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(&S.tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // tensor-memory allocation: BN fp32 accumulator columns (power of two >= 32)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)),
                 "n"(BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer =====
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % kStages;
        const unsigned ph = (kb / kStages) & 1;
        mbar_wait(&S.empty[s], ph ^ 1);
        mbar_expect_tx(&S.full[s], kStageBytes);
        tma_load_3d(S.a[s], &map_a, &S.full[s], kb * BK, m0, batch);
        tma_load_3d(S.b[s], &map_b, &S.full[s], kb * BK, n0, batch);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {  // ===== MMA issuer =====
      constexpr unsigned idesc = make_idesc(BN);
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % kStages;
        const unsigned ph = (kb / kStages) & 1;
        mbar_wait(&S.full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned long long da = make_smem_desc(S.a[s]), db = make_smem_desc(S.b[s]);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // advance 16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in the 16-byte start-address field
          umma_bf16(tmem, da + (unsigned long long)(k * 2), db + (unsigned long long)(k * 2), idesc,
                    (kb | k) != 0 ? 1u : 0u);
        }
        tcgen05_commit(&S.empty[s]);  // frees the stage once the MMAs above have read it
      }
      tcgen05_commit(&S.tmem_full);  // accumulator complete
    }
  } else if (warp >= 4) {  // ===== epilogue: TMEM -> registers -> global =====
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    mbar_wait(&S.tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + q * 32 + lane;
    float* crow = C + (size_t)batch * stride_c + (size_t)row * ldc;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      unsigned r[32];
      const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < M) {
        const int col = n0 + c0;
        if (col + 32 <= N && (ldc % 4) == 0 && ((uintptr_t)(crow + col) % 16) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            st_stream4(crow + col + j, make_float4(alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                                                   alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col + j < N) crow[col + j] = alpha * __uint_as_float(r[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// cta_group::2 variant: a cluster of two CTAs (one SM pair) computes one 256 x 256 output tile.
// Each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 columns), so the per-SM
// shared-memory traffic (TMA fill + MMA operand reads) drops from ~190 to ~128 B/clk -- the 1-CTA kernel
// above is limited by exactly that.  The leader CTA (rank 0) issues tcgen05.mma.cta_group::2 (M=256, N=256,
// K=16); both CTAs' TMA loads complete on the LEADER's `full` barrier (peer-masked mbarrier address),
// tcgen05.commit multicasts the `empty` / `tmem_full` arrivals to both CTAs, and each CTA drains its own
// 128 accumulator rows from its own tensor memory.
// ------------------------------------------------------------------------------------------------
template <int kStages>
struct Smem2 {
  alignas(1024) __nv_bfloat16 a[kStages][BM * BK];
  alignas(1024) __nv_bfloat16 b[kStages][128 * BK];
  unsigned long long full[kStages], empty[kStages], tmem_full;
  unsigned tmem_base;
};

__device__ __forceinline__ unsigned cluster_ctarank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion is signalled on the mbarrier at the same offset in the pair's LEADER CTA
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, unsigned long long* bar, int c0, int c1,
                                                int c2) {
  const unsigned bar_leader = smem_u32(bar) & 0xFEFFFFFFu;   // clear the CTA-rank bit of the shared::cluster address
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)),
      "l"(map), "r"(bar_leader), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_2sm(unsigned long long* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
      "h"((unsigned short)3)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(unsigned tmem_d, unsigned long long adesc, unsigned long long bdesc,
                                              unsigned idesc, unsigned accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_2cta_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                      float* __restrict__ C, long long ldc, long long stride_c, int M, int N, int K, float alpha) {
  extern __shared__ unsigned char smem_raw[];
  Smem2<kStages>& S = *reinterpret_cast<Smem2<kStages>*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int n0 = blockIdx.y * 256, m0 = (blockIdx.x >> 1) * 256 + (int)rank * 128, batch = blockIdx.z;
  const int k_blocks = (K + BK - 1) / BK;
  constexpr unsigned kStageBytes = (BM + 128) * BK * 2;   // per CTA

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    mbar_init(&S.tmem_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // one warp of EACH CTA of the pair takes part in the paired allocation
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(256)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs): own 128 rows of A, own half of the B tile =====
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % kStages;
        const unsigned ph = (kb / kStages) & 1;
        mbar_wait(&S.empty[s], ph ^ 1);
        if (leader) mbar_expect_tx(&S.full[s], 2 * kStageBytes);   // bytes of both CTAs land on the leader's barrier
        tma_load_3d_2sm(S.a[s], &map_a, &S.full[s], kb * BK, m0, batch);
        tma_load_3d_2sm(S.b[s], &map_b, &S.full[s], kb * BK, n0 + (int)rank * 128, batch);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {  // ===== MMA issuer (leader CTA only) =====
      constexpr unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(256 >> 3) << 17) | ((unsigned)(256 >> 4) << 24);
      for (int kb = 0; kb < k_blocks; ++kb) {
        const int s = kb % kStages;
        const unsigned ph = (kb / kStages) & 1;
        mbar_wait(&S.full[s], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned long long da = make_smem_desc(S.a[s]), db = make_smem_desc(S.b[s]);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k)
          umma_bf16_2sm(tmem, da + (unsigned long long)(k * 2), db + (unsigned long long)(k * 2), idesc,
                        (kb | k) != 0 ? 1u : 0u);
        tcgen05_commit_2sm(&S.empty[s]);   // frees the stage in BOTH CTAs
      }
      tcgen05_commit_2sm(&S.tmem_full);    // accumulators complete in BOTH CTAs
    }
  } else if (warp >= 4) {  // ===== epilogue: this CTA's 128 rows =====
    const int q = warp & 3;
    mbar_wait(&S.tmem_full, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int row = m0 + q * 32 + lane;
    float* crow = C + (size_t)batch * stride_c + (size_t)row * ldc;
#pragma unroll 1
    for (int c0 = 0; c0 < 256; c0 += 32) {
      unsigned r[32];
      const unsigned taddr = tmem + ((unsigned)(q * 32) << 16) + (unsigned)c0;
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
          "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
          "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
          : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
            "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
            "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
            "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
          : "r"(taddr));
      asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
      if (row < M) {
        const int col = n0 + c0;
        if (col + 32 <= N && (ldc % 4) == 0 && ((uintptr_t)(crow + col) % 16) == 0) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            st_stream4(crow + col + j, make_float4(alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                                                   alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3])));
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (col + j < N) crow[col + j] = alpha * __uint_as_float(r[j]);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();   // the peer may still be reading operands / tensor memory of this CTA's pair allocation
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(256) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// Stream-K variant of the SM-pair kernel: a PERSISTENT grid of pairs (one per co-resident cluster) splits the
// linear stream of (256 x 256 tile, 64-wide k block) units evenly, so the machine is full whatever the tile count is
// (S^T.X at 2500 x 1024 x 10000 has 40 tiles for 74 pairs: 0.45 of the tensor peak with one tile per pair).
// Hybrid schedule (data-parallel waves + stream-K remainder): while at least `pairs` whole tiles are left, pair p takes
// tile wave * pairs + p -- the pairs of a wave walk k in lockstep, so the ~10 tiles that share an operand block fetch it
// from DRAM once (a pure stream-K split of S^T.A put every pair at a different k offset of a different B block: 370 MB
// of live operand blocks for a 126 MB L2, 0.78 -> 0.58 of the tensor peak) -- and only the last tiles % pairs tiles are
// split by k.  A pair's stream-K range covers: at most one tile TAIL first (k range not starting at 0: the partial accumulator goes to the
// workspace slot of this CTA, then a release flag), whole tiles, and at most one tile HEAD last (k range starting at 0
// but not complete: the owner; it adds the partials of the following pairs in pair order -- a fixed order, so results
// are reproducible -- and stores the tile).  Waits only ever point at a higher pair's FIRST segment, which waits on
// nothing: no cycle, and no co-residency requirement.
// Two 256-column accumulators in tensor memory alternate between segments, so the epilogue of one segment overlaps the
// MMAs of the next (`tmem_full[2]` from the issuer, `tmem_empty[2]` on the leader with one arrival per epilogue warp
// of both CTAs).  Partial layout per CTA: float4 index (col / 4) * 128 + row, so a warp stores 512 contiguous bytes.
// Flags are reset by their (single) consumer: the workspace must be zero ONCE, when it is allocated.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxPairs = 80;
constexpr long long kPartialFloats = 128 * 256;   // per CTA

template <int kStages>
struct SmemSK {
  alignas(1024) __nv_bfloat16 a[kStages][BM * BK];
  alignas(1024) __nv_bfloat16 b[kStages][128 * BK];
  unsigned long long full[kStages], empty[kStages], tmem_full[2], tmem_empty[2];
  unsigned tmem_base;
};

struct SKSeg {
  int mt, nt, batch, ka, ke;
  long long tile;
};

__device__ __forceinline__ void mbar_arrive_cluster(unsigned long long* bar, unsigned cta) {
  unsigned remote;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(smem_u32(bar)), "r"(cta));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int kStages>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_bf16_streamk_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                         float* __restrict__ C, long long ldc, long long stride_c, int M, int N, int K, float alpha,
                         int tiles_m, int tiles_n, int full_waves, int pairs, long long total_units,
                         long long units_per_pair, float* __restrict__ ws_partial, unsigned* __restrict__ ws_flags) {
  extern __shared__ unsigned char smem_raw[];
  SmemSK<kStages>& S = *reinterpret_cast<SmemSK<kStages>*>(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int k_blocks = (K + BK - 1) / BK;
  const long long u0 = (long long)pair * units_per_pair;
  const long long u1 = min(total_units, u0 + units_per_pair);
  constexpr unsigned kStageBytes = (BM + 128) * BK * 2;   // per CTA

  // total_units / units_per_pair / u0 / u1 count the k blocks of the stream-K remainder (tiles after the whole waves);
  // s.tile is the tile index inside that remainder (-1 for a wave tile)
  struct Cursor {
    int wave;
    long long u;
  };
  auto next_segment = [&](Cursor& c, SKSeg& s) {
    long long tile;
    if (c.wave < full_waves) {
      tile = (long long)c.wave * pairs + pair;
      ++c.wave;
      s.tile = -1;
      s.ka = 0;
      s.ke = k_blocks;
    } else if (c.u < u1) {
      s.tile = c.u / k_blocks;
      s.ka = (int)(c.u - s.tile * k_blocks);
      s.ke = (int)min((long long)k_blocks, s.ka + (u1 - c.u));
      c.u += s.ke - s.ka;
      tile = (long long)full_waves * pairs + s.tile;
    } else {
      return false;
    }
    s.mt = (int)(tile % tiles_m);
    const long long r = tile / tiles_m;
    s.nt = (int)(r % tiles_n);
    s.batch = (int)(r / tiles_n);
    return true;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&S.full[s], 1);
      mbar_init(&S.empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&S.tmem_full[b], 1);
      mbar_init(&S.tmem_empty[b], 8);   // 4 epilogue warps of each CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "n"(512)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const unsigned tmem = S.tmem_base;

  if (warp == 0) {
    if (lane == 0) {  // ===== TMA producer (both CTAs) =====
      unsigned it = 0;
      Cursor cur = {0, u0};
      SKSeg sg;
      while (next_segment(cur, sg)) {
        for (int kb = sg.ka; kb < sg.ke; ++kb, ++it) {
          const int s = it % kStages;
          const unsigned ph = (it / kStages) & 1;
          mbar_wait(&S.empty[s], ph ^ 1);
          if (leader) mbar_expect_tx(&S.full[s], 2 * kStageBytes);
          tma_load_3d_2sm(S.a[s], &map_a, &S.full[s], kb * BK, sg.mt * 256 + (int)rank * 128, sg.batch);
          tma_load_3d_2sm(S.b[s], &map_b, &S.full[s], kb * BK, sg.nt * 256 + (int)rank * 128, sg.batch);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {  // ===== MMA issuer (leader CTA only) =====
      constexpr unsigned idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(256 >> 3) << 17) | ((unsigned)(256 >> 4) << 24);
      unsigned it = 0, si = 0;
      Cursor cur = {0, u0};
      SKSeg sg;
      for (; next_segment(cur, sg); ++si) {
        const unsigned buf = si & 1;
        mbar_wait(&S.tmem_empty[buf], ((si >> 1) & 1) ^ 1);   // the epilogue two segments back has drained this accumulator
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const unsigned acc = tmem + buf * 256;
        for (int kb = sg.ka; kb < sg.ke; ++kb, ++it) {
          const int s = it % kStages;
          const unsigned ph = (it / kStages) & 1;
          mbar_wait(&S.full[s], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const unsigned long long da = make_smem_desc(S.a[s]), db = make_smem_desc(S.b[s]);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k)
            umma_bf16_2sm(acc, da + (unsigned long long)(k * 2), db + (unsigned long long)(k * 2), idesc,
                          (kb > sg.ka || k != 0) ? 1u : 0u);
          tcgen05_commit_2sm(&S.empty[s]);
        }
        tcgen05_commit_2sm(&S.tmem_full[buf]);
      }
    }
  } else if (warp >= 4) {  // ===== epilogue: this CTA's 128 rows of every segment =====
    const int q = warp & 3;
    const int rl = q * 32 + lane;   // row inside the CTA's 128
    unsigned si = 0;
    Cursor cur = {0, u0};
    SKSeg sg;
    for (; next_segment(cur, sg); ++si) {
      const unsigned buf = si & 1;
      const bool partial = sg.ka > 0;
      const bool owner = !partial && sg.ke < k_blocks;
      // owner: pairs pair+1 .. last_pp start inside this tile
      int n_contrib = 0;
      if (owner) {
        const long long tile_end = (sg.tile + 1) * k_blocks;
        for (long long pp = pair + 1; pp * units_per_pair < tile_end && pp * units_per_pair < total_units; ++pp) ++n_contrib;
        for (int c = 0; c < n_contrib; ++c) {
          const unsigned* f = ws_flags + ((size_t)(pair + 1 + c) * 2 + rank) * 4 + q;
          if (lane == 0)
            while (ld_acquire_u32(f) == 0) __nanosleep(64);
        }
        __syncwarp();
      }
      mbar_wait(&S.tmem_full[buf], (si >> 1) & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int row = sg.mt * 256 + (int)rank * 128 + rl;
      const int n0 = sg.nt * 256;
      float* crow = C + (size_t)sg.batch * stride_c + (size_t)row * ldc;
      float4* mine = reinterpret_cast<float4*>(ws_partial + ((size_t)pair * 2 + rank) * kPartialFloats) + rl;
#pragma unroll 1
      for (int c0 = 0; c0 < 256; c0 += 32) {
        unsigned r[32];
        const unsigned taddr = tmem + buf * 256 + ((unsigned)(q * 32) << 16) + (unsigned)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
              "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
              "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (partial) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            __stcg(mine + (size_t)((c0 + j) >> 2) * 128,
                   make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                               __uint_as_float(r[j + 3])));
          continue;
        }
        for (int c = 0; c < n_contrib; ++c) {   // fixed order: pair + 1, pair + 2, ...
          const float4* theirs =
              reinterpret_cast<const float4*>(ws_partial + ((size_t)(pair + 1 + c) * 2 + rank) * kPartialFloats) + rl;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            const float4 v = __ldcg(theirs + (size_t)((c0 + j) >> 2) * 128);
            r[j] = __float_as_uint(__uint_as_float(r[j]) + v.x);
            r[j + 1] = __float_as_uint(__uint_as_float(r[j + 1]) + v.y);
            r[j + 2] = __float_as_uint(__uint_as_float(r[j + 2]) + v.z);
            r[j + 3] = __float_as_uint(__uint_as_float(r[j + 3]) + v.w);
          }
        }
        if (row < M) {
          const int col = n0 + c0;
          if (col + 32 <= N && (ldc % 4) == 0 && ((uintptr_t)(crow + col) % 16) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              st_stream4(crow + col + j, make_float4(alpha * __uint_as_float(r[j]), alpha * __uint_as_float(r[j + 1]),
                                                     alpha * __uint_as_float(r[j + 2]), alpha * __uint_as_float(r[j + 3])));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (col + j < N) crow[col + j] = alpha * __uint_as_float(r[j]);
          }
        }
      }
      // this warp's quarter of the accumulator is drained: hand the buffer back to the issuer
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(&S.tmem_empty[buf], 0);
      if (partial) {
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release_u32(ws_flags + ((size_t)pair * 2 + rank) * 4 + q, 1u);
      } else if (n_contrib > 0) {
        __syncwarp();
        if (lane == 0)
          for (int c = 0; c < n_contrib; ++c) ws_flags[((size_t)(pair + 1 + c) * 2 + rank) * 4 + q] = 0u;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  cluster_sync_all();
  if (warp == 2) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(512) : "memory");
  }
}

// ---- fp32 -> bf16 cast (optionally transposing) so that any operand becomes K-major ----------------
__global__ void cast_bf16_kernel(const float* __restrict__ src, long long ld_src, long long rows, long long cols,
                                 __nv_bfloat16* __restrict__ dst, long long ld_dst) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long b = blockIdx.y;
  if (i >= rows * cols) return;
  const long long r = i / cols, c = i % cols;
  dst[b * rows * ld_dst + r * ld_dst + c] = __float2bfloat16_rn(src[b * rows * ld_src + r * ld_src + c]);
}
// dst[b][c][r] = src[b][r][c]; 32x32 shared-memory tiles, both sides coalesced
__global__ void cast_bf16_transpose_kernel(const float* __restrict__ src, long long ld_src, long long rows,
                                           long long cols, __nv_bfloat16* __restrict__ dst, long long ld_dst) {
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  const float* s = src + b * rows * ld_src;
  __nv_bfloat16* d = dst + b * cols * ld_dst;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? s[r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const long long c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) d[c * ld_dst + r] = __float2bfloat16_rn(tile[threadIdx.x][j]);
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

// 3-D map over a K-major bf16 operand: dims (K, rows, batch); box (64, box_rows, 1); 128 B swizzle
int make_map(CUtensorMap* map, const void* base, long long K, long long rows, long long ld, long long batch,
             long long batch_stride, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) {
    mlg_set_error("mlg_gemm_bf16: cuTensorMapEncodeTiled entry point not available");
    return MLG_ERR_CUDA;
  }
  cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)batch};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)(batch > 1 ? batch_stride : rows * ld) * 2};
  cuuint32_t box[3] = {BK, (cuuint32_t)box_rows, 1};
  cuuint32_t elem[3] = {1, 1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, elem,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mlg_set_error("mlg_gemm_bf16: cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MLG_ERR_CUDA;
  }
  return MLG_OK;
}

template <int BN, int kStages>
int launch_gemm(const CUtensorMap& ma, const CUtensorMap& mb, float* C, long long ldc, long long stride_c, int M, int N,
                int K, int batch, float alpha, cudaStream_t st) {
  const int smem = (int)sizeof(Smem<BN, kStages>) + 1024;
  static bool done = false;
  if (!done) {
    MLG_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
  gemm_bf16_kernel<BN, kStages><<<grid, kThreads, smem, st>>>(ma, mb, C, ldc, stride_c, M, N, K, alpha);
  MLG_CHECK_LAUNCH("mlg_gemm_bf16");
  return MLG_OK;
}

}  // namespace

template <int kStages>
int launch_gemm_2cta(const CUtensorMap& ma, const CUtensorMap& mb, float* C, long long ldc, long long stride_c, int M,
                     int N, int K, int batch, float alpha, cudaStream_t st) {
  const int smem = (int)sizeof(Smem2<kStages>) + 1024;
  static bool done = false;
  if (!done) {
    MLG_CUDA(cudaFuncSetAttribute(gemm_bf16_2cta_kernel<kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  dim3 grid(2 * ((M + 255) / 256), (N + 255) / 256, batch);   // x: (256-row tile, CTA rank); cluster = 2 along x
  gemm_bf16_2cta_kernel<kStages><<<grid, kThreads, smem, st>>>(ma, mb, C, ldc, stride_c, M, N, K, alpha);
  MLG_CHECK_LAUNCH("mlg_gemm_bf16(2cta)");
  return MLG_OK;
}

// co-resident SM pairs for the stream-K kernel (cached per device)
template <int kStages>
int streamk_pairs(int smem) {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
  if (cached[dev]) return cached[dev];
  int pairs = 0, sms = 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * kMaxPairs, 1, 1);
  cfg.blockDim = dim3(kThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute at;
  at.id = cudaLaunchAttributeClusterDimension;
  at.val.clusterDim.x = 2;
  at.val.clusterDim.y = 1;
  at.val.clusterDim.z = 1;
  cfg.attrs = &at;
  cfg.numAttrs = 1;
  if (cudaOccupancyMaxActiveClusters(&pairs, gemm_bf16_streamk_kernel<kStages>, &cfg) != cudaSuccess || pairs <= 0) {
    cudaGetLastError();
    pairs = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms >= 2 ? sms / 2 : 1;
  }
  if (pairs > kMaxPairs) pairs = kMaxPairs;
  cached[dev] = pairs;
  return pairs;
}

template <int kStages>
int launch_gemm_streamk(const CUtensorMap& ma, const CUtensorMap& mb, float* C, long long ldc, long long stride_c, int M,
                        int N, int K, int batch, float alpha, void* workspace, cudaStream_t st) {
  const int smem = (int)sizeof(SmemSK<kStages>) + 1024;
  static bool done = false;
  if (!done) {
    MLG_CUDA(cudaFuncSetAttribute(gemm_bf16_streamk_kernel<kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    done = true;
  }
  const int tiles_m = (M + 255) / 256, tiles_n = (N + 255) / 256;
  const long long k_blocks = (K + BK - 1) / BK;
  const long long tiles = (long long)tiles_m * tiles_n * batch;
  const long long pairs = streamk_pairs<kStages>(smem);
  const long long waves = tiles / pairs;                        // whole data-parallel waves
  const long long total = (tiles - waves * pairs) * k_blocks;   // k blocks of the stream-K remainder
  const long long upp = total ? (total + pairs - 1) / pairs : 1;
  const long long used = waves ? pairs : (total + upp - 1) / upp;
  float* partial = (float*)workspace;
  unsigned* flags = (unsigned*)(partial + 2 * kMaxPairs * kPartialFloats);
  gemm_bf16_streamk_kernel<kStages><<<dim3((unsigned)(2 * used)), kThreads, smem, st>>>(
      ma, mb, C, ldc, stride_c, M, N, K, alpha, tiles_m, tiles_n, (int)waves, (int)pairs, total, upp, partial, flags);
  MLG_CHECK_LAUNCH("mlg_gemm_bf16(stream-K)");
  return MLG_OK;
}

extern "C" int64_t mlg_gemm_bf16_workspace_bytes(void) {
  return (int64_t)(2 * kMaxPairs * kPartialFloats * sizeof(float) + 2 * kMaxPairs * 4 * sizeof(unsigned));
}

extern "C" int mlg_cast_bf16(const float* src, int64_t ld_src, int64_t rows, int64_t cols, int64_t batch,
                             int transpose, void* dst_bf16, int64_t ld_dst, void* stream) {
  MLG_CHECK_ARG(src && dst_bf16 && rows > 0 && cols > 0 && batch > 0, "mlg_cast_bf16: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (transpose) {
    MLG_CHECK_ARG(batch < 65536 && (rows + 31) / 32 < 65536, "mlg_cast_bf16: grid too large");
    dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32), (unsigned)batch), block(32, 8);
    cast_bf16_transpose_kernel<<<grid, block, 0, st>>>(src, ld_src, rows, cols, (__nv_bfloat16*)dst_bf16, ld_dst);
  } else {
    MLG_CHECK_ARG(batch < 65536, "mlg_cast_bf16: batch too large");
    dim3 grid((unsigned)mlg_ceil_div(rows * cols, 256), (unsigned)batch);
    cast_bf16_kernel<<<grid, 256, 0, st>>>(src, ld_src, rows, cols, (__nv_bfloat16*)dst_bf16, ld_dst);
  }
  MLG_CHECK_LAUNCH("mlg_cast_bf16");
  return MLG_OK;
}

extern "C" int mlg_gemm_bf16(const void* A, int64_t lda, int64_t stride_a, const void* B, int64_t ldb,
                             int64_t stride_b, float* C, int64_t ldc, int64_t stride_c, int64_t M, int64_t N,
                             int64_t K, int64_t batch, float alpha, void* stream) {
  MLG_CHECK_ARG(A && B && C, "mlg_gemm_bf16: null pointer");
  MLG_CHECK_ARG(M > 0 && N > 0 && K > 0 && batch > 0 && batch < 65536 && M < (1ll << 31) && N < (1ll << 31) &&
                    K < (1ll << 31),
                "mlg_gemm_bf16: bad sizes");
  MLG_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0 &&
                    (batch == 1 || (stride_a % 8 == 0 && stride_b % 8 == 0)),
                "mlg_gemm_bf16: TMA needs 16-byte aligned operands with leading dimensions that are multiples of 8");
  MLG_CHECK_ARG(lda >= K && ldb >= K && ldc >= N, "mlg_gemm_bf16: leading dimension too small");
  cudaStream_t st = (cudaStream_t)stream;
  CUtensorMap ma, mb;
  // SM-pair kernel when the grid of 256 x 256 tiles fills the machine (MLG_GEMM_1CTA=1 forces the single-CTA kernel)
  {
    static const bool force_1cta = getenv("MLG_GEMM_1CTA") != nullptr;
    const long long tiles = ((M + 255) / 256) * ((N + 255) / 256) * batch;
    if (!force_1cta && M >= 256 && N >= 256 && tiles >= 64) {
      int rc2 = make_map(&ma, A, K, M, lda, batch, stride_a, BM);
      if (rc2) return rc2;
      rc2 = make_map(&mb, B, K, N, ldb, batch, stride_b, 128);
      if (rc2) return rc2;
      return launch_gemm_2cta<6>(ma, mb, C, ldc, stride_c, (int)M, (int)N, (int)K, (int)batch, alpha, st);
    }
  }
  // 256-wide tiles halve the A re-reads, but only pay when they still fill the 148 SMs
  const long long tiles256 = ((M + BM - 1) / BM) * ((N + 255) / 256) * batch;
  const bool wide = N > 128 && tiles256 >= 64;   // measured: 80 wide tiles (743 TF) beat 160 narrow ones (575 TF)
  int rc = make_map(&ma, A, K, M, lda, batch, stride_a, BM);
  if (rc) return rc;
  rc = make_map(&mb, B, K, N, ldb, batch, stride_b, wide ? 256 : 128);
  if (rc) return rc;
  if (wide) return launch_gemm<256, 4>(ma, mb, C, ldc, stride_c, (int)M, (int)N, (int)K, (int)batch, alpha, st);
  return launch_gemm<128, 6>(ma, mb, C, ldc, stride_c, (int)M, (int)N, (int)K, (int)batch, alpha, st);
}

extern "C" int mlg_gemm_bf16_ws(const void* A, int64_t lda, int64_t stride_a, const void* B, int64_t ldb,
                                int64_t stride_b, float* C, int64_t ldc, int64_t stride_c, int64_t M, int64_t N,
                                int64_t K, int64_t batch, float alpha, void* workspace, int64_t workspace_bytes,
                                void* stream) {
  static const bool no_streamk = getenv("MLG_GEMM_NO_STREAMK") != nullptr;
  if (no_streamk || !workspace || M < 256 || N < 256)
    return mlg_gemm_bf16(A, lda, stride_a, B, ldb, stride_b, C, ldc, stride_c, M, N, K, batch, alpha, stream);
  MLG_CHECK_ARG(A && B && C, "mlg_gemm_bf16_ws: null pointer");
  MLG_CHECK_ARG(workspace_bytes >= mlg_gemm_bf16_workspace_bytes() && (uintptr_t)workspace % 16 == 0,
                "mlg_gemm_bf16_ws: workspace smaller than mlg_gemm_bf16_workspace_bytes() or not 16-byte aligned");
  MLG_CHECK_ARG(K > 0 && batch > 0 && batch < 65536 && M < (1ll << 31) && N < (1ll << 31) && K < (1ll << 31),
                "mlg_gemm_bf16_ws: bad sizes");
  MLG_CHECK_ARG(lda % 8 == 0 && ldb % 8 == 0 && (uintptr_t)A % 16 == 0 && (uintptr_t)B % 16 == 0 &&
                    (batch == 1 || (stride_a % 8 == 0 && stride_b % 8 == 0)),
                "mlg_gemm_bf16_ws: TMA needs 16-byte aligned operands with leading dimensions that are multiples of 8");
  MLG_CHECK_ARG(lda >= K && ldb >= K && ldc >= N, "mlg_gemm_bf16_ws: leading dimension too small");
  CUtensorMap ma, mb;
  int rc = make_map(&ma, A, K, M, lda, batch, stride_a, BM);
  if (rc) return rc;
  rc = make_map(&mb, B, K, N, ldb, batch, stride_b, 128);
  if (rc) return rc;
  return launch_gemm_streamk<6>(ma, mb, C, ldc, stride_c, (int)M, (int)N, (int)K, (int)batch, alpha, workspace,
                                (cudaStream_t)stream);
}
