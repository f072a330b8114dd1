// Fully factored first SAGE layer of MultilevelGNN, forward, one WARP per gene (sm_100a).
//
//     out[b, i, :] = LeakyReLU( x[b, i] E_self[i, :] + (1 / cnt_i) sum_{q in row i} val_q x[b, idx_q] E_nbr[idx_q, :] + bias )
// (models/multilevel_gnn.py:150-151 feeding SAGEConv, gcn_lib/sparse/torch_vertex.py:269-294; see include/mlg_b200.h).
//
// r01's kernel gave every (row, 8 replicas) pair to a 16-lane group: each group re-fetched the row's table entries for its
// replica slice, 47.7 M warp instructions and 82 us for a 126 MB output (ncu: issue slots 56 % busy, DRAM 12 %).  Here a
// warp owns ONE gene for ALL (up to 32) replicas at once: lane l keeps channels (2l, 2l+1) of the 32 replicas in 64
// accumulator registers; per CSR entry the warp loads the neighbour's table row ONCE (one coalesced 256-byte load) and the
// neighbour's 32 replica values ONCE (one coalesced 128-byte load from the transposed node-value matrix xs_t [n][B]),
// four entries' loads in flight.  The scalar pass then broadcasts the replica values by shuffle (32 SHFL + 64 FFMA per
// entry); with all 32 replica slots live and C == 64 the packed pass below does the same arithmetic on replica pairs
// (8 broadcast LDS.128 + 32 FFMA2 per entry).
// The epilogue writes 32 coalesced 256-byte rows and (optionally) the 64 sign bits per (gene, replica) the backward kernel
// uses instead of re-reading the activation.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int EU = 4;   // entries whose loads are in flight together

struct R1F {
  const float* xs_t;   // [n][B]
  const float* e_self;
  const float* e_nbr;
  const int* rowptr;
  const int* idx;
  const float* val;
  const int* order;
  const float* bias;
  float* out;
  unsigned long long* mbits;   // optional [n][B]
  unsigned ld_self, ld_nbr, ld_out;
  float slope;
  int n, B;
  int nm;              // output rows node-major: (replica b, node i) at i * B + b instead of b * n + i
};

template <int CPL>
__device__ __forceinline__ void ld_cpl(float (&v)[CPL], const float* p) {
  if (CPL == 2) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x;
    v[CPL - 1] = t.y;
  } else {
    v[0] = __ldg(p);
  }
}

// FULL: all 32 replica slots of the pass are live (B a multiple of 32): no per-replica bound checks in the unrolled loops
template <int CPL, bool FULL>
__device__ __forceinline__ void rank1_rows_pass(const R1F& P, unsigned row, int beg, int end, float inv, const float (&es)[CPL],
                                                const float (&bs)[CPL], int rb0, int nb, int lane) {
  const bool live = FULL || lane < nb;
  float acc[32][CPL];
#pragma unroll
  for (int b = 0; b < 32; ++b)
#pragma unroll
    for (int k = 0; k < CPL; ++k) acc[b][k] = 0.f;
  const float* xcol = P.xs_t + rb0 + (live ? lane : 0);
  for (int base = beg; base < end; base += 32) {
    const int q = min(base + lane, end - 1);
    const unsigned my_idx = (unsigned)__ldg(P.idx + q);
    const float my_w = P.val ? __ldg(P.val + q) : 1.f;
    const int cnt = min(32, end - base);
    for (int j = 0; j < cnt; j += EU) {
      float e[EU][CPL], xw[EU];
#pragma unroll
      for (int u = 0; u < EU; ++u) {
        const int jj = min(j + u, cnt - 1);
        const unsigned s = __shfl_sync(0xffffffffu, my_idx, jj);
        const float w = (j + u < cnt) ? __shfl_sync(0xffffffffu, my_w, jj) : 0.f;
        ld_cpl<CPL>(e[u], P.e_nbr + (size_t)s * P.ld_nbr + CPL * lane);
        xw[u] = live ? __ldg(xcol + (size_t)s * P.B) * w : 0.f;
      }
#pragma unroll
      for (int b = 0; b < 32; ++b) {
        if (!FULL && b >= nb) break;   // warp-uniform
#pragma unroll
        for (int u = 0; u < EU; ++u) {
          const float f = __shfl_sync(0xffffffffu, xw[u], b);
#pragma unroll
          for (int k = 0; k < CPL; ++k) acc[b][k] = fmaf(f, e[u][k], acc[b][k]);
        }
      }
    }
  }
  const float xself = live ? __ldg(xcol + (size_t)row * P.B) : 0.f;
  unsigned lo_w = 0u, hi_w = 0u;     // lane b keeps the sign words of replica rb0 + b; one coalesced 8-byte store per lane
#pragma unroll
  for (int b = 0; b < 32; ++b) {
    if (!FULL && b >= nb) break;   // warp-uniform
    const float xb = __shfl_sync(0xffffffffu, xself, b);
    float y[CPL];
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const float z = fmaf(es[k], xb, fmaf(acc[b][k], inv, bs[k]));
      y[k] = z > 0.f ? z : z * P.slope;
    }
    float* dst = P.out + (P.nm ? (size_t)row * P.B + rb0 + b : (size_t)(rb0 + b) * P.n + row) * P.ld_out + CPL * lane;
    if (CPL == 2) *reinterpret_cast<float2*>(dst) = make_float2(y[0], y[CPL - 1]);
    else dst[0] = y[0];
    if (CPL == 2 && P.mbits) {
      // sign bits of this (gene, replica): channel c -> bit 16 * (c % 4) + c / 4 (the layout mlg_sage_rank1_bwd_rows reads):
      // ballot bit l of the first / second channel of lane l = channel 2l / 2l + 1; even lanes feed the low word, odd lanes
      // the high word, at bit (l >> 1) resp. 16 + (l >> 1)
      const unsigned b0 = __ballot_sync(0xffffffffu, y[0] > 0.f), b1 = __ballot_sync(0xffffffffu, y[CPL - 1] > 0.f);
      if (lane == b) { lo_w = b0; hi_w = b1; }
    }
  }
  if (CPL == 2 && P.mbits && live) {
    // compress: even bits of b0 -> low word bits 0..15, even bits of b1 -> low word bits 16..31; odd bits -> high word
    auto even = [](unsigned v) {   // gather bits 0, 2, 4, ... into the low 16 bits
      v &= 0x55555555u;
      v = (v | (v >> 1)) & 0x33333333u;
      v = (v | (v >> 2)) & 0x0f0f0f0fu;
      v = (v | (v >> 4)) & 0x00ff00ffu;
      v = (v | (v >> 8)) & 0x0000ffffu;
      return v;
    };
    const unsigned lo = even(lo_w) | (even(hi_w) << 16);
    const unsigned hi = even(lo_w >> 1) | (even(hi_w >> 1) << 16);
    P.mbits[(size_t)row * P.B + rb0 + lane] = ((unsigned long long)hi << 32) | lo;
  }
}

// CPL == 2, all 32 replica slots live: the same arithmetic on replica PAIRS.  The neighbour's 32 replica values (already
// scaled by the edge weight, as above) go through a warp-private shared-memory row and come back as eight broadcast
// LDS.128 -- every lane then holds all 32 values as 16 register pairs -- and the accumulators are pairs (replica 2j,
// replica 2j + 1) of one channel, so an entry costs 8 LDS.128 + 32 FFMA2 instead of 32 SHFL + 64 FFMA (the kernel is
// issue bound: 35.6 M warp instructions for a 126 MB output).  Per accumulator the FMAs and their operands are those of
// the scalar pass, in the same order: bit-identical results.
__device__ __forceinline__ void rank1_rows_pass_packed(const R1F& P, unsigned row, int beg, int end, float inv,
                                                       const float (&es)[2], const float (&bs)[2], int rb0, int lane,
                                                       float* xsm /* warp-private [EU][32], 16-byte aligned */) {
  u64 acc[2][16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[0][j] = acc[1][j] = 0ull;
  const float* xcol = P.xs_t + rb0 + lane;
  for (int base = beg; base < end; base += 32) {
    const int q = min(base + lane, end - 1);
    const unsigned my_idx = (unsigned)__ldg(P.idx + q);
    const float my_w = P.val ? __ldg(P.val + q) : 1.f;
    const int cnt = min(32, end - base);
    for (int j = 0; j < cnt; j += EU) {
      float e[EU][2], xw[EU];
#pragma unroll
      for (int u = 0; u < EU; ++u) {
        const int jj = min(j + u, cnt - 1);
        const unsigned s = __shfl_sync(0xffffffffu, my_idx, jj);
        const float w = (j + u < cnt) ? __shfl_sync(0xffffffffu, my_w, jj) : 0.f;
        ld_cpl<2>(e[u], P.e_nbr + (size_t)s * P.ld_nbr + 2 * lane);
        xw[u] = __ldg(xcol + (size_t)s * P.B) * w;
      }
#pragma unroll
      for (int u = 0; u < EU; ++u) xsm[u * 32 + lane] = xw[u];
      __syncwarp();
#pragma unroll
      for (int u = 0; u < EU; ++u) {
        const u64 e0 = pk2(e[u][0], e[u][0]), e1 = pk2(e[u][1], e[u][1]);
        const ulonglong2* xp = reinterpret_cast<const ulonglong2*>(xsm + u * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const ulonglong2 v = xp[i];   // replicas 4i .. 4i+3
          acc[0][2 * i] = fma2(v.x, e0, acc[0][2 * i]);
          acc[0][2 * i + 1] = fma2(v.y, e0, acc[0][2 * i + 1]);
          acc[1][2 * i] = fma2(v.x, e1, acc[1][2 * i]);
          acc[1][2 * i + 1] = fma2(v.y, e1, acc[1][2 * i + 1]);
        }
      }
      __syncwarp();   // the next batch overwrites the row
    }
  }
  xsm[lane] = __ldg(xcol + (size_t)row * P.B);
  __syncwarp();
  const u64 inv2 = pk2(inv, inv), es0 = pk2(es[0], es[0]), es1 = pk2(es[1], es[1]), bs0 = pk2(bs[0], bs[0]),
            bs1 = pk2(bs[1], bs[1]);
  unsigned lo_w = 0u, hi_w = 0u;     // lane b keeps the sign words of replica rb0 + b; one coalesced 8-byte store per lane
  const ulonglong2* xp = reinterpret_cast<const ulonglong2*>(xsm);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const ulonglong2 v = xp[i];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int j = 2 * i + h;
      const u64 xb = h ? v.y : v.x;
      float z[2][2];   // [replica of the pair][channel]
      upk2(fma2(es0, xb, fma2(acc[0][j], inv2, bs0)), z[0][0], z[1][0]);
      upk2(fma2(es1, xb, fma2(acc[1][j], inv2, bs1)), z[0][1], z[1][1]);
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int b = 2 * j + r;
        const float y0 = z[r][0] > 0.f ? z[r][0] : z[r][0] * P.slope;
        const float y1 = z[r][1] > 0.f ? z[r][1] : z[r][1] * P.slope;
        float* dst = P.out + (P.nm ? (size_t)row * P.B + rb0 + b : (size_t)(rb0 + b) * P.n + row) * P.ld_out + 2 * lane;
        *reinterpret_cast<float2*>(dst) = make_float2(y0, y1);
        if (P.mbits) {   // same bit layout as the scalar pass
          const unsigned b0 = __ballot_sync(0xffffffffu, y0 > 0.f), b1 = __ballot_sync(0xffffffffu, y1 > 0.f);
          if (lane == b) { lo_w = b0; hi_w = b1; }
        }
      }
    }
  }
  if (P.mbits) {
    auto even = [](unsigned v) {   // gather bits 0, 2, 4, ... into the low 16 bits
      v &= 0x55555555u;
      v = (v | (v >> 1)) & 0x33333333u;
      v = (v | (v >> 2)) & 0x0f0f0f0fu;
      v = (v | (v >> 4)) & 0x00ff00ffu;
      v = (v | (v >> 8)) & 0x0000ffffu;
      return v;
    };
    const unsigned lo = even(lo_w) | (even(hi_w) << 16);
    const unsigned hi = even(lo_w >> 1) | (even(hi_w >> 1) << 16);
    P.mbits[(size_t)row * P.B + rb0 + lane] = ((unsigned long long)hi << 32) | lo;
  }
  __syncwarp();
}

template <int CPL, bool PACKED>
__global__ void __launch_bounds__(kThreads, 2) sage_rank1_fwd_rows_kernel(const R1F P) {
  __shared__ __align__(16) float xsm[kThreads / 32][EU][32];
  const int lane = threadIdx.x & 31;
  const long long slot = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (slot >= P.n) return;
  const unsigned row = P.order ? (unsigned)__ldg(P.order + slot) : (unsigned)slot;
  const int beg = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
  const float inv = end > beg ? 1.f / (float)(end - beg) : 0.f;
  float es[CPL], bs[CPL];
  ld_cpl<CPL>(es, P.e_self + (size_t)row * P.ld_self + CPL * lane);
#pragma unroll
  for (int k = 0; k < CPL; ++k) bs[k] = P.bias ? __ldg(P.bias + CPL * lane + k) : 0.f;
  for (int rb0 = 0; rb0 < P.B; rb0 += 32) {
    const int nb = min(32, P.B - rb0);
    if (nb == 32) {
      if constexpr (CPL == 2 && PACKED) {
        rank1_rows_pass_packed(P, row, beg, end, inv, es, bs, rb0, lane, xsm[threadIdx.x >> 5][0]);
      } else {
        rank1_rows_pass<CPL, true>(P, row, beg, end, inv, es, bs, rb0, nb, lane);
      }
    } else {
      rank1_rows_pass<CPL, false>(P, row, beg, end, inv, es, bs, rb0, nb, lane);
    }
  }
}

__global__ void transpose_bn_kernel(const float* __restrict__ xs, int B, int n, float* __restrict__ xs_t) {
  __shared__ float tile[32][33];
  const int n0 = blockIdx.x * 32, b0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int b = b0 + r, i = n0 + threadIdx.x;
    tile[r][threadIdx.x] = (b < B && i < n) ? __ldg(xs + (size_t)b * n + i) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int i = n0 + r, b = b0 + threadIdx.x;
    if (i < n && b < B) xs_t[(size_t)i * B + b] = tile[threadIdx.x][r];
  }
}

}  // namespace

extern "C" int mlg_transpose_bn(const float* xs, int64_t B, int64_t n, float* xs_t, void* stream) {
  MLG_CHECK_ARG(xs && xs_t && B >= 1 && n >= 1, "mlg_transpose_bn: bad arguments");
  dim3 grid((unsigned)mlg_ceil_div(n, 32), (unsigned)mlg_ceil_div(B, 32)), block(32, 8);
  transpose_bn_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(xs, (int)B, (int)n, xs_t);
  MLG_CHECK_LAUNCH("mlg_transpose_bn");
  return MLG_OK;
}

extern "C" int mlg_sage_rank1_fwd_rows_supported(int64_t C) { return C == 32 || C == 64; }

static int rank1_fwd_rows_impl(const float* xs_t, const float* e_self, int64_t ld_self, const float* e_nbr, int64_t ld_nbr,
                               const int32_t* rowptr, const int32_t* idx, const float* val, const int32_t* order,
                               int64_t n_rows, int64_t C, int64_t replicas, const float* bias, float slope, float* out,
                               int64_t ld_out, uint64_t* mask_bits, int out_node_major, void* stream) {
  MLG_CHECK_ARG(xs_t && e_self && e_nbr && rowptr && idx && out, "mlg_sage_rank1_fwd_rows: null pointer");
  MLG_CHECK_ARG(mlg_sage_rank1_fwd_rows_supported(C), "mlg_sage_rank1_fwd_rows: C=%lld (needs 32 or 64)", (long long)C);
  MLG_CHECK_ARG(n_rows >= 0 && replicas >= 1 && replicas * n_rows < (1ll << 31) && ld_self >= C && ld_nbr >= C && ld_out >= C,
                "mlg_sage_rank1_fwd_rows: bad sizes");
  MLG_CHECK_ARG(ld_self % 2 == 0 && ld_nbr % 2 == 0 && ld_out % 2 == 0 &&
                    ((uintptr_t)e_self | (uintptr_t)e_nbr | (uintptr_t)out) % 8 == 0,
                "mlg_sage_rank1_fwd_rows: tables / out must be 8-byte aligned with even leading dimensions");
  MLG_CHECK_ARG(!mask_bits || C == 64, "mlg_sage_rank1_fwd_rows: mask_bits needs C == 64");
  if (n_rows == 0) return MLG_OK;
  R1F P;
  P.xs_t = xs_t; P.e_self = e_self; P.e_nbr = e_nbr; P.rowptr = rowptr; P.idx = idx; P.val = val; P.order = order;
  P.bias = bias; P.out = out; P.mbits = reinterpret_cast<unsigned long long*>(mask_bits);
  P.ld_self = (unsigned)ld_self; P.ld_nbr = (unsigned)ld_nbr; P.ld_out = (unsigned)ld_out; P.slope = slope;
  P.n = (int)n_rows; P.B = (int)replicas; P.nm = out_node_major;
  const unsigned grid = (unsigned)mlg_ceil_div(n_rows, kThreads / 32);
  cudaStream_t st = (cudaStream_t)stream;
  // the packed-pair pass needs 16-byte aligned rows of the transposed node values (B % 4 == 0, cudaMalloc'ed xs_t)
  const int packed = (replicas % 4 == 0 && (uintptr_t)xs_t % 16 == 0) ? 1 : 0;
  if (C == 64 && packed) sage_rank1_fwd_rows_kernel<2, true><<<grid, kThreads, 0, st>>>(P);
  else if (C == 64) sage_rank1_fwd_rows_kernel<2, false><<<grid, kThreads, 0, st>>>(P);
  else sage_rank1_fwd_rows_kernel<1, false><<<grid, kThreads, 0, st>>>(P);
  MLG_CHECK_LAUNCH("mlg_sage_rank1_fwd_rows");
  return MLG_OK;
}

extern "C" int mlg_sage_rank1_fwd_rows(const float* xs_t, const float* e_self, int64_t ld_self, const float* e_nbr, int64_t ld_nbr,
                                       const int32_t* rowptr, const int32_t* idx, const float* val, const int32_t* order,
                                       int64_t n_rows, int64_t C, int64_t replicas, const float* bias, float slope, float* out,
                                       int64_t ld_out, uint64_t* mask_bits, void* stream) {
  return rank1_fwd_rows_impl(xs_t, e_self, ld_self, e_nbr, ld_nbr, rowptr, idx, val, order, n_rows, C, replicas, bias, slope, out,
                             ld_out, mask_bits, 0, stream);
}

// the same layer with NODE-MAJOR output rows ((replica b, node i) at i * replicas + b: a gene's replica rows are one contiguous
// block, what the next layer's aggregation gathers per CSR entry)
extern "C" int mlg_sage_rank1_fwd_rows_nm(const float* xs_t, const float* e_self, int64_t ld_self, const float* e_nbr,
                                          int64_t ld_nbr, const int32_t* rowptr, const int32_t* idx, const float* val,
                                          const int32_t* order, int64_t n_rows, int64_t C, int64_t replicas, const float* bias,
                                          float slope, float* out, int64_t ld_out, uint64_t* mask_bits, void* stream) {
  return rank1_fwd_rows_impl(xs_t, e_self, ld_self, e_nbr, ld_nbr, rowptr, idx, val, order, n_rows, C, replicas, bias, slope, out,
                             ld_out, mask_bits, 1, stream);
}
