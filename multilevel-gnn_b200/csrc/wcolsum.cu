// Weighted column sums of a tall matrix:  u[c] = sum_r a[r] * G[r, c],  v[c] = sum_r G[r, c]   (sm_100a).
//
// Backward of the rank-1 affine edge term e_ij = a_e * p + q of mlg_gen_aggr_*_affine: with G = d loss / d e ([E, H]) the
// gradients of the two H-vectors are g_p = u and g_q = v.  One streaming pass over G (HBM-bound, 4*E*H bytes): a warp owns
// a row at a time (lanes over the columns, float4), per-lane partial sums, block partials through shared memory, and a
// fixed-order reduction over the blocks (deterministic).  C % 4 == 0, C <= 1024.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int kWarps = 8;
constexpr int kMaxV = 8;   // float4 per lane: C <= 1024

__global__ void __launch_bounds__(kWarps * 32)
wcolsum_kernel(const float* __restrict__ G, unsigned ld, const float* __restrict__ a, long long rows, int C,
               float* __restrict__ partial /* [gridDim.x][2][C] */) {
  extern __shared__ float stage[];   // [kWarps][2][C]
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int nv = (C + 127) / 128;
  float4 su[kMaxV], sv[kMaxV];
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) su[i] = sv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  const long long stride = (long long)gridDim.x * kWarps;
  for (long long r = (long long)blockIdx.x * kWarps + wib; r < rows; r += stride) {
    const float w = a ? __ldg(a + r) : 0.f;
    const float* row = G + (size_t)r * ld + lane * 4;
#pragma unroll
    for (int i = 0; i < kMaxV; ++i) {
      if (i < nv && i * 128 + lane * 4 < C) {
        const float4 g = ld_stream4(row + i * 128);
        su[i].x = fmaf(w, g.x, su[i].x); su[i].y = fmaf(w, g.y, su[i].y);
        su[i].z = fmaf(w, g.z, su[i].z); su[i].w = fmaf(w, g.w, su[i].w);
        sv[i].x += g.x; sv[i].y += g.y; sv[i].z += g.z; sv[i].w += g.w;
      }
    }
  }
  float* mine = stage + (size_t)wib * 2 * C;
#pragma unroll
  for (int i = 0; i < kMaxV; ++i) {
    const int c = i * 128 + lane * 4;
    if (i < nv && c < C) {
      *reinterpret_cast<float4*>(mine + c) = su[i];
      *reinterpret_cast<float4*>(mine + C + c) = sv[i];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * C; j += kWarps * 32) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += stage[(size_t)w * 2 * C + j];
    partial[(size_t)blockIdx.x * 2 * C + j] = s;
  }
}

__global__ void __launch_bounds__(1024)
wcolsum_reduce_kernel(const float* __restrict__ partial, int n_part, int C2, float* __restrict__ u, float* __restrict__ v) {
  // blockDim.x / 32 warps share the partials of 32 columns (warp w: p = w, w + nw, ...; two chains each); the warp sums are
  // added in warp order.  With 8 warps and ~600 partials this was 37 dependent round trips on 4 blocks (9.5 us).
  __shared__ float sm[32][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  float s0 = 0.f, s1 = 0.f;
  if (j < C2) {
    int p = w;
    for (; p + nw < n_part; p += 2 * nw) {
      s0 += partial[(size_t)p * C2 + j];
      s1 += partial[(size_t)(p + nw) * C2 + j];
    }
    for (; p < n_part; p += nw) s0 += partial[(size_t)p * C2 + j];
  }
  sm[w][lane] = s0 + s1;
  __syncthreads();
  if (w == 0 && j < C2) {
    float s = 0.f;
    for (int k = 0; k < nw; ++k) s += sm[k][lane];
    const int C = C2 / 2;
    if (j < C) { if (u) u[j] = s; }
    else if (v) v[j - C] = s;
  }
}

// first level of a two-level reduction over many per-block partials: slice y adds partial rows [y*per, (y+1)*per)
__global__ void __launch_bounds__(256)
colsum_slices_kernel(const float* __restrict__ partial, int n_part, int per, int C2, float* __restrict__ out /* [slices][C2] */) {
  __shared__ float sm[8][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int j = blockIdx.x * 32 + lane;
  const int p0 = blockIdx.y * per, p1 = min(n_part, p0 + per);
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  if (j < C2) {
    int p = p0 + w;
    for (; p + 24 < p1; p += 32) {
      s0 += partial[(size_t)p * C2 + j];
      s1 += partial[(size_t)(p + 8) * C2 + j];
      s2 += partial[(size_t)(p + 16) * C2 + j];
      s3 += partial[(size_t)(p + 24) * C2 + j];
    }
    for (; p < p1; p += 8) s0 += partial[(size_t)p * C2 + j];
  }
  sm[w][lane] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (w == 0 && j < C2) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += sm[k][lane];
    out[(size_t)blockIdx.y * C2 + j] = s;
  }
}

inline int wc_blocks(long long rows) {
  long long b = (rows + kWarps - 1) / kWarps;
  const long long cap = 148 * 4;
  return (int)(b < cap ? (b < 1 ? 1 : b) : cap);
}

}  // namespace

// u[c] = sum_p partial[p][c], v[c] = sum_p partial[p][C + c] for per-block partials [n_part][2C] (the fused edge-term
// gradient of gen_bwd_kernel); scratch holds kPartSlices * 2C floats.  Fixed summation order.
constexpr int kPartSlices = 64;
int mlg_detail_colsum_partials(const float* partial, long long n_part, long long C, float* u, float* v, float* scratch,
                               cudaStream_t st) {
  const int C2 = (int)(2 * C);
  const unsigned gx = (unsigned)mlg_ceil_div(C2, 32);
  if (n_part > 4 * kPartSlices) {
    const int per = (int)mlg_ceil_div(n_part, kPartSlices);
    const int slices = (int)mlg_ceil_div(n_part, per);
    colsum_slices_kernel<<<dim3(gx, slices), 256, 0, st>>>(partial, (int)n_part, per, C2, scratch);
    MLG_CHECK_LAUNCH("colsum_slices");
    wcolsum_reduce_kernel<<<gx, 256, 0, st>>>(scratch, slices, C2, u, v);
  } else {
    wcolsum_reduce_kernel<<<gx, n_part > 64 ? 1024 : 256, 0, st>>>(partial, (int)n_part, C2, u, v);
  }
  MLG_CHECK_LAUNCH("colsum_partials");
  return MLG_OK;
}
long long mlg_detail_colsum_partials_scratch_floats(long long C) { return (long long)kPartSlices * 2 * C; }

extern "C" int64_t mlg_wcolsum_workspace_bytes(int64_t rows, int64_t C) { return (int64_t)wc_blocks(rows) * 2 * C * 4; }

extern "C" int mlg_wcolsum(const float* G, int64_t ld, const float* a, int64_t rows, int64_t C, float* u, float* v,
                           void* workspace, int64_t workspace_bytes, void* stream) {
  MLG_CHECK_ARG(G && workspace && (u || v), "mlg_wcolsum: null pointer");
  MLG_CHECK_ARG(rows >= 0 && C >= 4 && C % 4 == 0 && C <= 128 * kMaxV && ld >= C && ld % 4 == 0 && (uintptr_t)G % 16 == 0,
                "mlg_wcolsum: need C %% 4 == 0, C <= 1024, 16-byte aligned rows");
  MLG_CHECK_ARG(!u || a, "mlg_wcolsum: the weighted sum u needs the weights a");
  MLG_CHECK_ARG(workspace_bytes >= mlg_wcolsum_workspace_bytes(rows, C), "mlg_wcolsum: workspace too small");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = wc_blocks(rows);
  const size_t smem = (size_t)kWarps * 2 * C * 4;
  if (smem > 48 * 1024) cudaFuncSetAttribute(wcolsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  wcolsum_kernel<<<grid, kWarps * 32, smem, st>>>(G, (unsigned)ld, a, rows, (int)C, (float*)workspace);
  MLG_CHECK_LAUNCH("mlg_wcolsum");
  wcolsum_reduce_kernel<<<(unsigned)mlg_ceil_div(2 * C, 32), grid > 64 ? 1024 : 256, 0, st>>>((const float*)workspace, grid, (int)(2 * C), u, v);
  MLG_CHECK_LAUNCH("mlg_wcolsum(reduce)");
  return MLG_OK;
}
