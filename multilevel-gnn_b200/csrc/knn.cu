// Dilated kNN graph construction: tiled fp32 distance + warp-level top-k (sm_100a).
//
// Replaces pairwise_distance + topk + index bookkeeping of
//   models/gcn_lib/sparse/torch_edge.py:53-104   (knn_matrix / knn_graph_matrix / Dilated)
//   models/gcn_lib/dense/torch_edge.py:32-58     (dense_knn_matrix / DenseDilated)
// The reference materialises the [B,N,N] distance matrix (400 MB per graph at N=10k) and runs a full-row
// radix select; here a 64x64 distance tile lives in registers/shared memory only and each query row keeps
// a sorted K-entry candidate list in shared memory.
//
// Arithmetic (exactness contract, SURVEY.md section 7 "hard parts"): fp32 FMA only (no TF32/BF16),
//   dot_ij = fma-chain over d ascending;  sq_i = the same chain on (x_i, x_i)  => d_ii == 0 exactly;
//   dist_ij = (sq_i + (-2*dot_ij)) + sq_j          (association of torch_edge.py:61-63)
// Order: ascending distance, ties broken by the lowest index (candidates arrive in index order and an
// equal distance never displaces an earlier entry).
// Compute-bound on the fp32 FMA pipe: flops = 2*N^2*D per graph.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int TQ = 64;   // query rows per block
constexpr int TC = 64;   // candidate columns per tile
constexpr int DK = 16;   // feature chunk
constexpr int kThreads = 256;
constexpr int kMaxK = 128;

__global__ void sqnorm_kernel(const float* __restrict__ x, long long rows, int D, float* __restrict__ sq) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* p = x + (size_t)r * D;
  float s = 0.f;
  for (int d = 0; d < D; ++d) s = fmaf(p[d], p[d], s);  // same chain as the tile dot product
  sq[r] = s;
}

// warp-cooperative sorted insertion of (d, j) into a K-entry ascending list (ties keep earlier entries first)
__device__ __forceinline__ void list_insert(float* bd, int* bi, int K, float d, int j, int lane) {
  int pos = 0;
  for (int e = lane; e < K; e += 32) pos += (bd[e] <= d) ? 1 : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pos += __shfl_xor_sync(0xffffffffu, pos, o);
  if (pos >= K) return;
  // shift [pos, K-2] -> [pos+1, K-1], highest chunk first so no entry is overwritten before it is read
  for (int base = ((K - 1) / 32) * 32; base >= 0; base -= 32) {
    const int e = base + lane;
    float vd = 0.f;
    int vi = 0;
    const bool mv = e >= pos && e + 1 < K;
    if (mv) { vd = bd[e]; vi = bi[e]; }
    __syncwarp();
    if (mv) { bd[e + 1] = vd; bi[e + 1] = vi; }
    __syncwarp();
  }
  if (lane == 0) { bd[pos] = d; bi[pos] = j; }
  __syncwarp();
}

__global__ void __launch_bounds__(kThreads)
knn_kernel(const float* __restrict__ x, const float* __restrict__ sq, int N, int D, int K, int dil, int add_offset,
           long long* __restrict__ out_nbr, long long* __restrict__ out_ctr, float* __restrict__ out_dist) {
  extern __shared__ float smem[];
  float* As = smem;                        // [DK][TQ+4]
  float* Bs = As + DK * (TQ + 4);          // [DK][TC+4]
  float* Ds = Bs + DK * (TC + 4);          // [TQ][TC+1]
  float* sqc = Ds + TQ * (TC + 1);         // [TC]
  float* bd = sqc + TC;                    // [TQ][K]
  int* bi = reinterpret_cast<int*>(bd + TQ * K);  // [TQ][K]

  const int b = blockIdx.y;
  const int q0 = blockIdx.x * TQ;
  const float* xb = x + (size_t)b * N * D;
  const float* sqb = sq + (size_t)b * N;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ty = tid / 16, tx = tid % 16;  // 16x16 threads, 4x4 outputs each

  for (int i = tid; i < TQ * K; i += kThreads) { bd[i] = INFINITY; bi[i] = -1; }
  float sqq[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int qi = q0 + ty * 4 + r;
    sqq[r] = qi < N ? __ldg(sqb + qi) : 0.f;
  }
  __syncthreads();

  for (int c0 = 0; c0 < N; c0 += TC) {
    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;
    if (tid < TC) sqc[tid] = (c0 + tid < N) ? __ldg(sqb + c0 + tid) : 0.f;

    for (int d0 = 0; d0 < D; d0 += DK) {
      // load 64 x 16 query and candidate sub-tiles, transposed into [k][row]
      for (int i = tid; i < TQ * DK; i += kThreads) {
        const int row = i / DK, k = i % DK;
        const int qi = q0 + row, ci = c0 + row, dd = d0 + k;
        As[k * (TQ + 4) + row] = (qi < N && dd < D) ? __ldg(xb + (size_t)qi * D + dd) : 0.f;
        Bs[k * (TC + 4) + row] = (ci < N && dd < D) ? __ldg(xb + (size_t)ci * D + dd) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < DK; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(As + k * (TQ + 4) + ty * 4);
        const float4 bb = *reinterpret_cast<const float4*>(Bs + k * (TC + 4) + tx * 4);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {bb.x, bb.y, bb.z, bb.w};
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
      }
      __syncthreads();
    }
    // distances of this tile -> shared
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const int cj = tx * 4 + c;
        const float dist = (sqq[r] + (-2.f * acc[r][c])) + sqc[cj];
        Ds[(ty * 4 + r) * (TC + 1) + cj] = (c0 + cj < N) ? dist : INFINITY;
      }
    __syncthreads();
    // selection: warp w owns query rows w*8 .. w*8+7
    for (int rr = 0; rr < TQ / 8; ++rr) {
      const int row = wid * (TQ / 8) + rr;
      if (q0 + row >= N) break;
      float* rbd = bd + row * K;
      int* rbi = bi + row * K;
#pragma unroll
      for (int half = 0; half < TC / 32; ++half) {
        const float dv = Ds[row * (TC + 1) + half * 32 + lane];
        float thr = rbd[K - 1];
        unsigned pass = __ballot_sync(0xffffffffu, dv < thr);
        while (pass) {
          const int src = __ffs(pass) - 1;
          pass &= pass - 1;
          const float dc = __shfl_sync(0xffffffffu, dv, src);
          if (dc < thr) {  // threshold may have tightened since the ballot
            list_insert(rbd, rbi, K, dc, c0 + half * 32 + src, lane);
            thr = rbd[K - 1];
          }
        }
      }
    }
    __syncthreads();
  }

  // write every dil-th rank
  const int k_out = K / dil;
  for (int i = tid; i < TQ * k_out; i += kThreads) {
    const int row = i / k_out, r = i % k_out;
    const int qi = q0 + row;
    if (qi >= N) continue;
    const long long off = add_offset ? (long long)b * N : 0;
    const size_t o = ((size_t)b * N + qi) * k_out + r;
    out_nbr[o] = (long long)bi[row * K + r * dil] + off;
    out_ctr[o] = (long long)qi + off;
    if (out_dist) out_dist[o] = bd[row * K + r * dil];
  }
}

}  // namespace

extern "C" int mlg_knn_graph(const float* x, int64_t B, int64_t N, int64_t D, int64_t k, int64_t dilation,
                             int add_offset, int64_t* out_nbr, int64_t* out_ctr, float* out_dist,
                             void* workspace, int64_t workspace_bytes, void* stream) {
  MLG_CHECK_ARG(x && out_nbr && out_ctr, "mlg_knn_graph: null pointer");
  MLG_CHECK_ARG(B > 0 && N > 0 && D > 0 && k > 0 && dilation > 0, "mlg_knn_graph: non-positive size");
  const long long K = k * dilation;
  MLG_CHECK_ARG(K <= N, "mlg_knn_graph: k*dilation=%lld exceeds the %lld points of a graph (torch.topk raises too)", K,
                (long long)N);
  MLG_CHECK_ARG(K <= kMaxK, "mlg_knn_graph: k*dilation=%lld > %d not supported", K, kMaxK);
  MLG_CHECK_ARG(B * N < (1ll << 31) && B < 65536, "mlg_knn_graph: sizes exceed limits");
  MLG_CHECK_ARG(workspace && workspace_bytes >= (int64_t)(B * N * 4), "mlg_knn_graph: workspace needs B*N*4 bytes");
  cudaStream_t st = (cudaStream_t)stream;
  float* sq = (float*)workspace;
  sqnorm_kernel<<<mlg_ceil_div(B * N, 256), 256, 0, st>>>(x, B * N, (int)D, sq);
  MLG_CHECK_LAUNCH("mlg_knn_graph(sqnorm)");
  const size_t smem = sizeof(float) * (DK * (TQ + 4) + DK * (TC + 4) + TQ * (TC + 1) + TC) + (size_t)TQ * K * 8;
  MLG_CUDA(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid(mlg_ceil_div(N, TQ), (unsigned)B);
  knn_kernel<<<grid, kThreads, smem, st>>>(x, sq, (int)N, (int)D, (int)K, (int)dilation, add_offset,
                                          (long long*)out_nbr, (long long*)out_ctr, out_dist);
  MLG_CHECK_LAUNCH("mlg_knn_graph");
  return MLG_OK;
}
