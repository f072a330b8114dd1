// Dilated kNN graph construction: tiled fp32 distance + warp-level top-k (sm_100a).
//
// Replaces pairwise_distance + topk + index bookkeeping of
//   models/gcn_lib/sparse/torch_edge.py:53-104   (knn_matrix / knn_graph_matrix / Dilated)
//   models/gcn_lib/dense/torch_edge.py:32-58     (dense_knn_matrix / DenseDilated)
// The reference materialises the [B,N,N] distance matrix (400 MB per graph at N=10k) and runs a full-row
// radix select; here a 128x128 distance tile lives in registers (8x8 per thread) and passes through shared
// memory 32 columns at a time for the selection; each query row keeps a sorted K-entry candidate list in
// shared memory.
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

constexpr int TC = 128;  // candidate columns per tile
constexpr int DK = 16;   // feature chunk
constexpr int LDT = 128 + 4;  // padded row length of the transposed operand tiles
constexpr int QW = 32;   // columns per selection pass
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

// 128 x 128 distance tile per iteration: 16 x 16 threads, 8 x 8 outputs each (rows {ty*4..+3, 64+ty*4..+3},
// cols {tx*4..+3, 64+tx*4..+3}: every shared-memory operand read is a conflict-free / broadcast LDS.128)
template <int TQ>   // query rows per block: 128 (8x8 outputs per thread) or 64 (4x8; more blocks for small graphs)
__global__ void __launch_bounds__(kThreads)
knn_kernel(const float* __restrict__ x, const float* __restrict__ sq, int N, int D, int K, int dil, int add_offset,
           long long* __restrict__ out_nbr, long long* __restrict__ out_ctr, float* __restrict__ out_dist) {
  extern __shared__ __align__(16) float smem[];
  float* As = smem;                        // [DK][LDT]
  float* Bs = As + DK * LDT;               // [DK][LDT]
  float* Ds = Bs + DK * LDT;               // [TQ][QW+1]
  float* sqc = Ds + TQ * (QW + 1);         // [TC]
  float* bd = sqc + TC;                    // [TQ][K]
  int* bi = reinterpret_cast<int*>(bd + TQ * K);  // [TQ][K]

  const int b = blockIdx.y;
  const int q0 = blockIdx.x * TQ;
  const float* xb = x + (size_t)b * N * D;
  const float* sqb = sq + (size_t)b * N;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int ty = tid / 16, tx = tid % 16;

  for (int i = tid; i < TQ * K; i += kThreads) { bd[i] = INFINITY; bi[i] = -1; }
  constexpr int RT = TQ / 16;   // rows per thread
  float sqq[RT];
#pragma unroll
  for (int r = 0; r < RT; ++r) {
    const int qi = q0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + r - 4);
    sqq[r] = qi < N ? __ldg(sqb + qi) : 0.f;
  }
  __syncthreads();

  for (int c0 = 0; c0 < N; c0 += TC) {
    float acc[RT][8];
#pragma unroll
    for (int r = 0; r < RT; ++r)
#pragma unroll
      for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;
    if (tid < TC) sqc[tid] = (c0 + tid < N) ? __ldg(sqb + c0 + tid) : 0.f;

    for (int d0 = 0; d0 < D; d0 += DK) {
      // 128 rows x 16 features of the query block and of the candidate block, transposed into [k][row]
      for (int i = tid; i < TC * DK; i += kThreads) {
        const int row = i / DK, k = i % DK;
        const int qi = q0 + row, ci = c0 + row, dd = d0 + k;
        if (row < TQ) As[k * LDT + row] = (qi < N && dd < D) ? __ldg(xb + (size_t)qi * D + dd) : 0.f;
        Bs[k * LDT + row] = (ci < N && dd < D) ? __ldg(xb + (size_t)ci * D + dd) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < DK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4*>(As + k * LDT + ty * 4);
        const float4 a1 = (RT == 8) ? *reinterpret_cast<const float4*>(As + k * LDT + 64 + ty * 4) : a0;
        const float4 b0 = *reinterpret_cast<const float4*>(Bs + k * LDT + tx * 4);
        const float4 b1 = *reinterpret_cast<const float4*>(Bs + k * LDT + 64 + tx * 4);
        const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
          for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(av[r], bv[c], acc[r][c]);
      }
      __syncthreads();
    }
    // selection in four 32-column passes: distances -> shared, then warp w scans rows 16w .. 16w+15
#pragma unroll
    for (int pass = 0; pass < TC / QW; ++pass) {
      const int half = pass >> 1;              // 0: columns tx*4.., 1: columns 64 + tx*4..
      const int txlo = (pass & 1) * 8;         // which 8 tx values own this pass's 32 columns
      if (tx >= txlo && tx < txlo + 8) {
#pragma unroll
        for (int r = 0; r < RT; ++r) {
          const int row = r < 4 ? ty * 4 + r : 64 + ty * 4 + r - 4;
#pragma unroll
          for (int c = 0; c < 4; ++c) {
            const int cl = (tx - txlo) * 4 + c;                  // column inside the pass
            const int cj = half * 64 + tx * 4 + c;               // column inside the tile
            const float dist = (sqq[r] + (-2.f * acc[r][half * 4 + c])) + sqc[cj];
            Ds[row * (QW + 1) + cl] = (c0 + cj < N) ? dist : INFINITY;
          }
        }
      }
      __syncthreads();
      const int colbase = c0 + half * 64 + txlo * 4;
      for (int rr = 0; rr < TQ / 8; ++rr) {
        const int row = wid * (TQ / 8) + rr;
        if (q0 + row >= N) break;
        float* rbd = bd + row * K;
        int* rbi = bi + row * K;
        const float dv = Ds[row * (QW + 1) + lane];
        float thr = rbd[K - 1];
        unsigned hit = __ballot_sync(0xffffffffu, dv < thr);
        while (hit) {
          const int src = __ffs(hit) - 1;
          hit &= hit - 1;
          const float dc = __shfl_sync(0xffffffffu, dv, src);
          if (dc < thr) {  // threshold may have tightened since the ballot
            list_insert(rbd, rbi, K, dc, colbase + src, lane);
            thr = rbd[K - 1];
          }
        }
      }
      __syncthreads();
    }
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
  // 128-row query tiles unless the grid would be tiny, then 64-row tiles (more, smaller blocks)
  const bool big = mlg_ceil_div(N, 128) * B >= 64;   // measured at N=10k, D=1024: 79 big blocks 16.6 TF vs 157 small 10.0 TF
  const int tq = big ? 128 : 64;
  const size_t smem = sizeof(float) * (2 * DK * LDT + tq * (QW + 1) + TC) + (size_t)tq * K * 8;
  dim3 grid(mlg_ceil_div(N, tq), (unsigned)B);
  if (big) {
    MLG_CUDA(cudaFuncSetAttribute(knn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<128><<<grid, kThreads, smem, st>>>(x, sq, (int)N, (int)D, (int)K, (int)dilation, add_offset,
                                                 (long long*)out_nbr, (long long*)out_ctr, out_dist);
  } else {
    MLG_CUDA(cudaFuncSetAttribute(knn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<64><<<grid, kThreads, smem, st>>>(x, sq, (int)N, (int)D, (int)K, (int)dilation, add_offset,
                                                (long long*)out_nbr, (long long*)out_ctr, out_dist);
  }
  MLG_CHECK_LAUNCH("mlg_knn_graph");
  return MLG_OK;
}
