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
//
// Large graphs (N >= 4096, k*dilation <= 16, D <= 128) take the tensor-core path at the end of this file: candidates by
// 3xTF32 distance products on tcgen05, then an fp32 re-evaluation with THIS kernel's arithmetic, a certificate per row, and
// this kernel as the repair pass for blocks without one -- same output, bit for bit.
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
           long long* __restrict__ out_nbr, long long* __restrict__ out_ctr, float* __restrict__ out_dist,
           const int* __restrict__ block_flags) {
  extern __shared__ __align__(16) float smem[];
  // repair pass behind the tensor-core path: only the query blocks it could not certify are recomputed
  if (block_flags && block_flags[blockIdx.y * gridDim.x + blockIdx.x] == 0) return;
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

// ---------------------------------------------------------------------------------------------------------------------
// Tensor-core path (N >= kTcMinN, k*dilation <= kTcMaxK, D <= 128).  The distance products of a graph are computed on
// the tcgen05 tensor cores with the 3xTF32 split (csrc/gemm_tf32x3.cu, kNN epilogue): every row keeps its kTcCand best
// candidates by the APPROXIMATE key.  The exactness contract above is then restored in fp32:
//   * knn_refine_kernel re-evaluates the kTcCand candidates of a row with exactly the arithmetic of knn_kernel (same fma
//     chain, same association) and orders them (ties: lowest index);
//   * the result is certified when the exact k*dilation-th distance is smaller than the worst kept approximate distance
//     minus a bound on |approximate - fp32-chain| (then no discarded point can belong to the answer, ties included);
//   * rows that cannot be certified (duplicated points, dense ties) flag their 128-row block and knn_kernel recomputes
//     those blocks exactly.  The output is therefore identical to the fp32 kernel's, always.
constexpr int kTcMinN = 4096, kTcMaxK = 16, kTcCand = 24;

// Xp [rows, Kp] = x zero-padded to Kp (the A operand: the tensor core truncates it to its tf32 "hi" part), hi / lo = the
// explicit split for the B operand, sq = the fma-chain squared norm, sqmax = max over all rows (bit pattern, sq >= 0)
__global__ void knn_tc_prepare_kernel(const float* __restrict__ x, long long rows, int D, int Kp, float* __restrict__ xp,
                                      float* __restrict__ hi, float* __restrict__ lo, float* __restrict__ sq,
                                      int* __restrict__ sqmax) {
  const long long r = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* p = x + (size_t)r * D;
  for (int d = lane; d < Kp; d += 32) {
    const float v = d < D ? __ldg(p + d) : 0.f;
    const float h = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
    xp[(size_t)r * Kp + d] = v;
    hi[(size_t)r * Kp + d] = h;
    lo[(size_t)r * Kp + d] = v - h;
  }
  if (lane == 0) {
    float s = 0.f;
    for (int d = 0; d < D; ++d) s = fmaf(p[d], p[d], s);   // same chain as the tile dot product of knn_kernel
    sq[r] = s;
    atomicMax(sqmax, __float_as_int(s));
  }
}

__global__ void knn_tc_init_lists_kernel(float* __restrict__ d, int* __restrict__ i, long long n) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t < n) { d[t] = INFINITY; i[t] = -1; }
}

// one warp per query row; lane e owns candidate e
__global__ void __launch_bounds__(256)
knn_refine_kernel(const float* __restrict__ x, const float* __restrict__ sq, const int* __restrict__ sqmax,
                  const float* __restrict__ list_d, const int* __restrict__ list_i, int B, int N, int D, int K, int dil,
                  int add_offset, long long* __restrict__ out_nbr, long long* __restrict__ out_ctr,
                  float* __restrict__ out_dist, int* __restrict__ block_flags, int blocks_per_graph) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= (long long)B * N) return;
  const int b = (int)(row / N), qi = (int)(row - (long long)b * N);
  const float* xb = x + (size_t)b * N * D;
  const float* sqb = sq + (size_t)b * N;
  const int j = lane < kTcCand ? list_i[row * kTcCand + lane] : -1;
  float dist = INFINITY;
  if (j >= 0) {
    const float* xi = xb + (size_t)qi * D;
    const float* xj = xb + (size_t)j * D;
    float acc = 0.f;
    for (int d = 0; d < D; ++d) acc = fmaf(__ldg(xi + d), __ldg(xj + d), acc);
    dist = (__ldg(sqb + qi) + (-2.f * acc)) + __ldg(sqb + j);
  }
  const int jj = j >= 0 ? j : 0x7fffffff;
  int rank = 0;
  for (int f = 0; f < kTcCand; ++f) {
    const float df = __shfl_sync(0xffffffffu, dist, f);
    const int jf = __shfl_sync(0xffffffffu, jj, f);
    rank += (df < dist || (df == dist && jf < jj)) ? 1 : 0;
  }
  // certification: exact K-th distance against the worst kept approximate distance
  const float sqi = __ldg(sqb + qi);
  float wkey = lane < kTcCand ? list_d[row * kTcCand + lane] : -INFINITY;   // the lists are unordered: largest kept key
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) wkey = fmaxf(wkey, __shfl_xor_sync(0xffffffffu, wkey, o));
  const float worst = wkey + sqi;
  const float eps = (sqi + __int_as_float(*sqmax)) * 1.9073486e-6f;   // 2^-19 (|x_i|^2 + max |x_j|^2): 8x the error model
  const unsigned kth = __ballot_sync(0xffffffffu, j >= 0 && rank == K - 1);
  bool safe = kth != 0;
  if (safe) {
    const float dk = __shfl_sync(0xffffffffu, dist, __ffs(kth) - 1);
    safe = dk < worst - eps;
  }
  if (!safe && lane == 0) block_flags[b * blocks_per_graph + qi / 128] = 1;
  const int k_out = K / dil;
  if (j >= 0 && rank < K && rank % dil == 0) {
    const long long off = add_offset ? (long long)b * N : 0;
    const size_t o = ((size_t)b * N + qi) * k_out + rank / dil;
    out_nbr[o] = (long long)j + off;
    out_ctr[o] = (long long)qi + off;
    if (out_dist) out_dist[o] = dist;
  }
}

inline int64_t align256(int64_t v) { return (v + 255) & ~(int64_t)255; }
inline int tc_kp(int64_t D) { return (int)((D + 31) / 32 * 32); }
inline bool tc_eligible(int64_t N, int64_t D, int64_t K) { return N >= kTcMinN && K <= kTcMaxK && tc_kp(D) <= 128; }

}  // namespace

extern "C" int64_t mlg_knn_workspace_bytes(int64_t B, int64_t N, int64_t D, int64_t k, int64_t dilation) {
  if (B < 1 || N < 1 || D < 1 || k < 1 || dilation < 1) return -1;
  const int64_t rows = B * N;
  int64_t bytes = align256(rows * 4);                                     // squared norms (both paths)
  if (tc_eligible(N, D, k * dilation)) {
    const int64_t kp = tc_kp(D);
    bytes += 256 + align256(B * ((N + 127) / 128) * 4)                    // sqmax, block flags
             + 3 * align256(rows * kp * 4)                                // padded x, hi, lo
             + 2 * align256(rows * kTcCand * 4);                          // candidate lists
  }
  return bytes;
}

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
  const int* flags = nullptr;
  if (tc_eligible(N, D, K) && workspace_bytes >= mlg_knn_workspace_bytes(B, N, D, k, dilation)) {
    // ---- tensor-core candidate search + fp32 refinement (see above); the exact kernel below repairs flagged blocks ----
    const int64_t rows = B * N;
    const int kp = tc_kp(D);
    const int nc = kp <= 64 ? 256 : 128;                     // candidate columns per launch (split weights must fit)
    const int bpg = (int)((N + 127) / 128);
    char* w = (char*)workspace + align256(rows * 4);
    int* sqmax = (int*)w;                       w += 256;
    int* bflags = (int*)w;                      w += align256((int64_t)B * bpg * 4);
    float* xp = (float*)w;                      w += align256(rows * kp * 4);
    float* xhi = (float*)w;                     w += align256(rows * kp * 4);
    float* xlo = (float*)w;                     w += align256(rows * kp * 4);
    float* list_d = (float*)w;                  w += align256(rows * kTcCand * 4);
    int* list_i = (int*)w;
    MLG_CUDA(cudaMemsetAsync(sqmax, 0, 256 + align256((int64_t)B * bpg * 4), st));
    knn_tc_prepare_kernel<<<mlg_ceil_div(rows, 8), 256, 0, st>>>(x, rows, (int)D, kp, xp, xhi, xlo, sq, sqmax);
    MLG_CHECK_LAUNCH("mlg_knn_graph(prepare)");
    knn_tc_init_lists_kernel<<<mlg_ceil_div(rows * kTcCand, 256), 256, 0, st>>>(list_d, list_i, rows * kTcCand);
    MLG_CHECK_LAUNCH("mlg_knn_graph(lists)");
    for (int64_t b = 0; b < B; ++b) {
      for (int64_t c0 = 0; c0 < N; c0 += nc) {
        const int64_t valid = N - c0 < nc ? N - c0 : nc;
        const size_t ro = (size_t)(b * N + c0) * kp;
        int rc = mlg_tf32x3_knn_chunk(xp + (size_t)b * N * kp, xhi + ro, xlo + ro, sq + b * N + c0, N, nc, valid, kp, kTcCand,
                                      list_d + (size_t)b * N * kTcCand, list_i + (size_t)b * N * kTcCand, (int)c0, stream);
        if (rc) return rc;
      }
    }
    knn_refine_kernel<<<mlg_ceil_div(rows, 8), 256, 0, st>>>(x, sq, sqmax, list_d, list_i, (int)B, (int)N, (int)D, (int)K,
                                                            (int)dilation, add_offset, (long long*)out_nbr,
                                                            (long long*)out_ctr, out_dist, bflags, bpg);
    MLG_CHECK_LAUNCH("mlg_knn_graph(refine)");
    flags = bflags;
  } else {
    sqnorm_kernel<<<mlg_ceil_div(B * N, 256), 256, 0, st>>>(x, B * N, (int)D, sq);
    MLG_CHECK_LAUNCH("mlg_knn_graph(sqnorm)");
  }
  // 128-row query tiles unless the grid would be tiny, then 64-row tiles (more, smaller blocks)
  const bool big = flags != nullptr || mlg_ceil_div(N, 128) * B >= 64;   // (the repair pass works on 128-row blocks)
  //   // measured at N=10k, D=1024: 79 big blocks 16.6 TF vs 157 small 10.0 TF
  const int tq = big ? 128 : 64;
  const size_t smem = sizeof(float) * (2 * DK * LDT + tq * (QW + 1) + TC) + (size_t)tq * K * 8;
  dim3 grid(mlg_ceil_div(N, tq), (unsigned)B);
  if (big) {
    MLG_CUDA(cudaFuncSetAttribute(knn_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<128><<<grid, kThreads, smem, st>>>(x, sq, (int)N, (int)D, (int)K, (int)dilation, add_offset,
                                                 (long long*)out_nbr, (long long*)out_ctr, out_dist, flags);
  } else {
    MLG_CUDA(cudaFuncSetAttribute(knn_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    knn_kernel<64><<<grid, kThreads, smem, st>>>(x, sq, (int)N, (int)D, (int)K, (int)dilation, add_offset,
                                                (long long*)out_nbr, (long long*)out_ctr, out_dist, nullptr);
  }
  MLG_CHECK_LAUNCH("mlg_knn_graph");
  return MLG_OK;
}
