// Weighted segment gather-sum over a CSR (SpMM with dense feature rows), sm_100a.
//
// Forward of the SAGE / RSAGE mean aggregation the three shipped configs use
//   SAGEConv.message + PyG mean aggr   models/gcn_lib/sparse/torch_vertex.py:279-286
// (the per-edge lin_r GEMM is hoisted AFTER the aggregation), its backward
// (same kernel on the by-source CSR) and the source-side pass of the GENConv backward.
// Deterministic: a row's entries are summed in CSR order by one lane group, no atomics.
// HBM-bound: algorithmic bytes = 4*C*n_rows*2 + 8*nnz.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int kThreads = 256;
constexpr int UN = 4;

struct GsP {
  const float* src;
  const int* rowptr;
  const int* idx;
  const float* val;
  const float* pre;
  const float* post;
  int n, C, src_mod, post_mode, relative, accumulate;
  float* out;
};

template <int LANES, int VEC>
__global__ void __launch_bounds__(kThreads) gather_sum_kernel(GsP P) {
  constexpr int RPW = 32 / LANES;
  constexpr int CW = LANES * VEC;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, sl = lane % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long row = warp * RPW + sub;
  if (row >= P.n) return;
  const int C = P.C;
  const int beg = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
  const int cnt_row = end - beg;
  float postf = 1.f;
  if (P.post_mode == 1) postf = cnt_row > 0 ? 1.f / (float)cnt_row : 0.f;
  else if (P.post_mode == 2) postf = __ldg(P.post + row);
  const int nchunks = (C + CW - 1) / CW;
  for (int ch = 0; ch < nchunks; ++ch) {
    const int c = ch * CW + sl * VEC;
    const bool cok = c < C;
    float acc[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] = 0.f;
    for (int base = beg; base < end; base += LANES) {
      const int q = min(base + sl, end - 1);
      const int my_idx = __ldg(P.idx + q);
      float my_w = P.val ? __ldg(P.val + q) : 1.f;
      if (P.pre) my_w *= __ldg(P.pre + my_idx);
      const int my_row = P.src_mod ? my_idx % P.src_mod : my_idx;
      const int cnt = min(LANES, end - base);
      for (int j = 0; j < cnt; j += UN) {
        float4 xv[UN];
        float w[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int jj = min(j + u, cnt - 1);
          const int s = __shfl_sync(gmask, my_row, jj, LANES);
          w[u] = __shfl_sync(gmask, my_w, jj, LANES);
          if (j + u >= cnt) w[u] = 0.f;
          const float* p = P.src + (size_t)s * C + c;
          if (VEC == 4) xv[u] = cok ? ld_gather4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
          else xv[u].x = cok ? __ldg(p) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if (j + u < cnt) {
            acc[0] = fmaf(w[u], xv[u].x, acc[0]);
            if (VEC == 4) {
              acc[1 % VEC] = fmaf(w[u], xv[u].y, acc[1 % VEC]);
              acc[2 % VEC] = fmaf(w[u], xv[u].z, acc[2 % VEC]);
              acc[3 % VEC] = fmaf(w[u], xv[u].w, acc[3 % VEC]);
            }
          }
        }
      }
    }
    if (!cok) continue;
#pragma unroll
    for (int k = 0; k < VEC; ++k) acc[k] *= postf;
    float* o = P.out + (size_t)row * C + c;
    if (P.relative) {
      const long long srow = P.src_mod ? row % P.src_mod : row;
      const float f = (P.post_mode == 1) ? (cnt_row > 0 ? 1.f : 0.f) : (float)cnt_row * postf;
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] -= f * __ldg(P.src + (size_t)srow * C + c + k);
    }
    if (P.accumulate) {
#pragma unroll
      for (int k = 0; k < VEC; ++k) acc[k] += o[k];
    }
    if (VEC == 4) st4(o, make_float4(acc[0], acc[1 % VEC], acc[2 % VEC], acc[3 % VEC]));
    else o[0] = acc[0];
  }
}

__global__ void edge_values_kernel(const float* __restrict__ ea, const int* __restrict__ eid,
                                   const int* __restrict__ rowptr, int n_rows, long long cap, float fill,
                                   float* __restrict__ val) {
  const long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (q >= cap) return;
  const int nnz = __ldg(rowptr + n_rows);
  float v = 0.f;
  if (q < nnz) {
    const int e = __ldg(eid + q);
    v = e >= 0 ? __ldg(ea + e) : fill;
  }
  val[q] = v;
}

__global__ void embed_scale_fwd_kernel(const float* __restrict__ xs, const float* __restrict__ emb,
                                       long long total4, int N, int C4, float4* __restrict__ out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const long long r = i / C4;
  const int c4 = (int)(i - r * C4);
  const int n = (int)(r % N);
  const float s = __ldg(xs + r);
  const float4 e = __ldg(reinterpret_cast<const float4*>(emb) + (size_t)n * C4 + c4);
  out[i] = make_float4(s * e.x, s * e.y, s * e.z, s * e.w);
}

__global__ void embed_scale_bwd_kernel(const float* __restrict__ xs, const float4* __restrict__ g, int B,
                                       int N, int C4, float4* __restrict__ g_emb) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)N * C4) return;
  const int n = (int)(i / C4);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int b = 0; b < B; ++b) {
    const float s = __ldg(xs + (size_t)b * N + n);
    const float4 v = __ldg(g + (size_t)b * N * C4 + i);
    a.x = fmaf(s, v.x, a.x);
    a.y = fmaf(s, v.y, a.y);
    a.z = fmaf(s, v.z, a.z);
    a.w = fmaf(s, v.w, a.w);
  }
  g_emb[i] = a;
}

}  // namespace

extern "C" int mlg_gather_sum(const float* src, const int32_t* rowptr, const int32_t* idx, const float* val,
                              const float* pre, const float* post, int64_t n_rows, int64_t C,
                              int64_t src_mod, int post_mode, int relative, int accumulate, float* out,
                              void* stream) {
  MLG_CHECK_ARG(src && rowptr && idx && out, "mlg_gather_sum: null src/rowptr/idx/out");
  MLG_CHECK_ARG(n_rows >= 0 && n_rows < (1ll << 31) && C > 0 && C < (1ll << 20),
                "mlg_gather_sum: bad sizes n_rows=%lld C=%lld", (long long)n_rows, (long long)C);
  MLG_CHECK_ARG(post_mode >= 0 && post_mode <= 2 && (post_mode != 2 || post), "mlg_gather_sum: bad post_mode");
  if (n_rows == 0) return MLG_OK;
  GsP P{src, rowptr, idx, val, pre, post, (int)n_rows, (int)C, (int)src_mod, post_mode, relative, accumulate, out};
  cudaStream_t st = (cudaStream_t)stream;
  const int wpb = kThreads / 32;
  if (C % 4 != 0) {
    gather_sum_kernel<32, 1><<<mlg_ceil_div(n_rows, wpb), kThreads, 0, st>>>(P);
  } else if (C <= 32) {
    gather_sum_kernel<8, 4><<<mlg_ceil_div(n_rows, wpb * 4), kThreads, 0, st>>>(P);
  } else if (C <= 64) {
    gather_sum_kernel<16, 4><<<mlg_ceil_div(n_rows, wpb * 2), kThreads, 0, st>>>(P);
  } else {
    gather_sum_kernel<32, 4><<<mlg_ceil_div(n_rows, wpb), kThreads, 0, st>>>(P);
  }
  MLG_CHECK_LAUNCH("mlg_gather_sum");
  return MLG_OK;
}

extern "C" int mlg_edge_values(const float* edge_attr, const int32_t* eid, const int32_t* rowptr,
                               int64_t n_rows, int64_t cap, float fill, float* val, void* stream) {
  MLG_CHECK_ARG(eid && rowptr && val, "mlg_edge_values: null eid/rowptr/val");
  MLG_CHECK_ARG(edge_attr || cap == 0, "mlg_edge_values: null edge_attr");
  if (cap == 0) return MLG_OK;
  edge_values_kernel<<<mlg_ceil_div(cap, 256), 256, 0, (cudaStream_t)stream>>>(edge_attr, eid, rowptr,
                                                                             (int)n_rows, cap, fill, val);
  MLG_CHECK_LAUNCH("mlg_edge_values");
  return MLG_OK;
}

extern "C" int mlg_embed_scale_fwd(const float* xs, const float* emb, int64_t B, int64_t N, int64_t C,
                                   float* out, void* stream) {
  MLG_CHECK_ARG(xs && emb && out, "mlg_embed_scale_fwd: null pointer");
  MLG_CHECK_ARG(C % 4 == 0 && C > 0, "mlg_embed_scale_fwd: C=%lld must be a multiple of 4", (long long)C);
  const long long total4 = B * N * (C / 4);
  if (total4 == 0) return MLG_OK;
  embed_scale_fwd_kernel<<<mlg_ceil_div(total4, 256), 256, 0, (cudaStream_t)stream>>>(
      xs, emb, total4, (int)N, (int)(C / 4), reinterpret_cast<float4*>(out));
  MLG_CHECK_LAUNCH("mlg_embed_scale_fwd");
  return MLG_OK;
}

extern "C" int mlg_embed_scale_bwd(const float* xs, const float* g_out, int64_t B, int64_t N, int64_t C,
                                   float* g_emb, void* stream) {
  MLG_CHECK_ARG(xs && g_out && g_emb, "mlg_embed_scale_bwd: null pointer");
  MLG_CHECK_ARG(C % 4 == 0 && C > 0, "mlg_embed_scale_bwd: C=%lld must be a multiple of 4", (long long)C);
  const long long total = N * (C / 4);
  if (total == 0) return MLG_OK;
  embed_scale_bwd_kernel<<<mlg_ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      xs, reinterpret_cast<const float4*>(g_out), (int)B, (int)N, (int)(C / 4),
      reinterpret_cast<float4*>(g_emb));
  MLG_CHECK_LAUNCH("mlg_embed_scale_bwd");
  return MLG_OK;
}
