// Weight folding of the fused SAGE layer, one launch each way (sm_100a).
//
// SAGEConv (models/gcn_lib/sparse/torch_vertex.py:279-291) computes  nn( cat[x, mean_j lin_r(x_j)] ).  lin_r is linear,
// so it commutes with the mean and folds into the update weight:
//     Wcat = [ W1 | W2 . W_r ]        nn.weight = [W1 | W2]  ([cout, cin + r]),  W_r = lin_r.weight ([r, cin])
// and the layer becomes one GEMM over [x | mean_j x_j].  Forward needs Wcat (saved for backward), its 3xTF32 hi / lo
// split for the update GEMM and the split of Wcat^T for the dX GEMM; backward maps dWcat back to the two parameters:
//     d nn.weight = [ dWcat[:, :cin] | dWeff . W_r^T ],   d lin_r.weight = W2^T . dWeff,    dWeff = dWcat[:, cin:].
// These were 3 + 5 library launches (SIMT sgemm, cat, transpose copy, split) of a few microseconds each per layer.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

__global__ void sage_fold_fwd_kernel(const float* __restrict__ nn_w, const float* __restrict__ w_r, int cout, int cin,
                                     int r, float* __restrict__ wcat, float* __restrict__ hi, float* __restrict__ lo,
                                     float* __restrict__ t_hi, float* __restrict__ t_lo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k2 = 2 * cin, kin = cin + r;
  if (i >= cout * k2) return;
  const int o = i / k2, k = i % k2;
  float v;
  if (k < cin) {
    v = __ldg(nn_w + (size_t)o * kin + k);
  } else {
    v = 0.f;
    const float* w2 = nn_w + (size_t)o * kin + cin;
#pragma unroll 8
    for (int m = 0; m < r; ++m) v = fmaf(__ldg(w2 + m), __ldg(w_r + (size_t)m * cin + (k - cin)), v);
  }
  const float h = tf32_hi(v);
  wcat[i] = v;
  hi[i] = h;
  lo[i] = v - h;
  t_hi[(size_t)k * cout + o] = h;
  t_lo[(size_t)k * cout + o] = v - h;
}

__global__ void sage_fold_bwd_kernel(const float* __restrict__ g_wcat, const float* __restrict__ nn_w,
                                     const float* __restrict__ w_r, int cout, int cin, int r, float* __restrict__ g_nn,
                                     float* __restrict__ g_wr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k2 = 2 * cin, kin = cin + r;
  const int n_nn = cout * kin;
  if (i < n_nn) {
    const int o = i / kin, k = i % kin;
    float v;
    if (k < cin) {
      v = __ldg(g_wcat + (size_t)o * k2 + k);
    } else {   // (dWeff . W_r^T)[o, m] = sum_j dWeff[o, j] * W_r[m, j],  m = k - cin
      v = 0.f;
      const float* ge = g_wcat + (size_t)o * k2 + cin;
      const float* wr = w_r + (size_t)(k - cin) * cin;
#pragma unroll 8
      for (int j = 0; j < cin; ++j) v = fmaf(__ldg(ge + j), __ldg(wr + j), v);
    }
    g_nn[i] = v;
  } else if (i < n_nn + r * cin) {   // (W2^T . dWeff)[m, k] = sum_o W2[o, m] * dWeff[o, k]
    const int j = i - n_nn;
    const int m = j / cin, k = j % cin;
    float v = 0.f;
#pragma unroll 8
    for (int o = 0; o < cout; ++o)
      v = fmaf(__ldg(nn_w + (size_t)o * kin + cin + m), __ldg(g_wcat + (size_t)o * k2 + cin + k), v);
    g_wr[j] = v;
  }
}

// Stacked layout for the layers that transform first (and the factored first layer): Wst = [W1 ; W2 . W_r]  [2cout, cin]
// (U = x W1^T + b, V = x (W2 W_r)^T), written directly with everything the two GEMMs around it need -- the 3xTF32 split of
// Wst (update GEMM), the split of Wst^T [cin, 2cout] (dX GEMM) and the bias [b | 0] -- so that no permute copy, zero fill,
// bias copy, transpose copy or separate split launch sits between the fold and the GEMM (5 launches of 2 - 4 us each on
// the step's critical path).
__global__ void sage_fold_stacked_fwd_kernel(const float* __restrict__ nn_w, const float* __restrict__ w_r,
                                             const float* __restrict__ nn_b, int cout, int cin, int r,
                                             float* __restrict__ wst, float* __restrict__ hi, float* __restrict__ lo,
                                             float* __restrict__ t_hi, float* __restrict__ t_lo, float* __restrict__ bias2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int k2 = 2 * cin, kin = cin + r;
  if (bias2 && i < 2 * cout) bias2[i] = (i < cout && nn_b) ? __ldg(nn_b + i) : 0.f;
  if (i >= cout * k2) return;
  const int o = i / k2, k = i % k2;
  const int half = k >= cin, kk = half ? k - cin : k;
  float v;
  if (!half) {
    v = __ldg(nn_w + (size_t)o * kin + k);
  } else {
    v = 0.f;
    const float* w2 = nn_w + (size_t)o * kin + cin;
#pragma unroll 8
    for (int m = 0; m < r; ++m) v = fmaf(__ldg(w2 + m), __ldg(w_r + (size_t)m * cin + kk), v);
  }
  const float h = tf32_hi(v);
  const int row = half * cout + o;                 // row of Wst
  const size_t s = (size_t)row * cin + kk;
  wst[s] = v;
  hi[s] = h;
  lo[s] = v - h;
  const size_t t = (size_t)kk * (2 * cout) + row;  // Wst^T [cin, 2cout]
  t_hi[t] = h;
  t_lo[t] = v - h;
}

// backward from the stacked gradient g_Wst [2cout, cin] = ga (+ gb): rows [0, cout) are g_W1, rows [cout, 2cout) are
// g_(W2 W_r).  gb (NULL ok) is a second addend with the same leading dimension (the odd-row diagonal block of the row-pair
// weight-gradient product, see mlg_xty_tc).
__global__ void sage_fold_stacked_bwd_kernel(const float* __restrict__ ga, const float* __restrict__ gb, int ld,
                                             const float* __restrict__ nn_w, const float* __restrict__ w_r, int cout,
                                             int cin, int r, float* __restrict__ g_nn, float* __restrict__ g_wr) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int kin = cin + r;
  const int n_nn = cout * kin;
  auto g = [&](int row, int k) {
    const size_t a = (size_t)row * ld + k;
    return gb ? __ldg(ga + a) + __ldg(gb + a) : __ldg(ga + a);
  };
  if (i < n_nn) {
    const int o = i / kin, k = i % kin;
    float v;
    if (k < cin) {
      v = g(o, k);
    } else {   // (dWeff . W_r^T)[o, m] = sum_j dWeff[o, j] * W_r[m, j],  m = k - cin
      v = 0.f;
      const float* wr = w_r + (size_t)(k - cin) * cin;
#pragma unroll 8
      for (int j = 0; j < cin; ++j) v = fmaf(g(cout + o, j), __ldg(wr + j), v);
    }
    g_nn[i] = v;
  } else if (i < n_nn + r * cin) {   // (W2^T . dWeff)[m, k] = sum_o W2[o, m] * dWeff[o, k]
    const int j = i - n_nn;
    const int m = j / cin, k = j % cin;
    float v = 0.f;
#pragma unroll 8
    for (int o = 0; o < cout; ++o) v = fmaf(__ldg(nn_w + (size_t)o * kin + cin + m), g(cout + o, k), v);
    g_wr[j] = v;
  }
}

}  // namespace

extern "C" int mlg_sage_fold_fwd(const float* nn_w, const float* lin_r_w, int64_t cout, int64_t cin, int64_t r, float* wcat,
                                 float* wcat_hi, float* wcat_lo, float* wcat_t_hi, float* wcat_t_lo, void* stream) {
  MLG_CHECK_ARG(nn_w && lin_r_w && wcat && wcat_hi && wcat_lo && wcat_t_hi && wcat_t_lo, "mlg_sage_fold_fwd: null pointer");
  MLG_CHECK_ARG(cout >= 1 && cin >= 1 && r >= 1 && cout <= 4096 && cin <= 4096 && r <= 4096, "mlg_sage_fold_fwd: bad sizes");
  const long long n = cout * 2 * cin;
  sage_fold_fwd_kernel<<<(unsigned)mlg_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(
      nn_w, lin_r_w, (int)cout, (int)cin, (int)r, wcat, wcat_hi, wcat_lo, wcat_t_hi, wcat_t_lo);
  MLG_CHECK_LAUNCH("mlg_sage_fold_fwd");
  return MLG_OK;
}

extern "C" int mlg_sage_fold_bwd(const float* g_wcat, const float* nn_w, const float* lin_r_w, int64_t cout, int64_t cin,
                                 int64_t r,                                 float* g_nn_w, float* g_lin_r_w, void* stream) {
  MLG_CHECK_ARG(g_wcat && nn_w && lin_r_w && g_nn_w && g_lin_r_w, "mlg_sage_fold_bwd: null pointer");
  MLG_CHECK_ARG(cout >= 1 && cin >= 1 && r >= 1 && cout <= 4096 && cin <= 4096 && r <= 4096, "mlg_sage_fold_bwd: bad sizes");
  const long long n = cout * (cin + r) + r * cin;
  sage_fold_bwd_kernel<<<(unsigned)mlg_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(g_wcat, nn_w, lin_r_w, (int)cout,
                                                                                       (int)cin, (int)r, g_nn_w, g_lin_r_w);
  MLG_CHECK_LAUNCH("mlg_sage_fold_bwd");
  return MLG_OK;
}

extern "C" int mlg_sage_fold_stacked_fwd(const float* nn_w, const float* lin_r_w, const float* nn_b, int64_t cout, int64_t cin,
                                         int64_t r, float* wst, float* wst_hi, float* wst_lo, float* wst_t_hi, float* wst_t_lo,
                                         float* bias2, void* stream) {
  MLG_CHECK_ARG(nn_w && lin_r_w && wst && wst_hi && wst_lo && wst_t_hi && wst_t_lo, "mlg_sage_fold_stacked_fwd: null pointer");
  MLG_CHECK_ARG(cout >= 1 && cin >= 1 && r >= 1 && cout <= 4096 && cin <= 4096 && r <= 4096,
                "mlg_sage_fold_stacked_fwd: bad sizes");
  const long long n = cout * 2 * cin;   // >= 2 * cout: the bias threads are covered
  sage_fold_stacked_fwd_kernel<<<(unsigned)mlg_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(
      nn_w, lin_r_w, nn_b, (int)cout, (int)cin, (int)r, wst, wst_hi, wst_lo, wst_t_hi, wst_t_lo, bias2);
  MLG_CHECK_LAUNCH("mlg_sage_fold_stacked_fwd");
  return MLG_OK;
}

extern "C" int mlg_sage_fold_stacked_bwd(const float* g_wst, const float* g_wst_add, int64_t ld, const float* nn_w,
                                         const float* lin_r_w, int64_t cout, int64_t cin, int64_t r, float* g_nn_w,
                                         float* g_lin_r_w, void* stream) {
  MLG_CHECK_ARG(g_wst && nn_w && lin_r_w && g_nn_w && g_lin_r_w, "mlg_sage_fold_stacked_bwd: null pointer");
  MLG_CHECK_ARG(cout >= 1 && cin >= 1 && r >= 1 && cout <= 4096 && cin <= 4096 && r <= 4096 && ld >= cin,
                "mlg_sage_fold_stacked_bwd: bad sizes");
  const long long n = cout * (cin + r) + r * cin;
  sage_fold_stacked_bwd_kernel<<<(unsigned)mlg_ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(
      g_wst, g_wst_add, (int)ld, nn_w, lin_r_w, (int)cout, (int)cin, (int)r, g_nn_w, g_lin_r_w);
  MLG_CHECK_LAUNCH("mlg_sage_fold_stacked_bwd");
  return MLG_OK;
}
