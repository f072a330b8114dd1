// Small fused element-wise epilogues of the SAGE layer (sm_100a).
//
//   mlg_bias_act:      z[r, c] = act(z[r, c] + bias[c])     in place, act = LeakyReLU(slope) (slope 0 = ReLU)
// replaces the bias broadcast + activation that follow the update GEMM in SAGEConv.update / MLP
// (models/gcn_lib/sparse/torch_vertex.py:288-291, torch_nn.py:54-75): the library path spends one
// full-size kernel on the bias (cublasLt::globalKernel, 130 us at the gbm shape) and another on the
// activation.  HBM-bound: bytes = 8 * rows * C.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

__global__ void __launch_bounds__(256) bias_act_kernel(float4* __restrict__ z, const float4* __restrict__ bias,
                                                       long long total4, int c4n, float slope) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total4) return;
  float4 v = z[i];
  if (bias) {
    const float4 b = __ldg(bias + (int)(i % c4n));
    v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
  }
  v.x = v.x > 0.f ? v.x : v.x * slope;
  v.y = v.y > 0.f ? v.y : v.y * slope;
  v.z = v.z > 0.f ? v.z : v.z * slope;
  v.w = v.w > 0.f ? v.w : v.w * slope;
  z[i] = v;
}

__global__ void __launch_bounds__(256) bias_act_scalar_kernel(float* __restrict__ z, const float* __restrict__ bias,
                                                              long long total, int C, float slope) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = z[i] + (bias ? __ldg(bias + (int)(i % C)) : 0.f);
  z[i] = v > 0.f ? v : v * slope;
}

}  // namespace

extern "C" int mlg_bias_act(float* z, const float* bias, int64_t rows, int64_t C, float slope, void* stream) {
  MLG_CHECK_ARG(z, "mlg_bias_act: null z");
  MLG_CHECK_ARG(rows >= 0 && C > 0, "mlg_bias_act: bad sizes");
  const long long total = rows * C;
  if (total == 0) return MLG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  if (C % 4 == 0 && (uintptr_t)z % 16 == 0 && (!bias || (uintptr_t)bias % 16 == 0)) {
    bias_act_kernel<<<mlg_ceil_div(total / 4, 256), 256, 0, st>>>((float4*)z, (const float4*)bias, total / 4,
                                                                  (int)(C / 4), slope);
  } else {
    bias_act_scalar_kernel<<<mlg_ceil_div(total, 256), 256, 0, st>>>(z, bias, total, (int)C, slope);
  }
  MLG_CHECK_LAUNCH("mlg_bias_act");
  return MLG_OK;
}
