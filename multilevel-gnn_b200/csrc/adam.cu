// Adam on one flat parameter / gradient / state buffer (sm_100a) -- the optimizer step of train.py:112,66
// (torch.optim.Adam, amsgrad off): the Trainer keeps every trained parameter and its gradient as a view of one
// contiguous fp32 buffer (the gradient bucket of the data-parallel all-reduce), so the update is ONE element-wise
// pass instead of a multi-tensor launch.  The step counter lives on the device so the kernel can be replayed
// inside a CUDA graph.  HBM-bound: 28 bytes per parameter.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const float* __restrict__ step_dev, long long n, float lr, float b1, float b2, float eps, float wd) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float t = __ldg(step_dev) + 1.f;                 // this is step t (1-based)
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  float gi = g[i];
  const float pi = p[i];
  if (wd != 0.f) gi = fmaf(wd, pi, gi);
  const float mi = fmaf(b1, m[i], (1.f - b1) * gi);
  const float vi = fmaf(b2, v[i], (1.f - b2) * gi * gi);
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / sqrtf(bc2) + eps;      // torch: (sqrt(v) / sqrt(bias_correction2)) + eps
  p[i] = pi - (lr / bc1) * (mi / denom);
}

__global__ void adam_tick_kernel(float* step_dev) { *step_dev += 1.f; }

}  // namespace

extern "C" int mlg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step_dev,
                             int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                             void* stream) {
  MLG_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step_dev && n >= 0, "mlg_adam_step: bad arguments");
  if (n == 0) return MLG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  adam_kernel<<<mlg_ceil_div(n, 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, step_dev, n, lr, beta1, beta2, eps,
                                                   weight_decay);
  MLG_CHECK_LAUNCH("mlg_adam_step");
  adam_tick_kernel<<<1, 1, 0, st>>>(step_dev);
  MLG_CHECK_LAUNCH("mlg_adam_step(tick)");
  return MLG_OK;
}
