// Adam on one flat parameter / gradient / state buffer (sm_100a) -- the optimizer step of train.py:112,66
// (torch.optim.Adam, amsgrad off): the Trainer keeps every trained parameter and its gradient as a view of one
// contiguous fp32 buffer (the gradient bucket of the data-parallel all-reduce), so the update is ONE element-wise
// pass instead of a multi-tensor launch.  The step counter lives on the device so the kernel can be replayed
// inside a CUDA graph.  HBM-bound: 28 bytes per parameter.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps, float wd,
                                         float step_size, float inv_sqrt_bc2) {
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = fmaf(b1, m, (1.f - b1) * g);
  v = fmaf(b2, v, (1.f - b2) * g * g);
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;     // torch: (sqrt(v) / sqrt(bias_correction2)) + eps
  p -= step_size * (m / denom);
}

// 4 parameters per thread (128-bit accesses; the two powf of the bias corrections are shared by the 4)
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
            const float* __restrict__ step_dev, long long n, float lr, float b1, float b2, float eps, float wd) {
  const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) * 4;
  if (i >= n) return;
  const float t = __ldg(step_dev) + 1.f;                 // this is step t (1-based)
  const float bc1 = 1.f - powf(b1, t), bc2 = 1.f - powf(b2, t);
  const float step_size = lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  if (i + 4 <= n) {
    float4 pi = *reinterpret_cast<const float4*>(p + i), mi = *reinterpret_cast<const float4*>(m + i),
           vi = *reinterpret_cast<const float4*>(v + i);
    const float4 gi = *reinterpret_cast<const float4*>(g + i);
    adam_one(pi.x, gi.x, mi.x, vi.x, b1, b2, eps, wd, step_size, inv_sqrt_bc2);
    adam_one(pi.y, gi.y, mi.y, vi.y, b1, b2, eps, wd, step_size, inv_sqrt_bc2);
    adam_one(pi.z, gi.z, mi.z, vi.z, b1, b2, eps, wd, step_size, inv_sqrt_bc2);
    adam_one(pi.w, gi.w, mi.w, vi.w, b1, b2, eps, wd, step_size, inv_sqrt_bc2);
    *reinterpret_cast<float4*>(p + i) = pi;
    *reinterpret_cast<float4*>(m + i) = mi;
    *reinterpret_cast<float4*>(v + i) = vi;
  } else {
    for (long long j = i; j < n; ++j) adam_one(p[j], g[j], m[j], v[j], b1, b2, eps, wd, step_size, inv_sqrt_bc2);
  }
}

__global__ void adam_tick_kernel(float* step_dev) { *step_dev += 1.f; }

}  // namespace

extern "C" int mlg_adam_step(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* step_dev,
                             int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay,
                             void* stream) {
  MLG_CHECK_ARG(params && grads && exp_avg && exp_avg_sq && step_dev && n >= 0, "mlg_adam_step: bad arguments");
  if (n == 0) return MLG_OK;
  cudaStream_t st = (cudaStream_t)stream;
  MLG_CHECK_ARG(((uintptr_t)params | (uintptr_t)grads | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
                "mlg_adam_step: buffers must be 16-byte aligned");
  adam_kernel<<<mlg_ceil_div(mlg_ceil_div(n, 4), 256), 256, 0, st>>>(params, grads, exp_avg, exp_avg_sq, step_dev, n, lr, beta1, beta2, eps,
                                                   weight_decay);
  MLG_CHECK_LAUNCH("mlg_adam_step");
  adam_tick_kernel<<<1, 1, 0, st>>>(step_dev);
  MLG_CHECK_LAUNCH("mlg_adam_step(tick)");
  return MLG_OK;
}
