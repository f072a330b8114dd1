// Shared device/host helpers for the sm_100a kernels of multilevel-gnn_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#define MLG_OK 0
#define MLG_ERR_ARG (-1)
#define MLG_ERR_CUDA (-2)
#define MLG_ERR_UNSUPPORTED (-3)
#define MLG_ERR_WORKSPACE (-4)

// thread-local error string, read through mlg_last_error()
void mlg_set_error(const char* fmt, ...);

#define MLG_CHECK_ARG(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      mlg_set_error(__VA_ARGS__);           \
      return MLG_ERR_ARG;                   \
    }                                       \
  } while (0)

#define MLG_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    cudaError_t _e = cudaGetLastError();                                         \
    if (_e != cudaSuccess) {                                                     \
      mlg_set_error("%s: CUDA launch failed: %s", name, cudaGetErrorString(_e)); \
      return MLG_ERR_CUDA;                                                       \
    }                                                                            \
  } while (0)

#define MLG_CUDA(call)                                                              \
  do {                                                                              \
    cudaError_t _e = (call);                                                        \
    if (_e != cudaSuccess) {                                                        \
      mlg_set_error("%s failed: %s", #call, cudaGetErrorString(_e));                \
      return MLG_ERR_CUDA;                                                          \
    }                                                                               \
  } while (0)

static inline int mlg_ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
#define MLG_LOG2E 1.4426950408889634f
#define MLG_LN2 0.6931471805599453f

// streaming 128-bit load: read-once data (edge features), keep it out of L1
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
// gather 128-bit load through the read-only path (rows re-read by many edges)
__device__ __forceinline__ float4 ld_gather4(const float* p) {
  return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
// streaming 128-bit store (written once, not re-read by this kernel)
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
  asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}

// packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 -- one issue slot, two IEEE-identical results)
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// single-MUFU 2^x (ex2.approx.ftz: rel. error 2^-22, flushes denormal results to 0)
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int W>
__device__ __forceinline__ float group_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = W / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o, W);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) { return group_sum<32>(v, 0xffffffffu); }

// block-level sum of NV values held by every thread; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(float (&v)[NV], float* smem /* >= NV*32 floats */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) smem[i * 32 + wid] = v[i];
  }
  __syncthreads();
  if (wid == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float t = lane < nw ? smem[i * 32 + lane] : 0.f;
      v[i] = warp_sum(t);
    }
  }
}

// csrc/gemm_tf32x3.cu: one column chunk of the tensor-core kNN candidate search (internal, used by csrc/knn.cu)
int mlg_tf32x3_knn_chunk(const float* A, const float* B_hi, const float* B_lo, const float* sq_cols, int64_t M, int64_t Nc,
                         int64_t valid, int64_t K, int kc, float* list_d, int* list_i, int col0, void* stream);
#endif
