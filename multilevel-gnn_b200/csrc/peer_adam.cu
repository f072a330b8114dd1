// Data-parallel optimizer step as ONE kernel over NVLink / NVSwitch peer memory (sm_100a):
//     reduce-scatter(gradients) -> Adam on the owned shard -> all-gather(parameters)
// The reference trains on one device (train.py:38-68: loss.backward(); optimizer.step()); patient graphs are independent,
// so the data-parallel wrapper only has to average the gradients.  Instead of an NCCL all-reduce
// followed by a replicated Adam, every rank keeps its flat gradient bucket, its flat parameter buffer and a small flag
// block in a cudaMalloc'ed arena that is mapped into every other rank of the box (CUDA IPC), and this kernel
//   0. tells every peer "my gradients of step e are final" and waits for the same from all of them,
//   1. for the 1/world shard it owns: loads that slice of every peer's gradients straight over NVLink (fixed rank order
//      0..world-1, so every replica gets bitwise-identical parameters), averages, applies torch.optim.Adam's update to
//      its shard of the optimizer state, and stores the new parameter values into EVERY peer's parameter buffer,
//   2. tells every peer "my writes of step e are done" and waits for the same from all before it retires.
// No gradient is ever written back, Adam state exists once per box (ZeRO-1), and since it is an ordinary kernel launch
// with fixed pointers the whole training step (forward, loss, backward, this kernel) is ONE CUDA graph per rank.
// Per-GPU traffic: (world-1)/world * 4n bytes read and the same written over NVLink + 16 B/parameter of local state.
//
// Flag block (uint32, one per rank, peer-writable):  [0..7] ready[src]   [8..15] done[src]   [16] epoch   [17] block
// counter   [18] status (1 = a wait timed out: peers out of step; the host checks it outside the timed path).
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int kMaxWorld = 8;
constexpr int kThreads = 512;
constexpr int F_READY = 0, F_DONE = 8, F_EPOCH = 16, F_COUNT = 17, F_STATUS = 18;

struct PeerP {
  const float* grad[kMaxWorld];
  float* param[kMaxWorld];
  unsigned* flags[kMaxWorld];
  float* m;          // exp_avg of the owned shard      [hi - lo]
  float* v;          // exp_avg_sq of the owned shard   [hi - lo]
  float* step_dev;   // Adam step counter (float, as mlg_adam_step)
  long long lo, hi;  // owned shard, multiples of 4
  int world, rank;
  float lr, b1, b2, eps, wd;
  long long timeout_clocks;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer4(const float* p) {   // never cached: the owner rewrites it every step
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// wait until flag[i] has reached `epoch` for every rank i < world (threads 0..world-1 poll one flag each)
__device__ __forceinline__ void wait_all(unsigned* mine, int base, int world, unsigned epoch, long long timeout) {
  if ((int)threadIdx.x < world) {
    const unsigned* f = mine + base + threadIdx.x;
    const long long t0 = clock64();
    while ((int)(ld_acquire_sys(f) - epoch) < 0) {
      if (clock64() - t0 > timeout) {
        mine[F_STATUS] = 1u;
        break;
      }
      __nanosleep(64);
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float b1, float b2, float eps, float wd,
                                         float step_size, float inv_sqrt_bc2) {
  if (wd != 0.f) g = fmaf(wd, p, g);
  m = fmaf(b1, m, (1.f - b1) * g);
  v = fmaf(b2, v, (1.f - b2) * g * g);
  const float denom = sqrtf(v) * inv_sqrt_bc2 + eps;     // torch: (sqrt(v) / sqrt(bias_correction2)) + eps
  p -= step_size * (m / denom);
}

__global__ void __launch_bounds__(kThreads) peer_adam_kernel(const PeerP P) {
  __shared__ int is_last;
  unsigned* mine = P.flags[P.rank];
  const int world = P.world;
  // the epoch word is only advanced by the LAST block of a run, after every block of that run has read it
  const unsigned epoch = *reinterpret_cast<volatile unsigned*>(mine + F_EPOCH) + 1u;

  // ---- 0. gradients of every rank are final ----
  if (blockIdx.x == 0 && (int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(P.flags[threadIdx.x] + F_READY + P.rank, epoch);
  }
  wait_all(mine, F_READY, world, epoch, P.timeout_clocks);

  // ---- 1. owned shard: sum over ranks in rank order, Adam, broadcast the new values ----
  const float t = *reinterpret_cast<volatile float*>(P.step_dev) + 1.f;
  const float bc1 = 1.f - powf(P.b1, t), bc2 = 1.f - powf(P.b2, t);
  const float step_size = P.lr / bc1, inv_sqrt_bc2 = 1.f / sqrtf(bc2);
  const float inv_world = 1.f / (float)world;
  const long long stride = (long long)gridDim.x * kThreads * 4;
  for (long long i = P.lo + ((long long)blockIdx.x * kThreads + threadIdx.x) * 4; i < P.hi; i += stride) {
    float4 g[kMaxWorld];
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) g[r] = ld_peer4(P.grad[r] + i);
    float4 s = g[0];
#pragma unroll
    for (int r = 1; r < kMaxWorld; ++r)
      if (r < world) {
        s.x += g[r].x; s.y += g[r].y; s.z += g[r].z; s.w += g[r].w;
      }
    s.x *= inv_world; s.y *= inv_world; s.z *= inv_world; s.w *= inv_world;
    const long long j = i - P.lo;
    float4 pi = *reinterpret_cast<const float4*>(P.param[P.rank] + i);
    float4 mi = *reinterpret_cast<const float4*>(P.m + j), vi = *reinterpret_cast<const float4*>(P.v + j);
    adam_one(pi.x, s.x, mi.x, vi.x, P.b1, P.b2, P.eps, P.wd, step_size, inv_sqrt_bc2);
    adam_one(pi.y, s.y, mi.y, vi.y, P.b1, P.b2, P.eps, P.wd, step_size, inv_sqrt_bc2);
    adam_one(pi.z, s.z, mi.z, vi.z, P.b1, P.b2, P.eps, P.wd, step_size, inv_sqrt_bc2);
    adam_one(pi.w, s.w, mi.w, vi.w, P.b1, P.b2, P.eps, P.wd, step_size, inv_sqrt_bc2);
    *reinterpret_cast<float4*>(P.m + j) = mi;
    *reinterpret_cast<float4*>(P.v + j) = vi;
#pragma unroll
    for (int r = 0; r < kMaxWorld; ++r)
      if (r < world) *reinterpret_cast<float4*>(P.param[r] + i) = pi;
  }

  // ---- 2. every rank's parameter writes have landed ----
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned c = atomicAdd(mine + F_COUNT, 1u);
    __threadfence();
    is_last = (c == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  if ((int)threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(P.flags[threadIdx.x] + F_DONE + P.rank, epoch);
  }
  wait_all(mine, F_DONE, world, epoch, P.timeout_clocks);
  if (threadIdx.x == 0) {
    mine[F_COUNT] = 0u;
    *reinterpret_cast<volatile float*>(P.step_dev) = t;
    __threadfence();
    *reinterpret_cast<volatile unsigned*>(mine + F_EPOCH) = epoch;
  }
}

}  // namespace

extern "C" int64_t mlg_peer_flag_bytes(void) { return 256; }

extern "C" void* mlg_peer_alloc(int64_t bytes) {
  void* p = nullptr;
  if (bytes <= 0) return nullptr;
  cudaError_t e = cudaMalloc(&p, (size_t)bytes);
  if (e != cudaSuccess) {
    mlg_set_error("mlg_peer_alloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
    return nullptr;
  }
  e = cudaMemset(p, 0, (size_t)bytes);
  if (e != cudaSuccess) {
    mlg_set_error("mlg_peer_alloc: memset: %s", cudaGetErrorString(e));
    cudaFree(p);
    return nullptr;
  }
  return p;
}

extern "C" int mlg_peer_free(void* ptr) {
  if (ptr && cudaFree(ptr) != cudaSuccess) {
    mlg_set_error("mlg_peer_free: %s", cudaGetErrorString(cudaGetLastError()));
    return MLG_ERR_CUDA;
  }
  return MLG_OK;
}

extern "C" int mlg_peer_export(void* ptr, void* handle64) {
  MLG_CHECK_ARG(ptr && handle64, "mlg_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), ptr);
  if (e != cudaSuccess) {
    mlg_set_error("mlg_peer_export: %s", cudaGetErrorString(e));
    return MLG_ERR_CUDA;
  }
  return MLG_OK;
}

extern "C" void* mlg_peer_open(const void* handle64) {
  if (!handle64) return nullptr;
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  void* p = nullptr;
  cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) {
    mlg_set_error("mlg_peer_open: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return nullptr;
  }
  return p;
}

extern "C" int mlg_peer_close(void* ptr) {
  if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) {
    mlg_set_error("mlg_peer_close: %s", cudaGetErrorString(cudaGetLastError()));
    return MLG_ERR_CUDA;
  }
  return MLG_OK;
}

extern "C" int mlg_peer_adam_step(const float* const* peer_grads, float* const* peer_params, void* const* peer_flags,
                                  int world, int rank, int64_t n_padded, float* exp_avg, float* exp_avg_sq,
                                  float* step_dev, float lr, float beta1, float beta2, float eps, float weight_decay,
                                  double timeout_s, void* stream) {
  MLG_CHECK_ARG(peer_grads && peer_params && peer_flags && exp_avg && exp_avg_sq && step_dev, "mlg_peer_adam_step: null pointer");
  MLG_CHECK_ARG(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world, "mlg_peer_adam_step: world must be 1..8");
  MLG_CHECK_ARG(n_padded >= 0 && n_padded % (4 * world) == 0, "mlg_peer_adam_step: n_padded must be a multiple of 4*world");
  if (n_padded == 0) return MLG_OK;
  PeerP P;
  memset(&P, 0, sizeof(P));
  for (int r = 0; r < world; ++r) {
    MLG_CHECK_ARG(peer_grads[r] && peer_params[r] && peer_flags[r], "mlg_peer_adam_step: null peer pointer");
    MLG_CHECK_ARG(((uintptr_t)peer_grads[r] | (uintptr_t)peer_params[r]) % 16 == 0, "mlg_peer_adam_step: 16-byte alignment");
    P.grad[r] = peer_grads[r];
    P.param[r] = peer_params[r];
    P.flags[r] = (unsigned*)peer_flags[r];
  }
  MLG_CHECK_ARG(((uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0, "mlg_peer_adam_step: 16-byte alignment");
  const long long shard = n_padded / world;
  P.m = exp_avg; P.v = exp_avg_sq; P.step_dev = step_dev;
  P.lo = shard * rank; P.hi = P.lo + shard;
  P.world = world; P.rank = rank;
  P.lr = lr; P.b1 = beta1; P.b2 = beta2; P.eps = eps; P.wd = weight_decay;
  P.timeout_clocks = (long long)((timeout_s > 0 ? timeout_s : 5.0) * 1.9e9);
  // enough blocks to keep ~world float4 loads per thread in flight on every NVLink, few enough that the whole grid is
  // resident next to whatever else the stream overlaps
  long long blocks = (shard / 4 + kThreads - 1) / kThreads;
  if (blocks > 96) blocks = 96;
  if (blocks < 1) blocks = 1;
  peer_adam_kernel<<<(unsigned)blocks, kThreads, 0, (cudaStream_t)stream>>>(P);
  MLG_CHECK_LAUNCH("mlg_peer_adam_step");
  return MLG_OK;
}

extern "C" int mlg_peer_status(const void* flags, int* status_out) {
  MLG_CHECK_ARG(flags && status_out, "mlg_peer_status: null pointer");
  unsigned s = 0;
  cudaError_t e = cudaMemcpy(&s, (const unsigned*)flags + F_STATUS, sizeof(s), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) {
    mlg_set_error("mlg_peer_status: %s", cudaGetErrorString(e));
    return MLG_ERR_CUDA;
  }
  *status_out = (int)s;
  return MLG_OK;
}
