// probe: throughput of red.global.add.v4.f32 scatter into an L2-resident [N, 128] fp32 matrix (the source-side sum of the
// GENConv backward done with reductions instead of a second pass over [E, H]).  nvcc -arch=sm_100a -O3 red_v4.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
__global__ void scatter(const int* __restrict__ idx, float* __restrict__ out, int E, int mode) {
  const int lane = threadIdx.x & 31;
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long nw = ((long long)gridDim.x * blockDim.x) >> 5;
  for (long long e = warp; e < E; e += nw) {
    const int r = __ldg(idx + e);
    float* p = out + (size_t)r * 128 + lane * 4;
    const float v = (float)(e & 7);
    if (mode == 0) {
      asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v), "f"(v), "f"(v), "f"(v) : "memory");
    } else if (mode == 1) {
      atomicAdd(p, v); atomicAdd(p + 1, v); atomicAdd(p + 2, v); atomicAdd(p + 3, v);
    } else {
      float4 o = *reinterpret_cast<float4*>(p);   // plain read-modify-write (racy): the bandwidth floor
      o.x += v; o.y += v; o.z += v; o.w += v;
      *reinterpret_cast<float4*>(p) = o;
    }
  }
}
int main() {
  const int N = 100000, E = 1600000;
  int* h = (int*)malloc(E * sizeof(int));
  srand(1);
  for (int i = 0; i < E; ++i) h[i] = rand() % N;
  int* idx; float* out;
  cudaMalloc(&idx, E * sizeof(int));
  cudaMalloc(&out, (size_t)N * 128 * 4);
  cudaMemcpy(idx, h, E * sizeof(int), cudaMemcpyHostToDevice);
  cudaMemset(out, 0, (size_t)N * 128 * 4);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int mode = 0; mode < 3; ++mode)
    for (int blocks : {148 * 4, 148 * 8, 148 * 16}) {
      scatter<<<blocks, 256>>>(idx, out, E, mode);
      cudaEventRecord(a);
      for (int i = 0; i < 5; ++i) scatter<<<blocks, 256>>>(idx, out, E, mode);
      cudaEventRecord(b); cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      printf("mode %d (%s) blocks %d: %.3f ms per pass, %.1f GB/s of payload, err=%s\n", mode,
             mode == 0 ? "red.v4.f32" : mode == 1 ? "4x atomicAdd" : "plain rmw", blocks, ms / 5, E * 512.0 / (ms / 5) / 1e6,
             cudaGetErrorString(cudaGetLastError()));
    }
  return 0;
}
