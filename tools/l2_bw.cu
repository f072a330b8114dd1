// L2 -> SM read bandwidth probe: every CTA sweeps an L2-resident buffer (default 32 MB) with 16-byte loads.
// Prints GB/s for coalesced streaming reads and for 256-byte-row random gathers (the SAGE aggregation pattern).
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__global__ void sweep(const float4* __restrict__ buf, size_t n4, int iters, float* sink) {
  float4 a = make_float4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x + (size_t)it * 977 * 32;
#pragma unroll 8
    for (size_t k = 0; k < n4 / stride; ++k) {
      size_t j = i + k * stride;
      if (j >= n4) j -= n4;
      const float4 v = __ldg(buf + j);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  }
  if (a.x + a.y + a.z + a.w == 123.456f) *sink = a.x;
}

// each 16-lane group reads a random 256 B row per step, 8 independent rows in flight
__global__ void gather(const float4* __restrict__ buf, unsigned rows, int steps, float* sink) {
  float4 a = make_float4(0, 0, 0, 0);
  const unsigned grp = (blockIdx.x * blockDim.x + threadIdx.x) >> 4, sl = threadIdx.x & 15;
  unsigned s = grp * 2654435761u + 12345u;
  for (int it = 0; it < steps; ++it) {
    float4 v[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      s = s * 1664525u + 1013904223u;
      v[r] = __ldg(buf + (size_t)((s >> 8) % rows) * 16 + sl);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) { a.x += v[r].x; a.y += v[r].y; a.z += v[r].z; a.w += v[r].w; }
  }
  if (a.x + a.y + a.z + a.w == 123.456f) *sink = a.x;
}

int main(int argc, char** argv) {
  const size_t mb = argc > 1 ? atoi(argv[1]) : 32;
  const size_t bytes = mb << 20, n4 = bytes / 16;
  float4* buf; float* sink;
  cudaMalloc(&buf, bytes); cudaMalloc(&sink, 4);
  cudaMemset(buf, 0, bytes);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024}) {
    const int grid = 148 * (2048 / threads);
    const int iters = 20;
    sweep<<<grid, threads>>>(buf, n4, 2, sink);
    cudaEventRecord(e0);
    sweep<<<grid, threads>>>(buf, n4, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double moved = (double)(n4 / ((size_t)grid * threads)) * grid * threads * 16.0 * iters;
    printf("sweep %zu MB threads=%d: %.0f GB/s\n", mb, threads, moved / ms / 1e6);
  }
  for (int bps : {3, 6, 8}) {
    const int grid = 148 * bps, threads = 256, steps = 400;
    gather<<<grid, threads>>>(buf, (unsigned)(bytes / 256), 10, sink);
    cudaEventRecord(e0);
    gather<<<grid, threads>>>(buf, (unsigned)(bytes / 256), steps, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("gather256 %zu MB blocks/SM=%d: %.0f GB/s (%s)\n", mb, bps, (double)grid * threads * steps * 8 * 16.0 / ms / 1e6,
           cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
