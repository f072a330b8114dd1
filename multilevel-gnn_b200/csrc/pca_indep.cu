// Pathway-wise independence term of MultilevelGNN.get_feature_loss (models/multilevel_gnn.py:336-346), sm_100a.
//
// With w = learnable_pca_params * info_mask ([G, P], no gradient: the reference reads .data) and the genes grouped in
// pathway segments, the reference accumulates, for every column i < P-1 against the LAST column (only the last j of its
// double loop survives),   |sum_g w_gi w_gL| / (sqrt(sum_g w_gi^2 * sum_g w_gL^2) + 1e-7)   per segment, takes the mean
// over the segments, sums over i and divides by P(P-1)/2.  The library path is ~14 tiny launches per step (cat,
// index_add, sqrt, div, abs, mean, sum, ...); here: one warp per segment (lanes stride over the segment's genes, shuffle
// tree) spread over ceil(nseg / 8) blocks, then one warp adds the per-segment terms in segment order (fixed summation
// order).  (r01 ran all segments through ONE block: 35 us of pure latency on the step's critical path; now ~2 x 3 us.)
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

constexpr int kMaxP = 8;

constexpr int kWarpsPI = 8;

__global__ void __launch_bounds__(kWarpsPI * 32) pca_indep_kernel(const float* __restrict__ w, const float* __restrict__ mask,
                                                                  const int* __restrict__ segptr, int nseg, int P,
                                                                  float* __restrict__ seg_val) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int s = blockIdx.x * kWarpsPI + warp;
  if (s >= nseg) return;
  const int beg = __ldg(segptr + s), end = __ldg(segptr + s + 1);
  float mul[kMaxP - 1], aa[kMaxP - 1], bb = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxP - 1; ++i) mul[i] = aa[i] = 0.f;
  for (int g = beg + lane; g < end; g += 32) {
    const float m = mask ? __ldg(mask + g) : 1.f;
    const float b = __ldg(w + (size_t)g * P + (P - 1)) * m;
    bb = fmaf(b, b, bb);
#pragma unroll
    for (int i = 0; i < kMaxP - 1; ++i)
      if (i < P - 1) {
        const float a = __ldg(w + (size_t)g * P + i) * m;
        mul[i] = fmaf(a, b, mul[i]);
        aa[i] = fmaf(a, a, aa[i]);
      }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    bb += __shfl_xor_sync(0xffffffffu, bb, o);
#pragma unroll
    for (int i = 0; i < kMaxP - 1; ++i)
      if (i < P - 1) {
        mul[i] += __shfl_xor_sync(0xffffffffu, mul[i], o);
        aa[i] += __shfl_xor_sync(0xffffffffu, aa[i], o);
      }
  }
  float total = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxP - 1; ++i)
    if (i < P - 1) total += fabsf(mul[i] / (sqrtf(aa[i] * bb) + 1e-7f));
  if (lane == 0) seg_val[s] = total;
}

// lane l adds segments l, l+32, ... in order, then the 32 partials are added in lane order: fixed for a given nseg
__global__ void pca_indep_sum_kernel(const float* __restrict__ seg_val, int nseg, float scale, float* __restrict__ out) {
  __shared__ float part[32];
  float t = 0.f;
  for (int s = threadIdx.x; s < nseg; s += 32) t += seg_val[s];
  part[threadIdx.x] = t;
  __syncwarp();
  if (threadIdx.x == 0) {
    float u = 0.f;
    for (int k = 0; k < 32; ++k) u += part[k];
    out[0] = u * scale;
  }
}

}  // namespace

extern "C" int mlg_pca_indep_loss(const float* w, const float* mask, const int32_t* segptr, int64_t nseg, int64_t P,
                                  float* out, float* workspace, void* stream) {
  MLG_CHECK_ARG(w && segptr && out && workspace, "mlg_pca_indep_loss: null pointer");
  MLG_CHECK_ARG(nseg >= 1 && P >= 2 && P <= kMaxP, "mlg_pca_indep_loss: needs nseg >= 1 and 2 <= P <= 8");
  const float scale = 1.f / ((float)nseg * (float)(P * (P - 1) / 2));
  cudaStream_t st = (cudaStream_t)stream;
  pca_indep_kernel<<<mlg_ceil_div(nseg, kWarpsPI), kWarpsPI * 32, 0, st>>>(w, mask, segptr, (int)nseg, (int)P, workspace);
  MLG_CHECK_LAUNCH("mlg_pca_indep_loss");
  pca_indep_sum_kernel<<<1, 32, 0, st>>>(workspace, (int)nseg, scale, out);
  MLG_CHECK_LAUNCH("mlg_pca_indep_loss(sum)");
  return MLG_OK;
}
