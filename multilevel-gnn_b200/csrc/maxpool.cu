// Max-pool of MultilevelGNN's head over a channel-LAST activation, writing the NCHW result the flatten expects (sm_100a).
//
// models/multilevel_gnn.py:286 of the reference applies nn.MaxPool2d((pathway_pool_dim, pca_pool_dim)) (stride = kernel,
// no padding, floor mode) to the [B, C, 146, 3P] conv output.  Here that tensor lives channel-last in memory
// ([B, H, W, C], the layout the pooled features and the 1x1 convs use), so the library path needed an NCHW copy before
// the pool and another one in backward (4 launches, ~45 us).  One kernel each way: lanes run over the channels
// (coalesced channel-last reads), the output is written in [B, C, Ho, Wo] order; backward routes each output gradient
// to the FIRST maximum of its window in row-major scan order (ATen's tie rule) and zero-fills the rest, including the
// rows / columns the floor mode drops.
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

__global__ void maxpool_cl_fwd_kernel(const float* __restrict__ x, int B, int H, int W, int C, int kh, int kw, int Ho,
                                      int Wo, float* __restrict__ out, unsigned char* __restrict__ arg) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // over (b, ho, wo, c), c fastest
  const long long total = (long long)B * Ho * Wo * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int wo = (int)(r % Wo); r /= Wo;
  const int ho = (int)(r % Ho);
  const int b = (int)(r / Ho);
  const float* base = x + (((size_t)b * H + (size_t)ho * kh) * W + (size_t)wo * kw) * C + c;
  float best = -INFINITY;
  int bi = 0;
  for (int dh = 0; dh < kh; ++dh)
    for (int dw = 0; dw < kw; ++dw) {
      const float v = __ldg(base + ((size_t)dh * W + dw) * C);
      if (v > best || v != v) {   // first maximum wins; NaN propagates like ATen
        if (!(best != best)) { best = v; bi = dh * kw + dw; }
      }
    }
  const size_t o = (((size_t)b * C + c) * Ho + ho) * Wo + wo;   // NCHW
  out[o] = best;
  arg[o] = (unsigned char)bi;
}

__global__ void maxpool_cl_bwd_kernel(const float* __restrict__ g, const unsigned char* __restrict__ arg, int B, int H,
                                      int W, int C, int kh, int kw, int Ho, int Wo, float* __restrict__ gx) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;   // over (b, h, w, c) of the INPUT, c fastest
  const long long total = (long long)B * H * W * C;
  if (i >= total) return;
  const int c = (int)(i % C);
  long long r = i / C;
  const int w = (int)(r % W); r /= W;
  const int h = (int)(r % H);
  const int b = (int)(r / H);
  const int ho = h / kh, wo = w / kw;
  float v = 0.f;
  if (ho < Ho && wo < Wo) {
    const size_t o = (((size_t)b * C + c) * Ho + ho) * Wo + wo;
    if ((int)arg[o] == (h - ho * kh) * kw + (w - wo * kw)) v = __ldg(g + o);
  }
  gx[i] = v;
}

}  // namespace

extern "C" int mlg_maxpool_cl_fwd(const float* x_cl, int64_t B, int64_t H, int64_t W, int64_t C, int64_t kh, int64_t kw,
                                  float* out_nchw, uint8_t* argmax, void* stream) {
  MLG_CHECK_ARG(x_cl && out_nchw && argmax, "mlg_maxpool_cl_fwd: null pointer");
  MLG_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1 && kh >= 1 && kw >= 1 && kh * kw <= 255 && kh <= H && kw <= W,
                "mlg_maxpool_cl_fwd: bad sizes (window %lld x %lld on %lld x %lld)", (long long)kh, (long long)kw,
                (long long)H, (long long)W);
  const long long Ho = H / kh, Wo = W / kw, total = B * Ho * Wo * C;
  if (total == 0) return MLG_OK;
  maxpool_cl_fwd_kernel<<<(unsigned)mlg_ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      x_cl, (int)B, (int)H, (int)W, (int)C, (int)kh, (int)kw, (int)Ho, (int)Wo, out_nchw, argmax);
  MLG_CHECK_LAUNCH("mlg_maxpool_cl_fwd");
  return MLG_OK;
}

extern "C" int mlg_maxpool_cl_bwd(const float* g_out_nchw, const uint8_t* argmax, int64_t B, int64_t H, int64_t W,
                                  int64_t C, int64_t kh, int64_t kw, float* g_x_cl, void* stream) {
  MLG_CHECK_ARG(g_out_nchw && argmax && g_x_cl, "mlg_maxpool_cl_bwd: null pointer");
  MLG_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1 && kh >= 1 && kw >= 1 && kh * kw <= 255 && kh <= H && kw <= W,
                "mlg_maxpool_cl_bwd: bad sizes");
  const long long total = B * H * W * C;
  if (total == 0) return MLG_OK;
  maxpool_cl_bwd_kernel<<<(unsigned)mlg_ceil_div(total, 256), 256, 0, (cudaStream_t)stream>>>(
      g_out_nchw, argmax, (int)B, (int)H, (int)W, (int)C, (int)kh, (int)kw, (int)(H / kh), (int)(W / kw), g_x_cl);
  MLG_CHECK_LAUNCH("mlg_maxpool_cl_bwd");
  return MLG_OK;
}
