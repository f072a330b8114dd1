// Per-pathway decoders of the VAE ("unpool") as ONE grouped kernel per direction (sm_100a).
//
// Reference: VAE.foreach_decoder (models/vae.py:216-222) over the blocks built at :54-74:
//   pred[:, genes of pathway i] = Linear_i2(ReLU(Linear_i0(h[:, i, :])))          for i in range(438), torch.cat on dim -1
// i.e. 438 x (Linear, ReLU, Linear) + a concatenation = ~1750 library launches forward and ~3500 backward, each on a
// [B x 96] x [96 x D] sized problem.  The blocks share nothing but the batch: every weight is used by exactly B rows.
//
// Here all blocks' parameters live in ONE packed buffer (W1_i [D_i, F], b1_i [D_i], W2_i [n_i, D_i], b2_i [n_i] back to
// back; `table` holds the offsets) and one CTA owns one pathway: its B input rows and the B x D_i hidden activations stay
// in shared memory, the weights are streamed from global memory once per CTA, and the outputs are written straight into
// their column range of pred (no concatenation).  Backward does the five contractions of the block (dW2, dH, dW1, dX and
// the two bias sums) in the same CTA from the saved post-ReLU activations; parameter gradients have a single writer per
// element (the pathway's CTA), so nothing is atomic and the result is deterministic.
// HBM-bound on the packed weights: algorithmic bytes forward = 4 * (sum_i (F + 1) D_i + (D_i + 1) n_i) + 4 B (S F + N_out).
//
// Batches larger than what fits next to the hidden activations in shared memory are processed in row chunks (parameter
// gradients accumulate over the chunks in launch order).
#include "block_mm.cuh"

namespace dec {

using namespace bmm;

constexpr int kCols = 8;   // table columns: w1, b1, w2, b2 (float offsets into packed), D, n_out, out_off, h_off

struct Params {
  const float* x;        // [B, S, F] (this chunk's first row)
  const float* packed;
  const long long* table;   // [S, kCols]
  int B, S, F;
  long long total_out, total_hidden;
  float* out;            // [B, total_out]
  float* h;              // [B, total_hidden] post-ReLU hidden activations (forward: written when non-null; backward: read)
  // backward
  const float* g_out;    // [B, total_out]
  float* g_x;            // [B, S, F]
  float* g_packed;
  int accumulate;        // parameter gradients: 0 = overwrite, 1 = add (later row chunks)
};

struct Block { const float *W1, *b1, *W2, *b2; float *gW1, *gb1, *gW2, *gb2; int D, n; long long out_off, h_off; };

MLG_DEV Block load_block(const Params& P, int i) {
  const long long* t = P.table + (size_t)i * kCols;
  Block b;
  b.W1 = P.packed + t[0]; b.b1 = P.packed + t[1]; b.W2 = P.packed + t[2]; b.b2 = P.packed + t[3];
  b.gW1 = b.gb1 = b.gW2 = b.gb2 = nullptr;
  if (P.g_packed) { b.gW1 = P.g_packed + t[0]; b.gb1 = P.g_packed + t[1]; b.gW2 = P.g_packed + t[2]; b.gb2 = P.g_packed + t[3]; }
  b.D = (int)t[4]; b.n = (int)t[5]; b.out_off = t[6]; b.h_off = t[7];
  return b;
}

// shared memory: Xs [B x P4(F)] | Hs [B x P4(Dmax)] (| dHs [B x P4(Dmax)] in backward)
MLG_DEV void load_x(const Params& P, int i, float* Xs) {
  const int Fp = P4(P.F);
  MLG_PFOR(t, P.B * P.F) {
    const int b = t / P.F, f = t - b * P.F;
    Xs[b * Fp + f] = P.x[((size_t)b * P.S + i) * P.F + f];
  }
  MLG_SYNC();
}

MLG_DEV void forward_body(const Params& P, float* sm, int i, int Dmax) {
  const Block k = load_block(P, i);
  const int B = P.B, F = P.F, Fp = P4(F), Dp = P4(k.D);
  float* Xs = sm;
  float* Hs = sm + B * Fp;
  (void)Dmax;
  load_x(P, i, Xs);
  // H = ReLU(X W1^T + b1)
  mm(Hs, Dp, Xs, Fp, 1, true, k.W1, 1, F, false, B, k.D, F, 1.f, false);
  MLG_PFOR(t, B * k.D) {
    const int b = t / k.D, d = t - b * k.D;
    const float v = fmaxf(Hs[b * Dp + d] + k.b1[d], 0.f);
    Hs[b * Dp + d] = v;
    if (P.h) P.h[(size_t)b * P.total_hidden + k.h_off + d] = v;
  }
  MLG_SYNC();
  // pred[:, out_off : out_off + n] = H W2^T + b2   (bias first, then the product accumulates onto it)
  float* Y = P.out + k.out_off;
  MLG_PFOR(t, B * k.n) {
    const int b = t / k.n, j = t - b * k.n;
    Y[(size_t)b * P.total_out + j] = k.b2[j];
  }
  MLG_SYNC();
  mm(Y, (int)P.total_out, Hs, Dp, 1, true, k.W2, 1, k.D, false, B, k.n, k.D, 1.f, true);
}

MLG_DEV void backward_body(const Params& P, float* sm, int i, int Dmax) {
  const Block k = load_block(P, i);
  const int B = P.B, F = P.F, Fp = P4(F), Dp = P4(k.D);
  float* Xs = sm;
  float* Hs = sm + B * Fp;
  float* dHs = Hs + B * P4(Dmax);
  const bool acc = P.accumulate != 0;
  const int ldo = (int)P.total_out;
  const float* dY = P.g_out + k.out_off;    // [B x n], row stride total_out (global)
  load_x(P, i, Xs);
  MLG_PFOR(t, B * k.D) {
    const int b = t / k.D, d = t - b * k.D;
    Hs[b * Dp + d] = P.h[(size_t)b * P.total_hidden + k.h_off + d];
  }
  MLG_SYNC();
  // gW2 [n x D] = dY^T H ; gb2 = colsum dY
  mm(k.gW2, k.D, dY, 1, ldo, false, Hs, Dp, 1, true, k.n, k.D, B, 1.f, acc);
  MLG_PFOR(j, k.n) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dY[(size_t)b * ldo + j];
    k.gb2[j] = acc ? k.gb2[j] + s : s;
  }
  // dH = (dY W2) masked by H > 0
  mm(dHs, Dp, dY, ldo, 1, false, k.W2, k.D, 1, false, B, k.D, k.n, 1.f, false);
  MLG_PFOR(t, B * k.D) {
    const int b = t / k.D, d = t - b * k.D;
    if (!(Hs[b * Dp + d] > 0.f)) dHs[b * Dp + d] = 0.f;
  }
  MLG_SYNC();
  // gW1 [D x F] = dH^T X ; gb1 = colsum dH
  mm(k.gW1, F, dHs, 1, Dp, true, Xs, Fp, 1, true, k.D, F, B, 1.f, acc);
  MLG_PFOR(d, k.D) {
    float s = 0.f;
    for (int b = 0; b < B; ++b) s += dHs[b * Dp + d];
    k.gb1[d] = acc ? k.gb1[d] + s : s;
  }
  // dX [B x F] = dH W1   (into g_x[:, i, :], row stride S F)
  if (P.g_x) mm(P.g_x + (size_t)i * F, P.S * F, dHs, Dp, 1, true, k.W1, F, 1, false, B, F, k.D, 1.f, false);
}

// floats of shared memory for a chunk of B rows
static inline long long smem_floats(int B, int F, int Dmax, bool backward) {
  const long long Fp = (F + 3) & ~3, Dp = (Dmax + 3) & ~3;
  return (long long)B * (Fp + (backward ? 2 : 1) * Dp);
}

}  // namespace dec

#ifndef MLG_HOST_EMU
namespace {

constexpr int kThreadsDec = 512;
constexpr int kMaxSmemBytes = 227 * 1024;

__global__ void __launch_bounds__(kThreadsDec, 1) decoder_fwd_kernel(const dec::Params P, int Dmax) {
  extern __shared__ __align__(16) float dec_sm[];
  dec::forward_body(P, dec_sm, blockIdx.x, Dmax);
}

__global__ void __launch_bounds__(kThreadsDec, 1) decoder_bwd_kernel(const dec::Params P, int Dmax) {
  extern __shared__ __align__(16) float dec_sm[];
  dec::backward_body(P, dec_sm, blockIdx.x, Dmax);
}

int max_rows(int64_t F, int64_t Dmax, bool backward) {
  const long long per_row = dec::smem_floats(1, (int)F, (int)Dmax, backward) * 4;
  return (int)(kMaxSmemBytes / per_row);
}

}  // namespace

extern "C" int64_t mlg_decoder_max_rows(int64_t F, int64_t Dmax, int backward) {
  if (F < 1 || Dmax < 1) return 0;
  return max_rows(F, Dmax, backward != 0);
}

extern "C" int mlg_decoder_fwd(const float* x, const float* packed, const int64_t* table, int64_t B, int64_t S, int64_t F,
                               int64_t Dmax, int64_t total_out, int64_t total_hidden, float* out, float* h_saved,
                               void* stream) {
  MLG_CHECK_ARG(x && packed && table && out && B >= 1 && S >= 1 && F >= 1 && Dmax >= 1, "mlg_decoder_fwd: bad arguments");
  MLG_CHECK_ARG(total_out < (1ll << 31) && S * F < (1ll << 31), "mlg_decoder_fwd: sizes exceed int32");
  const int rows = max_rows(F, Dmax, false);
  MLG_CHECK_ARG(rows >= 1, "mlg_decoder_fwd: one row (F=%lld, D=%lld) does not fit in shared memory", (long long)F, (long long)Dmax);
  MLG_CUDA(cudaFuncSetAttribute(decoder_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes));
  for (int64_t b0 = 0; b0 < B; b0 += rows) {
    dec::Params P;
    memset(&P, 0, sizeof(P));
    P.B = (int)(B - b0 < rows ? B - b0 : rows);
    P.S = (int)S; P.F = (int)F;
    P.total_out = total_out; P.total_hidden = total_hidden;
    P.x = x + (size_t)b0 * S * F;
    P.packed = packed;
    P.table = (const long long*)table;
    P.out = out + (size_t)b0 * total_out;
    P.h = h_saved ? h_saved + (size_t)b0 * total_hidden : nullptr;
    const size_t smem = (size_t)dec::smem_floats(P.B, P.F, (int)Dmax, false) * 4;
    decoder_fwd_kernel<<<(unsigned)S, kThreadsDec, smem, (cudaStream_t)stream>>>(P, (int)Dmax);
    MLG_CHECK_LAUNCH("mlg_decoder_fwd");
  }
  return MLG_OK;
}

extern "C" int mlg_decoder_bwd(const float* g_out, const float* x, const float* h_saved, const float* packed,
                               const int64_t* table, int64_t B, int64_t S, int64_t F, int64_t Dmax, int64_t total_out,
                               int64_t total_hidden, float* g_x, float* g_packed, void* stream) {
  MLG_CHECK_ARG(g_out && x && h_saved && packed && table && g_packed && B >= 1 && S >= 1 && F >= 1 && Dmax >= 1,
                "mlg_decoder_bwd: bad arguments");
  MLG_CHECK_ARG(total_out < (1ll << 31) && S * F < (1ll << 31), "mlg_decoder_bwd: sizes exceed int32");
  const int rows = max_rows(F, Dmax, true);
  MLG_CHECK_ARG(rows >= 1, "mlg_decoder_bwd: one row (F=%lld, D=%lld) does not fit in shared memory", (long long)F, (long long)Dmax);
  MLG_CUDA(cudaFuncSetAttribute(decoder_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxSmemBytes));
  for (int64_t b0 = 0; b0 < B; b0 += rows) {
    dec::Params P;
    memset(&P, 0, sizeof(P));
    P.B = (int)(B - b0 < rows ? B - b0 : rows);
    P.S = (int)S; P.F = (int)F;
    P.total_out = total_out; P.total_hidden = total_hidden;
    P.x = x + (size_t)b0 * S * F;
    P.packed = packed;
    P.table = (const long long*)table;
    P.h = const_cast<float*>(h_saved) + (size_t)b0 * total_hidden;
    P.g_out = g_out + (size_t)b0 * total_out;
    P.g_x = g_x ? g_x + (size_t)b0 * S * F : nullptr;
    P.g_packed = g_packed;
    P.accumulate = b0 > 0;
    const size_t smem = (size_t)dec::smem_floats(P.B, P.F, (int)Dmax, true) * 4;
    decoder_bwd_kernel<<<(unsigned)S, kThreadsDec, smem, (cudaStream_t)stream>>>(P, (int)Dmax);
    MLG_CHECK_LAUNCH("mlg_decoder_bwd");
  }
  return MLG_OK;
}
#endif  // MLG_HOST_EMU
