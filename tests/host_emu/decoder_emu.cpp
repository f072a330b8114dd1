// TEST INFRASTRUCTURE ONLY: compiles csrc/decoder_grouped.cu as plain host C++ (-DMLG_HOST_EMU) so that the grouped
// decoder's forward / backward algebra can be checked against the CPU oracle without a GPU (tests/test_decoder_emu.py).
// The product never loads this.
#define MLG_HOST_EMU 1
#include <stdlib.h>
#include "../../multilevel-gnn_b200/csrc/decoder_grouped.cu"

static void base(dec::Params& P, const float* x, const float* packed, const int64_t* table, int B, int S, int F,
                 long long total_out, long long total_hidden) {
  memset(&P, 0, sizeof(P));
  P.x = x; P.packed = packed; P.table = (const long long*)table; P.B = B; P.S = S; P.F = F;
  P.total_out = total_out; P.total_hidden = total_hidden;
}

extern "C" int emu_decoder_fwd(const float* x, const float* packed, const int64_t* table, int B, int S, int F, int Dmax,
                               long long total_out, long long total_hidden, float* out, float* h) {
  dec::Params P; base(P, x, packed, table, B, S, F, total_out, total_hidden);
  P.out = out; P.h = h;
  float* sm = (float*)aligned_alloc(64, ((size_t)dec::smem_floats(B, F, Dmax, false) * 4 + 63) / 64 * 64);
  for (int i = 0; i < S; ++i) dec::forward_body(P, sm, i, Dmax);
  free(sm);
  return 0;
}

extern "C" int emu_decoder_bwd(const float* g_out, const float* x, const float* h, const float* packed, const int64_t* table,
                               int B, int S, int F, int Dmax, long long total_out, long long total_hidden, float* g_x,
                               float* g_packed, int accumulate) {
  dec::Params P; base(P, x, packed, table, B, S, F, total_out, total_hidden);
  P.h = (float*)h; P.g_out = g_out; P.g_x = g_x; P.g_packed = g_packed; P.accumulate = accumulate;
  float* sm = (float*)aligned_alloc(64, ((size_t)dec::smem_floats(B, F, Dmax, true) * 4 + 63) / 64 * 64);
  for (int i = 0; i < S; ++i) dec::backward_body(P, sm, i, Dmax);
  free(sm);
  return 0;
}
