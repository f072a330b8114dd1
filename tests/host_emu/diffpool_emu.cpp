// TEST INFRASTRUCTURE ONLY: compiles csrc/diffpool_fused.cu as plain host C++ (-DMLG_HOST_EMU: the items of every
// "parallel for + barrier" phase run sequentially) so that the forward / backward algebra of the fused DiffPool kernel can
// be checked against the CPU oracle without a GPU (tests/test_diffpool_emu.py).  The product never loads this.
#define MLG_HOST_EMU 1
#include <stdlib.h>
#include "../../multilevel-gnn_b200/csrc/diffpool_fused.cu"

static int fill(dpf::Params& P, int layers, const int64_t* dims, const float* const* weights, const float* x, const float* adj, int b) {
  memset(&P, 0, sizeof(P));
  P.layers = layers;
  int off = 0;
  for (int l = 0; l < layers; ++l) {
    P.d[l].n = (int)dims[4 * l]; P.d[l].c = (int)dims[4 * l + 1]; P.d[l].k = (int)dims[4 * l + 2]; P.d[l].h = (int)dims[4 * l + 3];
    const float* const* w = weights + 9 * l;
    P.pool[l] = {w[0], w[1], w[2]}; P.embed[l] = {w[3], w[4], w[5]}; P.after[l] = {w[6], w[7], w[8]};
    const int c = P.d[l].c, k = P.d[l].k, h = P.d[l].h;
    const int sizes[9] = {k * c, k * c, k, h * c, h * c, h, h * h, h * h, h};
    for (int q = 0; q < 9; ++q) { P.grad_off[l][q] = off; off += sizes[q]; }
  }
  P.grad_floats = off; P.x = x; P.adj = adj; P.b = b;
  return off;
}

extern "C" long emu_diffpool_smem_floats(int layers, const int64_t* dims) {
  dpf::Params P; memset(&P, 0, sizeof(P)); P.layers = layers;
  for (int l = 0; l < layers; ++l) P.d[l] = {(int)dims[4 * l], (int)dims[4 * l + 1], (int)dims[4 * l + 2], (int)dims[4 * l + 3]};
  dpf::MemMap mp; return dpf::build_map(P, mp);
}

extern "C" long emu_diffpool_state_floats(int layers, const int64_t* dims) {
  dpf::Params P; memset(&P, 0, sizeof(P)); P.layers = layers;
  for (int l = 0; l < layers; ++l) P.d[l] = {(int)dims[4 * l], (int)dims[4 * l + 1], (int)dims[4 * l + 2], (int)dims[4 * l + 3]};
  dpf::MemMap mp; dpf::build_map(P, mp); return mp.state_floats;
}

extern "C" int emu_diffpool_fwd(const float* x, const float* adj, const float* const* weights, int layers, const int64_t* dims,
                                int b, float* out, float* stats, float* state) {
  dpf::Params P; fill(P, layers, dims, weights, x, adj, b);
  P.out = out; P.stats = stats; P.state = state;
  dpf::MemMap mp; const int nfl = dpf::build_map(P, mp);
  float* sm = (float*)calloc(nfl, 4);
  dpf::forward_body(P, sm, mp, 0, 1);
  free(sm);
  return 0;
}

extern "C" int emu_diffpool_bwd(const float* g_out, const float* coef, const float* x, const float* adj, const float* const* weights,
                                int layers, const int64_t* dims, int b, float* g_x, float* g_weights, float* state) {
  dpf::Params P; const int n = fill(P, layers, dims, weights, x, adj, b);
  P.g_out = g_out; P.coef = coef; P.g_x = g_x; P.partial = g_weights;   // one "CTA": its partial IS the result
  P.state = state;
  dpf::MemMap mp; const int nfl = dpf::build_map(P, mp);
  float* sm = (float*)calloc(nfl, 4);
  dpf::backward_body(P, sm, mp, 0, 1);
  free(sm);
  return n;
}
