// GENConv message + aggregation + MsgNorm/residual epilogue, forward and backward (sm_100a).
//
// Replaces, in one pass over the target-sorted CSR, the ~10 torch_scatter/ATen kernels behind
//   GENConv.message              models/gcn_lib/sparse/torch_vertex.py:94-101
//   GenMessagePassing.aggregate  models/gcn_lib/sparse/torch_message.py:44-85
//   MsgNorm.forward              models/gcn_lib/sparse/torch_message.py:175-179
//   h = x + m                    models/gcn_lib/sparse/torch_vertex.py:89
//
// Mapping: a group of LANES lanes (8/16/32) owns one target row; every lane owns VEC=4 consecutive
// channels (128-bit loads), so one group covers a chunk of 4*LANES channels and loops over chunks
// for wider rows.  Column / edge ids of a row are fetched coalesced by the group and broadcast with
// shuffles; UN edges' source rows and edge-feature rows are requested before any is consumed.
// Softmax is evaluated online with a LAZY running max (refreshed only when an element exceeds it by
// 2^24), so the common path per element is  add, max, add, fma, ex2, add, fma.
// HBM-bound: algorithmic bytes fwd = 4H(E + 2N) + 4E + 4(N+1)   (SURVEY.md section 8d).
#include <type_traits>

#include "common.cuh"
#include "../../include/mlg_b200.h"

#ifndef MLG_GEN_FWD_MIN_BLOCKS
#define MLG_GEN_FWD_MIN_BLOCKS 3
#endif
#ifndef MLG_GEN_UN
#define MLG_GEN_UN 4
#endif

namespace {

struct GenP {
  const float* x;
  const float* e;
  const float* ea;   // S_XA: per-edge scalar a_e [E] ...
  const float* ep;   // ... and the vectors p, q [H]: edge term e_ij = a_e * p + q (never materialised)
  const float* eq;
  float* pq_part;    // S_XA backward, optional: per-block partial sums [grid][2][H] of (a_e * g_edge, g_edge) over the edges
  const int* rowptr;
  const int* col;
  const int* eid;
  int n;
  unsigned H;
  int mode, learn, epi;
  float t, p, eps;
  const float* t_dev;
  const float* p_dev;
  const float* y_dev;
  const float* scale_dev;
  float* m;
  float* aux;
  float* h;
  // backward
  const float* g;
  const float* m_in;
  const float* aux_in;
  float* g_edge;     // ring backward with src_sum: may be nullptr (edge gradients not wanted)
  float* g_x;
  float* partials;
  int src_sum;       // ring backward: also add every edge gradient into g_x[source] (red.global.add.v4.f32; g_x zeroed first)
};

constexpr int kThreads = 256;
constexpr int UN = MLG_GEN_UN;
constexpr float kLo = 1e-7f, kHi = 1e1f;
constexpr float kLazy = 24.f;  // log2 head-room before the running max is refreshed

// kernel families
constexpr int K_SOFTMAX = 0, K_POWER = 1, K_SIMPLE = 2;  // SIMPLE: add / mean / max chosen at run time
// message sources
constexpr int S_XE = 0, S_X = 1, S_RAW = 2, S_XA = 3;  // relu(x_j + e) + eps | relu(x_j) + eps | e (given messages) |
                                                        // relu(x_j + a_e * p + q) + eps (rank-1 affine edge term)

template <int VEC>
struct Vec {
  float v[VEC];
};

template <int VEC, bool FULL>
__device__ __forceinline__ Vec<VEC> ld_g(const float* p, bool ok) {
  Vec<VEC> r;
  if (VEC == 4) {
    float4 t = (FULL || ok) ? ld_gather4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.v[0] = t.x; r.v[1 % VEC] = t.y; r.v[2 % VEC] = t.z; r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = (FULL || ok) ? __ldg(p) : 0.f;
  }
  return r;
}
template <int VEC, bool FULL>
__device__ __forceinline__ Vec<VEC> ld_s(const float* p, bool ok) {
  Vec<VEC> r;
  if (VEC == 4) {
    float4 t = (FULL || ok) ? ld_stream4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.v[0] = t.x; r.v[1 % VEC] = t.y; r.v[2 % VEC] = t.z; r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = (FULL || ok) ? __ldg(p) : 0.f;
  }
  return r;
}
template <int VEC, bool FULL>
__device__ __forceinline__ void st_v(float* p, const Vec<VEC>& r, bool ok, bool stream) {
  if (!FULL && !ok) return;
  if (VEC == 4) {
    float4 t = make_float4(r.v[0], r.v[1 % VEC], r.v[2 % VEC], r.v[3 % VEC]);
    if (stream) st_stream4(p, t); else st4(p, t);
  } else {
    *p = r.v[0];
  }
}

__device__ __forceinline__ float sigmoidf_(float y) { return 1.f / (1.f + expf(-y)); }
// row r of a row-major [*, H] fp32 matrix: one IMAD.WIDE.U32 + 64-bit add
__device__ __forceinline__ const float* row_ptr(const float* base, unsigned r, unsigned H) {
  return base + (size_t)r * H;
}
__device__ __forceinline__ float* row_ptr(float* base, unsigned r, unsigned H) { return base + (size_t)r * H; }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int LANES, int VEC, int KIND, int SRC, bool FULL>
__global__ void __launch_bounds__(kThreads, MLG_GEN_FWD_MIN_BLOCKS) gen_fwd_kernel(const GenP P) {
  constexpr int RPW = 32 / LANES;
  constexpr int CW = LANES * VEC;  // channels per chunk
  constexpr bool HAS_X = SRC != S_RAW, HAS_E = (SRC == S_XE || SRC == S_RAW), RAW = SRC == S_RAW, AFF = SRC == S_XA;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, sl = lane % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long row = warp * RPW + sub;
  if (row >= P.n) return;  // group-uniform exit; no block-level sync in this kernel

  const unsigned H = P.H;
  const int mode = P.mode;
  const int beg = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
  const int deg = end - beg;
  const float t = P.t_dev ? __ldg(P.t_dev) : P.t;
  const float pw = P.p_dev ? __ldg(P.p_dev) : P.p;
  const float tl2 = t * MLG_LOG2E;
  const float eps = P.eps;
  float degpow = 1.f;
  if (P.y_dev) degpow = powf((float)deg, sigmoidf_(__ldg(P.y_dev)));
  const int nchunks = (H + CW - 1) / CW;
  const float* xrow = HAS_X ? row_ptr(P.x, (unsigned)row, H) : nullptr;
  float* mrow = row_ptr(P.m, (unsigned)row, H);
  const bool has_eid = (HAS_E || AFF) && P.eid != nullptr;

  float sx2 = 0.f, sm2 = 0.f;
  Vec<VEC> out, xi;
#pragma unroll
  for (int k = 0; k < VEC; ++k) out.v[k] = xi.v[k] = 0.f;

  for (int ch = 0; ch < nchunks; ++ch) {
    const unsigned c = ch * CW + sl * VEC;
    const bool cok = FULL || c < H;
    const float* xc = HAS_X ? P.x + c : nullptr;
    const float* ec = HAS_E ? P.e + c : nullptr;
    Vec<VEC> pv, qv;
    if (AFF) {
      pv = ld_g<VEC, FULL>(P.ep + c, cok);
      qv = ld_g<VEC, FULL>(P.eq + c, cok);
    }
    float a0[VEC], a1[VEC], a2[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      a0[k] = (KIND == K_SOFTMAX || (KIND == K_SIMPLE && mode == MLG_AGGR_MAX)) ? -INFINITY : 0.f;
      a1[k] = 0.f;
      a2[k] = 0.f;
    }
    for (int base = beg; base < end; base += LANES) {
      const int q = min(base + sl, end - 1);
      const unsigned my_col = HAS_X ? (unsigned)__ldg(P.col + q) : 0u;
      const unsigned my_e = has_eid ? (unsigned)__ldg(P.eid + q) : (unsigned)q;  // identity if pre-sorted
      const float my_a = AFF ? __ldg(P.ea + my_e) : 0.f;
      const int cnt = min(LANES, end - base);
      // NE edges per step: all 2*NE row loads are issued before any is consumed
      auto step = [&](auto ne_tag, int j) {
        constexpr int NE = decltype(ne_tag)::value;
        Vec<VEC> xv[NE], ev[NE];
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          if (HAS_X) {
            const unsigned s = __shfl_sync(gmask, my_col, j + u, LANES);
            xv[u] = ld_g<VEC, FULL>(row_ptr(xc, s, H), cok);
          }
          if (AFF) {
            const float a = __shfl_sync(gmask, my_a, j + u, LANES);
#pragma unroll
            for (int k = 0; k < VEC; ++k) ev[u].v[k] = fmaf(a, pv.v[k], qv.v[k]);
          }
          if (HAS_E) {
            const unsigned ee = __shfl_sync(gmask, my_e, j + u, LANES);
            ev[u] = ld_s<VEC, FULL>(row_ptr(ec, ee, H), cok);
          }
        }
#pragma unroll
        for (int u = 0; u < NE; ++u) {
          float v[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            if (RAW) {
              v[k] = ev[u].v[k];
            } else {
              const float pre = (HAS_E || AFF) ? xv[u].v[k] + ev[u].v[k] : xv[u].v[k];
              v[k] = fmaxf(pre, 0.f) + eps;
            }
          }
          if (KIND == K_SOFTMAX) {
            float d[VEC];
            float dm = -INFINITY;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              d[k] = fmaf(v[k], tl2, -a0[k]);
              dm = fmaxf(dm, d[k]);
            }
            if (dm > kLazy) {  // rare (always on a row's first edge, where a0 = -inf)
#pragma unroll
              for (int k = 0; k < VEC; ++k) {
                const bool up = d[k] > 0.f;
                const float sc = ex2_approx(up ? -d[k] : 0.f);  // 2^(old_max - new_max); 0 on the first edge
                a1[k] *= sc;
                a2[k] *= sc;
                a0[k] = up ? v[k] * tl2 : a0[k];
                d[k] = up ? 0.f : d[k];
              }
            }
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float pz = ex2_approx(d[k]);
              a1[k] += pz;
              a2[k] = fmaf(v[k], pz, a2[k]);
            }
          } else if (KIND == K_POWER) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float vc = fminf(fmaxf(v[k], kLo), kHi);
              a0[k] += (pw == 1.f) ? vc : powf(vc, pw);
            }
          } else {
#pragma unroll
            for (int k = 0; k < VEC; ++k) a0[k] = (mode == MLG_AGGR_MAX) ? fmaxf(a0[k], v[k]) : a0[k] + v[k];
          }
        }
      };
      int j = 0;
      for (; j + UN <= cnt; j += UN) step(std::integral_constant<int, UN>{}, j);
      for (; j < cnt; ++j) step(std::integral_constant<int, 1>{}, j);
    }
    // finalise this chunk
    Vec<VEC> o, ax;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float r = 0.f, au = 0.f;
      if (KIND == K_SOFTMAX) {
        if (deg > 0) {
          r = a2[k] / a1[k];
          au = a0[k] + log2f(a1[k]);
        }
      } else if (KIND == K_POWER) {
        const float mean = a0[k] / (float)max(deg, 1);
        au = mean;
        const float cl = fminf(fmaxf(mean, kLo), kHi);
        r = (pw == 1.f) ? cl : powf(cl, 1.f / pw);
      } else if (mode == MLG_AGGR_MAX) {
        r = deg > 0 ? a0[k] : 0.f;
      } else if (mode == MLG_AGGR_MEAN) {
        r = a0[k] / (float)max(deg, 1);
      } else {
        r = a0[k];
      }
      if (P.y_dev) r *= degpow;
      o.v[k] = r;
      ax.v[k] = au;
    }
    if (P.m) st_v<VEC, FULL>(mrow + c, o, cok, false);
    if (P.aux) st_v<VEC, FULL>(row_ptr(P.aux, (unsigned)row, H) + c, ax, cok, true);
    if (P.epi != MLG_EPI_NONE) {
      Vec<VEC> xr = ld_g<VEC, FULL>(xrow + c, cok);
      if (P.epi == MLG_EPI_RESIDUAL) {
        Vec<VEC> hv;
#pragma unroll
        for (int k = 0; k < VEC; ++k) hv.v[k] = xr.v[k] + o.v[k];
        st_v<VEC, FULL>(row_ptr(P.h, (unsigned)row, H) + c, hv, cok, false);
      } else {
        if (cok) {  // lanes past H hold eps-valued garbage: keep them out of the norms
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            sx2 = fmaf(xr.v[k], xr.v[k], sx2);
            sm2 = fmaf(o.v[k], o.v[k], sm2);
          }
        }
        out = o;
        xi = xr;
      }
    }
  }

  if (P.epi == MLG_EPI_MSGNORM) {
    sx2 = group_sum<LANES>(sx2, gmask);
    sm2 = group_sum<LANES>(sm2, gmask);
    const float r = sqrtf(sx2);
    const float nm = fmaxf(sqrtf(sm2), 1e-12f);
    const float f = __ldg(P.scale_dev) * r / nm;
    float* hrow = row_ptr(P.h, (unsigned)row, H);
    for (int ch = 0; ch < nchunks; ++ch) {
      const unsigned c = ch * CW + sl * VEC;
      const bool cok = FULL || c < H;
      Vec<VEC> mo, xr;
      if (nchunks == 1) {
        mo = out;
        xr = xi;
      } else {
        // re-read this lane's own writes of m (same thread wrote them: visible without a fence)
#pragma unroll
        for (int k = 0; k < VEC; ++k) mo.v[k] = cok ? mrow[c + k] : 0.f;
        xr = ld_g<VEC, FULL>(xrow + c, cok);
      }
      Vec<VEC> hv;
#pragma unroll
      for (int k = 0; k < VEC; ++k) hv.v[k] = fmaf(f, mo.v[k], xr.v[k]);
      st_v<VEC, FULL>(hrow + c, hv, cok, false);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int LANES, int VEC, int KIND, int SRC, bool FULL>
__global__ void __launch_bounds__(kThreads) gen_bwd_kernel(const GenP P) {
  constexpr int RPW = 32 / LANES;
  constexpr int CW = LANES * VEC;
  constexpr bool HAS_X = SRC != S_RAW, HAS_E = (SRC == S_XE || SRC == S_RAW), RAW = SRC == S_RAW, AFF = SRC == S_XA;
  __shared__ float red[3 * 32];
  extern __shared__ __align__(16) float pq_stage[];   // AFF && pq_part: [lane groups][2][H] per-row edge-gradient sums
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, sl = lane % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long row = warp * RPW + sub;
  const bool active = row < P.n;
  const bool pq_on = AFF && P.pq_part != nullptr;
  constexpr int G = kThreads / LANES;
  if (pq_on) {
    for (unsigned j = threadIdx.x; j < G * 2 * P.H; j += kThreads) pq_stage[j] = 0.f;
    __syncthreads();
  }
  float* my_stage = pq_stage + (size_t)(threadIdx.x / LANES) * 2 * P.H;

  float acc[3] = {0.f, 0.f, 0.f};  // d/dt (or d/dp), d/dy_raw, d/dmsg_scale

  if (active) {
    const unsigned H = P.H;
    const int mode = P.mode;
    const bool learn = P.learn != 0;
    const int beg = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
    const int deg = end - beg;
    const float t = P.t_dev ? __ldg(P.t_dev) : P.t;
    const float pw = P.p_dev ? __ldg(P.p_dev) : P.p;
    const float tl2 = t * MLG_LOG2E;
    const float eps = P.eps;
    float degpow = 1.f, ycoef = 0.f;
    if (P.y_dev) {
      const float sg = sigmoidf_(__ldg(P.y_dev));
      degpow = powf((float)deg, sg);
      ycoef = deg > 0 ? logf((float)deg) * sg * (1.f - sg) : 0.f;
    }
    const float inv_deg = 1.f / (float)max(deg, 1);
    const int nchunks = (H + CW - 1) / CW;
    const float* xrow = HAS_X ? row_ptr(P.x, (unsigned)row, H) : nullptr;
    const float* mrow = row_ptr(P.m_in, (unsigned)row, H);
    const float* grow = row_ptr(P.g, (unsigned)row, H);
    const bool has_eid = P.eid != nullptr;

    // --- MsgNorm statistics of this row (one sweep over channels) ---
    float f_gm = 1.f, f_u = 0.f, f_x = 0.f;  // g_m = f_gm * g - f_u * m ; g_x = g + f_x * x
    if (P.epi == MLG_EPI_MSGNORM) {
      float sx2 = 0.f, sm2 = 0.f, sgm = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        const unsigned c = ch * CW + sl * VEC;
        const bool cok = FULL || c < H;
        Vec<VEC> xr = ld_g<VEC, FULL>(xrow + c, cok), mr = ld_g<VEC, FULL>(mrow + c, cok),
                 gr = ld_g<VEC, FULL>(grow + c, cok);
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          sx2 = fmaf(xr.v[k], xr.v[k], sx2);
          sm2 = fmaf(mr.v[k], mr.v[k], sm2);
          sgm = fmaf(gr.v[k], mr.v[k], sgm);
        }
      }
      sx2 = group_sum<LANES>(sx2, gmask);
      sm2 = group_sum<LANES>(sm2, gmask);
      sgm = group_sum<LANES>(sgm, gmask);
      const float s = __ldg(P.scale_dev);
      const float r = sqrtf(sx2), nm = sqrtf(sm2);
      const float nn = fmaxf(nm, 1e-12f);
      const float dot = sgm / nn;  // g . u
      f_gm = s * r / nn;
      f_u = (nm >= 1e-12f) ? f_gm * dot / nn : 0.f;  // clamp_min passes grad iff norm >= eps
      f_x = (r > 0.f) ? s * dot / r : 0.f;
      if (sl == 0) acc[2] = dot * r;
    }

    for (int ch = 0; ch < nchunks; ++ch) {
      const unsigned c = ch * CW + sl * VEC;
      const bool cok = FULL || c < H;
      const float* xc = HAS_X ? P.x + c : nullptr;
      const float* ec = HAS_E ? P.e + c : nullptr;
      Vec<VEC> pv, qv;
      if (AFF) {
        pv = ld_g<VEC, FULL>(P.ep + c, cok);
        qv = ld_g<VEC, FULL>(P.eq + c, cok);
      }
      float* gec = P.g_edge + c;
      Vec<VEC> gr = ld_g<VEC, FULL>(grow + c, cok);
      Vec<VEC> mr = ld_g<VEC, FULL>(mrow + c, cok);
      Vec<VEC> au;
#pragma unroll
      for (int k = 0; k < VEC; ++k) au.v[k] = 0.f;
      if (KIND != K_SIMPLE) au = ld_g<VEC, FULL>(row_ptr(P.aux_in, (unsigned)row, H) + c, cok);
      float gin[VEC], oi[VEC], k1[VEC], k2[VEC];
      bool taken[VEC];
      // direct term into g_x and gradient w.r.t. the aggregated message
      {
        Vec<VEC> gx;
        if (P.epi == MLG_EPI_NONE) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) gx.v[k] = 0.f;
        } else if (P.epi == MLG_EPI_RESIDUAL) {
          gx = gr;
        } else {
          Vec<VEC> xr = ld_g<VEC, FULL>(xrow + c, cok);
#pragma unroll
          for (int k = 0; k < VEC; ++k) gx.v[k] = fmaf(f_x, xr.v[k], gr.v[k]);
        }
        st_v<VEC, FULL>(row_ptr(P.g_x, (unsigned)row, H) + c, gx, cok, false);
      }
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float gm = (P.epi == MLG_EPI_MSGNORM) ? (f_gm * gr.v[k] - f_u * mr.v[k]) : gr.v[k];
        if (P.y_dev && cok) acc[1] = fmaf(gm * mr.v[k], ycoef, acc[1]);
        gin[k] = gm * degpow;
        oi[k] = (P.y_dev && deg > 0) ? mr.v[k] / degpow : mr.v[k];
        k1[k] = 0.f;
        k2[k] = 0.f;
        taken[k] = false;
        if (KIND == K_POWER) {
          const float mean = au.v[k];
          const bool inr = mean >= kLo && mean <= kHi;
          const float cl = fminf(fmaxf(mean, kLo), kHi);
          // d out / d v_e = inr * (oi/cl) * vc^(p-1) / deg
          k1[k] = inr ? gin[k] * oi[k] / cl * inv_deg : 0.f;
          if (learn && cok) {
            acc[0] = fmaf(gin[k] * oi[k], -logf(cl) / (pw * pw), acc[0]);
            k2[k] = inr ? gin[k] * oi[k] / (pw * cl) * inv_deg : 0.f;
          }
        } else if (KIND == K_SIMPLE && mode == MLG_AGGR_MEAN) {
          k1[k] = gin[k] * inv_deg;
        }
      }

      float su[VEC], sv[VEC];   // AFF: sum over this row's edges of a_e * g_edge and g_edge (this lane's channels)
#pragma unroll
      for (int k = 0; k < VEC; ++k) su[k] = sv[k] = 0.f;
      for (int base = beg; base < end; base += LANES) {
        const int q = min(base + sl, end - 1);
        const unsigned my_col = HAS_X ? (unsigned)__ldg(P.col + q) : 0u;
        const unsigned my_e = has_eid ? (unsigned)__ldg(P.eid + q) : (unsigned)q;
        const float my_a = AFF ? __ldg(P.ea + my_e) : 0.f;
        const int cnt = min(LANES, end - base);
        auto step = [&](auto ne_tag, int j) {
          constexpr int NE = decltype(ne_tag)::value;
          Vec<VEC> xv[NE], ev[NE];
          unsigned eo[NE];
          float av[NE];
#pragma unroll
          for (int u = 0; u < NE; ++u) {
            eo[u] = __shfl_sync(gmask, my_e, j + u, LANES);
            av[u] = 0.f;
            if (AFF) {
              const float a = __shfl_sync(gmask, my_a, j + u, LANES);
              av[u] = a;
#pragma unroll
              for (int k = 0; k < VEC; ++k) ev[u].v[k] = fmaf(a, pv.v[k], qv.v[k]);
            }
            if (HAS_X) {
              const unsigned s = __shfl_sync(gmask, my_col, j + u, LANES);
              xv[u] = ld_g<VEC, FULL>(row_ptr(xc, s, H), cok);
            }
            if (HAS_E) ev[u] = ld_s<VEC, FULL>(row_ptr(ec, eo[u], H), cok);
          }
#pragma unroll
          for (int u = 0; u < NE; ++u) {
            Vec<VEC> ge;
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              float pre, v;
              if (RAW) {
                pre = v = ev[u].v[k];
              } else {
                pre = (HAS_E || AFF) ? xv[u].v[k] + ev[u].v[k] : xv[u].v[k];
                v = fmaxf(pre, 0.f) + eps;
              }
              float gv;
              if (KIND == K_SOFTMAX) {
                const float w = ex2_approx(fmaf(v, tl2, -au.v[k]));
                const float gw = gin[k] * w;
                if (learn) {
                  const float dv = v - oi[k];
                  gv = gw * fmaf(t, dv, 1.f);
                  if (cok) acc[0] = fmaf(gw * v, dv, acc[0]);
                } else {
                  gv = gw;
                }
              } else if (KIND == K_POWER) {
                const float vc = fminf(fmaxf(v, kLo), kHi);
                const float vp1 = (pw == 1.f) ? 1.f : powf(vc, pw - 1.f);
                gv = (v <= kHi) ? k1[k] * vp1 : 0.f;
                if (learn && cok) acc[0] = fmaf(k2[k] * vp1 * vc, logf(vc), acc[0]);
              } else if (mode == MLG_AGGR_MAX) {
                const bool hit = (!taken[k]) && (v == oi[k]);
                gv = hit ? gin[k] : 0.f;
                taken[k] = taken[k] || hit;
              } else if (mode == MLG_AGGR_MEAN) {
                gv = k1[k];
              } else {
                gv = gin[k];
              }
              ge.v[k] = (RAW || pre > 0.f) ? gv : 0.f;
              if (AFF) {
                su[k] = fmaf(av[u], ge.v[k], su[k]);
                sv[k] += ge.v[k];
              }
            }
            st_v<VEC, FULL>(row_ptr(gec, eo[u], H), ge, cok, true);
          }
        };
        int j = 0;
        for (; j + UN <= cnt; j += UN) step(std::integral_constant<int, UN>{}, j);
        for (; j < cnt; ++j) step(std::integral_constant<int, 1>{}, j);
      }
      if (pq_on && cok) {
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          my_stage[c + k] = su[k];
          my_stage[H + c + k] = sv[k];
        }
      }
    }
  }

  if (pq_on) {   // this block's rows, added in lane-group order
    __syncthreads();
    const unsigned H2 = 2 * P.H;
    for (unsigned j = threadIdx.x; j < H2; j += kThreads) {
      float sum = 0.f;
#pragma unroll 4
      for (int gi = 0; gi < G; ++gi) sum += pq_stage[(size_t)gi * H2 + j];
      P.pq_part[(size_t)blockIdx.x * H2 + j] = sum;
    }
  }

  block_sum<3>(acc, red);
  if (threadIdx.x == 0) {
    float* o = P.partials + (size_t)blockIdx.x * 4;
    o[0] = acc[0];
    o[1] = acc[1];
    o[2] = acc[2];
    o[3] = 0.f;
  }
}

// ------------------------------------------------------------------------------------------------
// forward, shared-memory staged variant (softmax family, x + e messages, H = 128 * NV):
// every warp owns RPW consecutive target rows and walks their edges as ONE flat stream; each edge's
// source row and edge-feature row are copied global -> shared with 16-byte cp.async (one commit group
// per edge, kDepth edges in flight per warp, no registers held), each lane reading back exactly the 16
// bytes it copied (no barrier needed, cp.async.wait_group only).  The edge stream carries an L2
// evict-first policy so that it does not push the re-used node rows out of L2.
// ------------------------------------------------------------------------------------------------
#ifndef MLG_RING_WARPS
#define MLG_RING_WARPS 4
#endif
#ifndef MLG_RING_RPW
#define MLG_RING_RPW 8
#endif
#ifndef MLG_RING_DEPTH
#define MLG_RING_DEPTH 8
#endif
constexpr int kRingWarps = MLG_RING_WARPS;   // warps per block
constexpr int kRPW = MLG_RING_RPW;           // rows per warp
constexpr int kDepth = MLG_RING_DEPTH;       // edges in flight per warp (power of two)

__device__ __forceinline__ void cp_async16_pol(unsigned dst, const void* src, unsigned long long pol) {
  asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void cp_async16_plain(unsigned dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float4 lds128(unsigned addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

constexpr int kGroup = 4;                       // edges per commit group
constexpr int kGroups = kDepth / kGroup;        // groups in flight per warp

template <int NV>
__global__ void __launch_bounds__(kRingWarps * 32) gen_fwd_ring_kernel(const GenP P) {
  extern __shared__ __align__(16) unsigned char ring_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * kRingWarps + wib;
  const long long r0 = gw * kRPW;
  if (r0 >= P.n) return;
  const int r1 = (int)min((long long)P.n, r0 + kRPW);
  const unsigned H = P.H;
  constexpr unsigned kSlotBytes = 2 * NV * 512;   // x chunk(s) + e chunk(s) of one edge
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(ring_raw) + wib * kDepth * kSlotBytes + lane * 16;
  unsigned long long pol_stream;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));

  // row pointers of this warp's rows: lane j holds rowptr[r0 + j] (j <= RPW)
  const int rp = __ldg(P.rowptr + min((long long)P.n, r0 + min(lane, kRPW)));
  const int qb = __shfl_sync(0xffffffffu, rp, 0);
  const int qe = __shfl_sync(0xffffffffu, rp, r1 - (int)r0);
  const int* eidp = P.eid;
  const float t = P.t_dev ? __ldg(P.t_dev) : P.t;
  const float tl2 = t * MLG_LOG2E;
  const float eps = P.eps;
  const float ysig = P.y_dev ? sigmoidf_(__ldg(P.y_dev)) : 0.f;
  const float scale = (P.epi == MLG_EPI_MSGNORM) ? __ldg(P.scale_dev) : 0.f;
  const float* xc = P.x + lane * 4;
  const float* ec = P.e + lane * 4;
  const int n_groups = (qe - qb + kGroup - 1) / kGroup;

  // ---- issue side: one commit group = kGroup edges; the group's source / edge ids are fetched (uniform,
  //      L1-broadcast loads) one group ahead of their use so the address chain never stalls the issue ----
  unsigned nxt_s[kGroup], nxt_e[kGroup];
  auto fetch_ids = [&](int g) {
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int p = min(qb + g * kGroup + u, qe - 1);
      nxt_s[u] = (unsigned)__ldg(P.col + p);
      nxt_e[u] = eidp ? (unsigned)__ldg(eidp + p) : (unsigned)p;
    }
  };
  auto issue = [&](int g) {   // uses the ids fetched for group g, then fetches those of group g + 1
    if (g < n_groups) {
      const unsigned slot0 = sbase + (unsigned)((g % kGroups) * kGroup) * kSlotBytes;
#pragma unroll
      for (int u = 0; u < kGroup; ++u) {
        const unsigned slot = slot0 + u * kSlotBytes;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          cp_async16_plain(slot + v * 512, row_ptr(xc, nxt_s[u], H) + v * 128);
          cp_async16_pol(slot + (NV + v) * 512, row_ptr(ec, nxt_e[u], H) + v * 128, pol_stream);
        }
      }
      if (g + 1 < n_groups) fetch_ids(g + 1);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  if (n_groups > 0) fetch_ids(0);
#pragma unroll
  for (int i = 0; i < kGroups - 1; ++i) issue(i);

  // ---- consume side ----
  int row = (int)r0;
  int rend = __shfl_sync(0xffffffffu, rp, 1);
  int rbeg = qb;
  float a0[NV][4], a1[NV][4], a2[NV][4];
  float4 xi[NV];
  auto reset = [&]() {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
#pragma unroll
      for (int k = 0; k < 4; ++k) { a0[v][k] = -INFINITY; a1[v][k] = 0.f; a2[v][k] = 0.f; }
      xi[v] = (P.epi != MLG_EPI_NONE) ? ld_gather4(row_ptr(xc, (unsigned)row, H) + v * 128) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto finalize = [&]() {
    const int deg = rend - rbeg;
    const float degpow = P.y_dev ? powf((float)deg, ysig) : 1.f;
    float o[NV][4];
    float sx2 = 0.f, sm2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      float au[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float r = 0.f;
        au[k] = 0.f;
        if (deg > 0) {
          r = a2[v][k] / a1[v][k];
          au[k] = a0[v][k] + log2f(a1[v][k]);
        }
        o[v][k] = r * degpow;
        sm2 = fmaf(o[v][k], o[v][k], sm2);
      }
      sx2 += xi[v].x * xi[v].x + xi[v].y * xi[v].y + xi[v].z * xi[v].z + xi[v].w * xi[v].w;
      const unsigned off = lane * 4 + v * 128;
      if (P.m) st4(row_ptr(P.m, (unsigned)row, H) + off, make_float4(o[v][0], o[v][1], o[v][2], o[v][3]));
      if (P.aux) st_stream4(row_ptr(P.aux, (unsigned)row, H) + off, make_float4(au[0], au[1], au[2], au[3]));
    }
    if (P.epi != MLG_EPI_NONE) {
      float f = 1.f;
      if (P.epi == MLG_EPI_MSGNORM) {
        sx2 = warp_sum(sx2);
        sm2 = warp_sum(sm2);
        f = scale * sqrtf(sx2) / fmaxf(sqrtf(sm2), 1e-12f);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const unsigned off = lane * 4 + v * 128;
        st4(row_ptr(P.h, (unsigned)row, H) + off,
            make_float4(fmaf(f, o[v][0], xi[v].x), fmaf(f, o[v][1], xi[v].y), fmaf(f, o[v][2], xi[v].z),
                        fmaf(f, o[v][3], xi[v].w)));
      }
    }
  };
  reset();
  for (int g = 0; g < n_groups; ++g) {
    issue(g + kGroups - 1);
    asm volatile("cp.async.wait_group %0;" ::"n"(kGroups - 1) : "memory");
    const unsigned slot0 = sbase + (unsigned)((g % kGroups) * kGroup) * kSlotBytes;
    const int q0 = qb + g * kGroup;
#pragma unroll
    for (int u = 0; u < kGroup; ++u) {
      const int q = q0 + u;
      if (q >= qe) break;
      while (q >= rend) {  // row finished (also skips empty rows)
        finalize();
        ++row;
        rbeg = rend;
        rend = __shfl_sync(0xffffffffu, rp, row - (int)r0 + 1);
        reset();
      }
      const unsigned slot = slot0 + u * kSlotBytes;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 xv = lds128(slot + v * 512);
        const float4 ev = lds128(slot + (NV + v) * 512);
        const float vv[4] = {fmaxf(xv.x + ev.x, 0.f) + eps, fmaxf(xv.y + ev.y, 0.f) + eps,
                             fmaxf(xv.z + ev.z, 0.f) + eps, fmaxf(xv.w + ev.w, 0.f) + eps};
        float d[4];
        float dm = -INFINITY;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          d[k] = fmaf(vv[k], tl2, -a0[v][k]);
          dm = fmaxf(dm, d[k]);
        }
        if (dm > kLazy) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool up = d[k] > 0.f;
            const float sc = ex2_approx(up ? -d[k] : 0.f);
            a1[v][k] *= sc;
            a2[v][k] *= sc;
            a0[v][k] = up ? vv[k] * tl2 : a0[v][k];
            d[k] = up ? 0.f : d[k];
          }
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float pz = ex2_approx(d[k]);
          a1[v][k] += pz;
          a2[v][k] = fmaf(vv[k], pz, a2[v][k]);
        }
      }
    }
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  while (row < r1) {  // last row with edges + trailing empty rows
    finalize();
    ++row;
    if (row < r1) {
      rbeg = rend;
      rend = __shfl_sync(0xffffffffu, rp, row - (int)r0 + 1);
      reset();
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward, shared-memory staged variant (softmax family, H = 128 * NV; AFF = false: x + e messages, AFF = true: the
// rank-1 affine edge term a_e * p + q): the same flat edge stream per warp as gen_fwd_ring_kernel -- source rows (and
// edge-feature rows) arrive by 16-byte cp.async, kDepth edges in flight per warp without holding registers -- instead of
// the register-staged gather of gen_bwd_kernel (UN rows per lane group in flight: 0.42 of HBM at N = 100 k, k = 16,
// H = 128).  Per row: MsgNorm statistics and the direct term into g_x first, then one pass over its edges writes
// g_edge (streaming stores).  Per-block partials as gen_bwd_kernel: (d/dt, d/dy_raw, d/dmsg_scale) and, for AFF with
// pq_part, the sums of (a_e * g_edge, g_edge) over the block's edges.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add4(float* p, float4 v, unsigned long long pol) {
  asm volatile("red.global.add.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ float4 ld_pol4(const float* p, unsigned long long pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ void st_pol4(float* p, float4 v, unsigned long long pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}

// P.src_sum: the source-side sum g_x[j] += sum over out-edges (j -> i) of g_edge is done HERE with 16-byte vector reductions
// into the L2-resident g_x (probe tools/probes/red_v4.cu: 1.6 M scattered 512-byte rows in 0.144 ms = 5.7 TB/s of payload,
// the speed of a plain read-modify-write) instead of a second kernel that reads all of g_edge [E, H] again; with the affine
// edge term g_edge is then not written at all.  The order of the additions is not reproducible (as in the reference's own
// scatter-add backward); the two-pass path stays available (functional.GEN_BWD_SRC_ATOMIC = False).
// L2 residency: x (gathered) and g_x (reduced into) are the re-used 2 x 4 n H bytes; everything else -- e, g_edge, and the
// g / m / aux rows -- passes through once.  The first capture of this kernel (profiles/r02_genbwd_raw.csv: 1.75 GB read +
// 1.18 GB written against 1.07 + 0.87 GB algorithmic) showed the streams evicting g_x / x lines, every reduction on an
// evicted line costing a DRAM fill and a write-back; hence evict_last on the gathers / reductions and evict_first on all
// streams.
// The per-edge loop is instruction-issue bound (first capture, profiles/r02_genaffb_raw.csv: 162 warp instructions per
// edge, 68 % issue-active, DRAM at 1 TB/s): the element-wise chain runs on PAIRS of channels with the packed fp32
// instructions of sm_100 (fma / add / mul .rn.f32x2 = FFMA2 / FADD2 / FMUL2: one issue slot for two IEEE-identical
// results), the run-time switches (source-side sum, g_edge wanted, learn_t) are template parameters, and the edge stream
// is walked as plain nested loops (rows, then the row's entries; one group test per entry).
template <int NV, bool AFF, bool SRC, bool EDGE, bool LEARN>
__global__ void __launch_bounds__(kRingWarps * 32) gen_bwd_ring_kernel(const GenP P) {
  extern __shared__ __align__(16) unsigned char ring_raw[];
  __shared__ float red[3 * 32];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const long long gw = (long long)blockIdx.x * kRingWarps + wib;
  const long long r0 = gw * kRPW;
  const unsigned H = P.H;
  constexpr unsigned kSlotBytes = (AFF ? 1 : 2) * NV * 512;
  const unsigned sbase = (unsigned)__cvta_generic_to_shared(ring_raw) + wib * kDepth * kSlotBytes + lane * 16;
  float acc[3] = {0.f, 0.f, 0.f};
  u64 acc0p = 0ull;              // learn_t: (even, odd) channel halves of the d/dt sum
  u64 su2[NV][2], sv2[NV][2];    // AFF: sums over this warp's edges of a_e * g_edge and g_edge
#pragma unroll
  for (int v = 0; v < NV; ++v) su2[v][0] = su2[v][1] = sv2[v][0] = sv2[v][1] = 0ull;

  if (r0 < P.n) {
    const int r1 = (int)min((long long)P.n, r0 + kRPW);
    unsigned long long pol_stream, pol_keep;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    const int rp = __ldg(P.rowptr + min((long long)P.n, r0 + min(lane, kRPW)));
    const int qb = __shfl_sync(0xffffffffu, rp, 0);
    const int qe = __shfl_sync(0xffffffffu, rp, r1 - (int)r0);
    // kernel parameters used inside the edge loop, read once
    const int* const eidp = P.eid;
    const int* const colp = P.col;
    const float* const eap = P.ea;
    float* const gedge = P.g_edge;
    float* const gxp = P.g_x;
    const float* const gp = P.g;
    const float* const mp = P.m_in;
    const float* const auxp = P.aux_in;
    const int epi = P.epi;
    const bool has_y = P.y_dev != nullptr;
    const float t = P.t_dev ? __ldg(P.t_dev) : P.t;
    const float eps = P.eps;
    const u64 t2 = pk2(t, t), tl2 = pk2(t * MLG_LOG2E, t * MLG_LOG2E), eps2 = pk2(eps, eps), one2 = pk2(1.f, 1.f);
    float ysig = 0.f;
    if (has_y) ysig = sigmoidf_(__ldg(P.y_dev));
    const float msg_scale = epi == MLG_EPI_MSGNORM ? __ldg(P.scale_dev) : 0.f;
    const float* xc = P.x + lane * 4;
    const float* ec = AFF ? nullptr : P.e + lane * 4;
    u64 pv2[NV][2], qv2[NV][2];
    if (AFF) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const float4 pp = ld_gather4(P.ep + lane * 4 + v * 128), qq = ld_gather4(P.eq + lane * 4 + v * 128);
        pv2[v][0] = pk2(pp.x, pp.y); pv2[v][1] = pk2(pp.z, pp.w);
        qv2[v][0] = pk2(qq.x, qq.y); qv2[v][1] = pk2(qq.z, qq.w);
      }
    }
    const int n_groups = (qe - qb + kGroup - 1) / kGroup;

    unsigned nxt_s[kGroup], nxt_e[kGroup];
    auto fetch_ids = [&](int g) {
#pragma unroll
      for (int u = 0; u < kGroup; ++u) {
        const int q = min(qb + g * kGroup + u, qe - 1);
        nxt_s[u] = (unsigned)__ldg(colp + q);
        nxt_e[u] = eidp ? (unsigned)__ldg(eidp + q) : (unsigned)q;
      }
    };
    auto issue = [&](int g) {
      if (g < n_groups) {
        const unsigned slot0 = sbase + (unsigned)((g % kGroups) * kGroup) * kSlotBytes;
#pragma unroll
        for (int u = 0; u < kGroup; ++u) {
          const unsigned slot = slot0 + u * kSlotBytes;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            cp_async16_pol(slot + v * 512, row_ptr(xc, nxt_s[u], H) + v * 128, pol_keep);
            if (!AFF) cp_async16_pol(slot + (NV + v) * 512, row_ptr(ec, nxt_e[u], H) + v * 128, pol_stream);
          }
        }
        if (g + 1 < n_groups) fetch_ids(g + 1);
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    if (n_groups > 0) fetch_ids(0);
#pragma unroll
    for (int i = 0; i < kGroups - 1; ++i) issue(i);

    for (int row = (int)r0; row < r1; ++row) {
      const int rbeg = __shfl_sync(0xffffffffu, rp, row - (int)r0);
      const int rend = __shfl_sync(0xffffffffu, rp, row - (int)r0 + 1);
      // ---- per-row state: MsgNorm statistics, the direct term into g_x, and (gin, -oi, -aux) for the row's edges ----
      u64 gin2[NV][2], noi2[NV][2], nau2[NV][2];
      {
        const int deg = rend - rbeg;
        float degpow = 1.f, ycoef = 0.f;
        if (has_y) {
          degpow = powf((float)deg, ysig);
          ycoef = deg > 0 ? logf((float)deg) * ysig * (1.f - ysig) : 0.f;
        }
        float4 gr[NV], mr[NV], xr[NV], ar[NV];
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const unsigned off = lane * 4 + v * 128;
          gr[v] = ld_pol4(row_ptr(gp, (unsigned)row, H) + off, pol_stream);
          mr[v] = ld_pol4(row_ptr(mp, (unsigned)row, H) + off, pol_stream);
          ar[v] = ld_pol4(row_ptr(auxp, (unsigned)row, H) + off, pol_stream);
          xr[v] = (epi == MLG_EPI_MSGNORM) ? ld_gather4(row_ptr(P.x, (unsigned)row, H) + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        float f_gm = 1.f, f_u = 0.f, f_x = 0.f;   // g_m = f_gm * g - f_u * m ; g_x = g + f_x * x
        if (epi == MLG_EPI_MSGNORM) {
          float sx2 = 0.f, sm2 = 0.f, sgm = 0.f;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            sx2 += xr[v].x * xr[v].x + xr[v].y * xr[v].y + xr[v].z * xr[v].z + xr[v].w * xr[v].w;
            sm2 += mr[v].x * mr[v].x + mr[v].y * mr[v].y + mr[v].z * mr[v].z + mr[v].w * mr[v].w;
            sgm += gr[v].x * mr[v].x + gr[v].y * mr[v].y + gr[v].z * mr[v].z + gr[v].w * mr[v].w;
          }
          sx2 = warp_sum(sx2);
          sm2 = warp_sum(sm2);
          sgm = warp_sum(sgm);
          const float r = sqrtf(sx2), nm = sqrtf(sm2);
          const float nn = fmaxf(nm, 1e-12f);
          const float dot = sgm / nn;
          f_gm = msg_scale * r / nn;
          f_u = (nm >= 1e-12f) ? f_gm * dot / nn : 0.f;
          f_x = (r > 0.f) ? msg_scale * dot / r : 0.f;
          if (lane == 0) acc[2] += dot * r;
        }
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const unsigned off = lane * 4 + v * 128;
          const float g4[4] = {gr[v].x, gr[v].y, gr[v].z, gr[v].w};
          const float m4[4] = {mr[v].x, mr[v].y, mr[v].z, mr[v].w};
          const float x4[4] = {xr[v].x, xr[v].y, xr[v].z, xr[v].w};
          float gx[4], gi[4], oi[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            gx[k] = epi == MLG_EPI_NONE ? 0.f : (epi == MLG_EPI_RESIDUAL ? g4[k] : fmaf(f_x, x4[k], g4[k]));
            const float gm = (epi == MLG_EPI_MSGNORM) ? (f_gm * g4[k] - f_u * m4[k]) : g4[k];
            if (has_y) acc[1] = fmaf(gm * m4[k], ycoef, acc[1]);
            gi[k] = gm * degpow;
            oi[k] = (has_y && deg > 0) ? m4[k] / degpow : m4[k];
          }
          gin2[v][0] = pk2(gi[0], gi[1]); gin2[v][1] = pk2(gi[2], gi[3]);
          noi2[v][0] = pk2(-oi[0], -oi[1]); noi2[v][1] = pk2(-oi[2], -oi[3]);
          nau2[v][0] = pk2(-ar[v].x, -ar[v].y); nau2[v][1] = pk2(-ar[v].z, -ar[v].w);
          if (!SRC) st4(row_ptr(gxp, (unsigned)row, H) + off, make_float4(gx[0], gx[1], gx[2], gx[3]));
          else if (epi != MLG_EPI_NONE) red_add4(row_ptr(gxp, (unsigned)row, H) + off, make_float4(gx[0], gx[1], gx[2], gx[3]), pol_keep);
        }
      }
      // ---- the row's edges ----
      for (int q = rbeg; q < rend; ++q) {
        const int rel = q - qb;
        if ((rel & (kGroup - 1)) == 0) {
          issue(rel / kGroup + kGroups - 1);
          asm volatile("cp.async.wait_group %0;" ::"n"(kGroups - 1) : "memory");
        }
        const unsigned slot = sbase + (unsigned)(rel & (kDepth - 1)) * kSlotBytes;
        const unsigned eo = eidp ? (unsigned)__ldg(eidp + q) : (unsigned)q;   // uniform, L1-resident (fetched at issue time)
        const unsigned src = SRC ? (unsigned)__ldg(colp + q) : 0u;
        const float a_e = AFF ? __ldg(eap + eo) : 0.f;
        const u64 a2 = pk2(a_e, a_e);
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const float4 xv = lds128(slot + v * 512);
          float4 ev4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (!AFF) ev4 = lds128(slot + (NV + v) * 512);
          const u64 x2[2] = {pk2(xv.x, xv.y), pk2(xv.z, xv.w)};
          float ge[4];
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const u64 e2 = AFF ? fma2(a2, pv2[v][j], qv2[v][j]) : (j == 0 ? pk2(ev4.x, ev4.y) : pk2(ev4.z, ev4.w));
            const u64 pre2 = add2(x2[j], e2);
            float p0, p1;
            upk2(pre2, p0, p1);
            const u64 val2 = add2(pk2(fmaxf(p0, 0.f), fmaxf(p1, 0.f)), eps2);
            float a0, a1;
            upk2(fma2(val2, tl2, nau2[v][j]), a0, a1);
            const u64 gwt2 = mul2(gin2[v][j], pk2(ex2_approx(a0), ex2_approx(a1)));
            u64 gv2 = gwt2;
            if (LEARN) {
              const u64 dv2 = add2(val2, noi2[v][j]);
              gv2 = mul2(gwt2, fma2(t2, dv2, one2));
              acc0p = fma2(mul2(gwt2, val2), dv2, acc0p);
            }
            float g0, g1;
            upk2(gv2, g0, g1);
            ge[2 * j] = p0 > 0.f ? g0 : 0.f;
            ge[2 * j + 1] = p1 > 0.f ? g1 : 0.f;
            if (AFF) {
              const u64 ge2 = pk2(ge[2 * j], ge[2 * j + 1]);
              su2[v][j] = fma2(a2, ge2, su2[v][j]);
              sv2[v][j] = add2(sv2[v][j], ge2);
            }
          }
          const float4 ge4 = make_float4(ge[0], ge[1], ge[2], ge[3]);
          if (EDGE) st_pol4(row_ptr(gedge, eo, H) + lane * 4 + v * 128, ge4, pol_stream);
          if (SRC) red_add4(row_ptr(gxp, src, H) + lane * 4 + v * 128, ge4, pol_keep);
        }
      }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  if (LEARN) {
    float lo, hi;
    upk2(acc0p, lo, hi);
    acc[0] = lo + hi;
  }

  if (AFF && P.pq_part) {   // block sums in warp order: reuse the (now idle) ring as [warps][2][H]
    __syncthreads();
    float* stage = reinterpret_cast<float*>(ring_raw);
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        float s0, s1, t0, t1;
        upk2(su2[v][j], s0, s1);
        upk2(sv2[v][j], t0, t1);
        const size_t o = (size_t)wib * 2 * H + lane * 4 + v * 128 + 2 * j;
        stage[o] = s0; stage[o + 1] = s1;
        stage[o + H] = t0; stage[o + H + 1] = t1;
      }
    __syncthreads();
    for (unsigned j = threadIdx.x; j < 2 * H; j += kRingWarps * 32) {
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < kRingWarps; ++w) sum += stage[(size_t)w * 2 * H + j];
      P.pq_part[(size_t)blockIdx.x * 2 * H + j] = sum;
    }
  }
  block_sum<3>(acc, red);
  if (threadIdx.x == 0) {
    float* o = P.partials + (size_t)blockIdx.x * 4;
    o[0] = acc[0];
    o[1] = acc[1];
    o[2] = acc[2];
    o[3] = 0.f;
  }
}

// host side of the ring backward: true when it ran
inline bool ring_bwd_ok(const GenP& P) {
  return P.mode == MLG_AGGR_SOFTMAX && P.x && (P.e || P.ea) && (P.H == 128 || P.H == 256) && ((uintptr_t)P.x % 16 == 0) &&
         (!P.e || (uintptr_t)P.e % 16 == 0) && (P.g_edge || P.src_sum) && ((uintptr_t)P.g_edge % 16 == 0) &&
         ((uintptr_t)P.g % 16 == 0) &&
         ((uintptr_t)P.m_in % 16 == 0) && ((uintptr_t)P.aux_in % 16 == 0) && ((uintptr_t)P.g_x % 16 == 0) &&
         (!P.ea || ((uintptr_t)P.ep % 16 == 0 && (uintptr_t)P.eq % 16 == 0));
}
inline long long ring_grid(long long n) { return (n + kRingWarps * kRPW - 1) / (kRingWarps * kRPW); }

template <int NV, bool AFF, bool SRC, bool EDGE, bool LEARN>
int launch_ring_bwd_k(const GenP& P, cudaStream_t st) {
  int smem = kRingWarps * kDepth * (AFF ? 1 : 2) * NV * 512;
  const int stage = kRingWarps * 2 * (int)P.H * 4;      // the pq block sums reuse the ring
  if (AFF && P.pq_part && stage > smem) smem = stage;
  static bool attr = false;
  if (!attr) {
    MLG_CUDA(cudaFuncSetAttribute(gen_bwd_ring_kernel<NV, AFF, SRC, EDGE, LEARN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  64 * 1024));
    attr = true;
  }
  gen_bwd_ring_kernel<NV, AFF, SRC, EDGE, LEARN><<<(unsigned)ring_grid(P.n), kRingWarps * 32, smem, st>>>(P);
  return MLG_OK;
}
template <int NV, bool AFF, bool SRC, bool EDGE>
int launch_ring_bwd_l(const GenP& P, cudaStream_t st) {
  return P.learn ? launch_ring_bwd_k<NV, AFF, SRC, EDGE, true>(P, st) : launch_ring_bwd_k<NV, AFF, SRC, EDGE, false>(P, st);
}
template <int NV, bool AFF>
int launch_ring_bwd(const GenP& P, cudaStream_t st) {
  if (!P.src_sum) return launch_ring_bwd_l<NV, AFF, false, true>(P, st);
  return P.g_edge ? launch_ring_bwd_l<NV, AFF, true, true>(P, st) : launch_ring_bwd_l<NV, AFF, true, false>(P, st);
}

struct Cfg {
  int lanes, vec;
  bool full;
};
inline Cfg pick_cfg(long long H) {
  if (H % 4 != 0) return {32, 1, false};
  if (H <= 32) return {8, 4, H == 32};
  if (H <= 64) return {16, 4, H == 64};
  return {32, 4, H % 128 == 0};
}
inline long long grid_for(long long n, const Cfg& c) {
  const int rows_per_block = (kThreads / 32) * (32 / c.lanes);
  return (n + rows_per_block - 1) / rows_per_block;
}

template <bool FWD, int LANES, int VEC, int KIND, int SRC, bool FULL>
void launch(const GenP& P, unsigned grid, cudaStream_t st) {
  if (FWD) gen_fwd_kernel<LANES, VEC, KIND, SRC, FULL><<<grid, kThreads, 0, st>>>(P);
  else {
    size_t smem = 0;
    if (SRC == S_XA && P.pq_part) {
      smem = (size_t)(kThreads / LANES) * 2 * P.H * sizeof(float);
      if (smem > 48 * 1024)
        cudaFuncSetAttribute(gen_bwd_kernel<LANES, VEC, KIND, SRC, FULL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    }
    gen_bwd_kernel<LANES, VEC, KIND, SRC, FULL><<<grid, kThreads, smem, st>>>(P);
  }
}

template <bool FWD, int KIND, int SRC>
void launch_cfg(const GenP& P, const Cfg& c, cudaStream_t st) {
  const unsigned grid = (unsigned)grid_for(P.n, c);
  if (c.vec == 1) launch<FWD, 32, 1, KIND, SRC, false>(P, grid, st);
  else if (c.lanes == 8) { if (c.full) launch<FWD, 8, 4, KIND, SRC, true>(P, grid, st); else launch<FWD, 8, 4, KIND, SRC, false>(P, grid, st); }
  else if (c.lanes == 16) { if (c.full) launch<FWD, 16, 4, KIND, SRC, true>(P, grid, st); else launch<FWD, 16, 4, KIND, SRC, false>(P, grid, st); }
  else { if (c.full) launch<FWD, 32, 4, KIND, SRC, true>(P, grid, st); else launch<FWD, 32, 4, KIND, SRC, false>(P, grid, st); }
}

template <bool FWD>
int dispatch(const GenP& P, cudaStream_t st) {
  const Cfg c = pick_cfg(P.H);
  const int src = P.x == nullptr ? S_RAW : (P.e != nullptr ? S_XE : (P.ea != nullptr ? S_XA : S_X));
  int kind;
  switch (P.mode) {
    case MLG_AGGR_SOFTMAX: kind = K_SOFTMAX; break;
    case MLG_AGGR_POWER: kind = K_POWER; break;
    case MLG_AGGR_ADD: case MLG_AGGR_MEAN: case MLG_AGGR_MAX: kind = K_SIMPLE; break;
    default:
      mlg_set_error("mlg_gen_aggr: unknown mode %d", P.mode);
      return MLG_ERR_ARG;
  }
#define MLG_SRC(K)                                                    \
  if (src == S_XE) launch_cfg<FWD, K, S_XE>(P, c, st);                \
  else if (src == S_X) launch_cfg<FWD, K, S_X>(P, c, st);             \
  else if (src == S_XA) launch_cfg<FWD, K, S_XA>(P, c, st);           \
  else launch_cfg<FWD, K, S_RAW>(P, c, st);
  if (kind == K_SOFTMAX) { MLG_SRC(K_SOFTMAX) }
  else if (kind == K_POWER) { MLG_SRC(K_POWER) }
  else { MLG_SRC(K_SIMPLE) }
#undef MLG_SRC
  return MLG_OK;
}

int check_common(const char* who, const float* x, const float* e, const int32_t* rowptr, const int32_t* col,
                 int64_t n, int64_t H, int epilogue, const float* scale) {
  MLG_CHECK_ARG(rowptr && col, "%s: null rowptr/col", who);
  MLG_CHECK_ARG(x || (e && epilogue == MLG_EPI_NONE),
                "%s: x == NULL (raw messages in e) needs e and MLG_EPI_NONE", who);
  MLG_CHECK_ARG(n >= 0 && n < (1ll << 31) && H > 0 && H < (1ll << 20), "%s: bad sizes n=%lld H=%lld", who,
                (long long)n, (long long)H);
  MLG_CHECK_ARG(epilogue >= 0 && epilogue <= 2, "%s: bad epilogue %d", who, epilogue);
  MLG_CHECK_ARG(epilogue != MLG_EPI_MSGNORM || scale, "%s: MsgNorm epilogue needs msg_scale_dev", who);
  return MLG_OK;
}

}  // namespace

extern "C" int mlg_gen_aggr_fwd(const float* x, const float* e, const int32_t* rowptr, const int32_t* col,
                                const int32_t* eid, int64_t n, int64_t H, int mode, float t,
                                const float* t_dev, float p, const float* p_dev, const float* y_dev,
                                float eps, int epilogue, const float* msg_scale_dev, float* m, float* aux,
                                float* h, void* stream) {
  int rc = check_common("mlg_gen_aggr_fwd", x, e, rowptr, col, n, H, epilogue, msg_scale_dev);
  if (rc) return rc;
  MLG_CHECK_ARG(epilogue == MLG_EPI_NONE || h, "mlg_gen_aggr_fwd: epilogue needs h");
  {
    // m may be omitted (inference: only h is wanted) unless the MsgNorm epilogue has to re-read it (rows wider
    // than one 4*LANES chunk on the register-staged path)
    const bool single_chunk = (H % 4 == 0) && H <= 128;
    const bool ring = mode == MLG_AGGR_SOFTMAX && x && e && (H == 128 || H == 256);
    MLG_CHECK_ARG(m || (epilogue != MLG_EPI_NONE && (ring || single_chunk || epilogue == MLG_EPI_RESIDUAL)),
                  "mlg_gen_aggr_fwd: m may only be NULL when h is produced without re-reading m");
  }
  if (n == 0) return MLG_OK;
  GenP P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.e = e; P.rowptr = rowptr; P.col = col; P.eid = eid;
  P.n = (int)n; P.H = (unsigned)H; P.mode = mode; P.epi = epilogue;
  P.t = t; P.p = p; P.eps = eps; P.t_dev = t_dev; P.p_dev = p_dev; P.y_dev = y_dev;
  P.scale_dev = msg_scale_dev; P.m = m; P.aux = aux; P.h = h;
#ifndef MLG_GEN_NO_RING
  if (mode == MLG_AGGR_SOFTMAX && x && e && (H == 128 || H == 256) && ((uintptr_t)x % 16 == 0) && ((uintptr_t)e % 16 == 0)) {
    const int nv = (int)(H / 128);
    const int smem = kRingWarps * kDepth * 2 * nv * 512;
    const unsigned grid = (unsigned)mlg_ceil_div(n, kRingWarps * kRPW);
    static bool attr = false;
    if (!attr) {
      MLG_CUDA(cudaFuncSetAttribute(gen_fwd_ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kRingWarps * kDepth * 2 * 1 * 512));
      MLG_CUDA(cudaFuncSetAttribute(gen_fwd_ring_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    kRingWarps * kDepth * 2 * 2 * 512));
      attr = true;
    }
    if (nv == 1) gen_fwd_ring_kernel<1><<<grid, kRingWarps * 32, smem, (cudaStream_t)stream>>>(P);
    else gen_fwd_ring_kernel<2><<<grid, kRingWarps * 32, smem, (cudaStream_t)stream>>>(P);
    MLG_CHECK_LAUNCH("mlg_gen_aggr_fwd(ring)");
    return MLG_OK;
  }
#endif
  rc = dispatch<true>(P, (cudaStream_t)stream);
  if (rc) return rc;
  MLG_CHECK_LAUNCH("mlg_gen_aggr_fwd");
  return MLG_OK;
}

extern "C" int64_t mlg_gen_aggr_bwd_partial_rows(int64_t n, int64_t H) {
  if (n <= 0) return 1;
  return grid_for(n, pick_cfg(H));
}

// src_sum: 1 = g_x also receives the source-side sums (ring kernel only; g_edge may then be nullptr)
static int gen_aggr_bwd_impl(const float* g, const float* x, const float* e, const int32_t* rowptr,
                             const int32_t* col, const int32_t* eid, int64_t n, int64_t H, int mode,
                             int learn, float t, const float* t_dev, float p, const float* p_dev,
                             const float* y_dev, float eps, int epilogue, const float* msg_scale_dev,
                             const float* m, const float* aux, float* g_edge, float* g_x,
                             float* partials, int src_sum, void* stream) {
  int rc = check_common("mlg_gen_aggr_bwd", x, e, rowptr, col, n, H, epilogue, msg_scale_dev);
  if (rc) return rc;
  MLG_CHECK_ARG(g && m && (g_edge || src_sum) && g_x && partials, "mlg_gen_aggr_bwd: null g/m/g_edge/g_x/partials");
  MLG_CHECK_ARG(aux || (mode != MLG_AGGR_SOFTMAX && mode != MLG_AGGR_POWER),
                "mlg_gen_aggr_bwd: softmax/power backward needs aux from the forward");
  if (n == 0) return MLG_OK;
  GenP P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.e = e; P.rowptr = rowptr; P.col = col; P.eid = eid;
  P.n = (int)n; P.H = (unsigned)H; P.mode = mode; P.epi = epilogue; P.learn = learn;
  P.t = t; P.p = p; P.eps = eps; P.t_dev = t_dev; P.p_dev = p_dev; P.y_dev = y_dev;
  P.scale_dev = msg_scale_dev; P.g = g; P.m_in = m; P.aux_in = aux;
  P.g_edge = g_edge; P.g_x = g_x; P.partials = partials; P.src_sum = src_sum;
  MLG_CHECK_ARG(!src_sum || ring_bwd_ok(P), "mlg_gen_aggr_bwd_src: shape / mode not supported (see mlg_gen_aggr_bwd_src_supported)");
#ifndef MLG_GEN_NO_RING
  if (ring_bwd_ok(P)) {
    if (src_sum) MLG_CUDA(cudaMemsetAsync(g_x, 0, (size_t)n * H * 4, (cudaStream_t)stream));
    // fewer blocks than the register-staged kernel: the unused partial rows must read as zero
    const long long used = ring_grid(n), rows = mlg_gen_aggr_bwd_partial_rows(n, H);
    if (rows > used) MLG_CUDA(cudaMemsetAsync(partials + used * 4, 0, (size_t)(rows - used) * 16, (cudaStream_t)stream));
    rc = H == 128 ? launch_ring_bwd<1, false>(P, (cudaStream_t)stream) : launch_ring_bwd<2, false>(P, (cudaStream_t)stream);
    if (rc) return rc;
    MLG_CHECK_LAUNCH("mlg_gen_aggr_bwd(ring)");
    return MLG_OK;
  }
#endif
  rc = dispatch<false>(P, (cudaStream_t)stream);
  if (rc) return rc;
  MLG_CHECK_LAUNCH("mlg_gen_aggr_bwd");
  return MLG_OK;
}

extern "C" int mlg_gen_aggr_bwd(const float* g, const float* x, const float* e, const int32_t* rowptr,
                                const int32_t* col, const int32_t* eid, int64_t n, int64_t H, int mode,
                                int learn, float t, const float* t_dev, float p, const float* p_dev,
                                const float* y_dev, float eps, int epilogue, const float* msg_scale_dev,
                                const float* m, const float* aux, float* g_edge, float* g_x,
                                float* partials, void* stream) {
  return gen_aggr_bwd_impl(g, x, e, rowptr, col, eid, n, H, mode, learn, t, t_dev, p, p_dev, y_dev, eps, epilogue,
                           msg_scale_dev, m, aux, g_edge, g_x, partials, 0, stream);
}

extern "C" int mlg_gen_aggr_bwd_src_supported(int64_t H, int mode, int have_x) {
#ifdef MLG_GEN_NO_RING
  return 0;
#else
  return mode == MLG_AGGR_SOFTMAX && (H == 128 || H == 256) && have_x;
#endif
}

extern "C" int mlg_gen_aggr_bwd_src(const float* g, const float* x, const float* e, const int32_t* rowptr,
                                    const int32_t* col, const int32_t* eid, int64_t n, int64_t H, int mode,
                                    int learn, float t, const float* t_dev, float p, const float* p_dev,
                                    const float* y_dev, float eps, int epilogue, const float* msg_scale_dev,
                                    const float* m, const float* aux, float* g_edge, float* g_x,
                                    float* partials, void* stream) {
  return gen_aggr_bwd_impl(g, x, e, rowptr, col, eid, n, H, mode, learn, t, t_dev, p, p_dev, y_dev, eps, epilogue,
                           msg_scale_dev, m, aux, g_edge, g_x, partials, 1, stream);
}

// ------------------------------------------------------------------------------------------------
// Rank-1 affine edge term: e_ij = a_e * p + q with a per-edge SCALAR a_e and two H-vectors.  This is what DeeperGCN's
// edge path produces when the raw edge attribute is one number per edge (the reference's weighted gene graph,
// dataloader/multiloader.py): edge_encoder = Linear(1 -> H) followed by every GENConv's own Linear(H -> H)
// (models/deepergcn.py:209, gcn_lib/sparse/torch_vertex.py:76-77) is affine in a_e, so the [E, H] edge embedding -- and
// its per-layer GEMM, its gradient GEMMs and the [E, H] gradient accumulation across layers -- never have to exist.
// Forward reads x rows only; backward writes g_edge (= d loss / d e_ij, the source-side pass needs it) and the caller
// reduces it to g_p = sum_e a_e * g_edge[e], g_q = sum_e g_edge[e].
// ------------------------------------------------------------------------------------------------
extern "C" int mlg_gen_aggr_fwd_affine(const float* x, const float* edge_scalar, const float* edge_p, const float* edge_q,
                                       const int32_t* rowptr, const int32_t* col, const int32_t* eid, int64_t n, int64_t H,
                                       int mode, float t, const float* t_dev, float p, const float* p_dev,
                                       const float* y_dev, float eps, int epilogue, const float* msg_scale_dev, float* m,
                                       float* aux, float* h, void* stream) {
  MLG_CHECK_ARG(x && edge_scalar && edge_p && edge_q, "mlg_gen_aggr_fwd_affine: null x / edge_scalar / edge_p / edge_q");
  int rc = check_common("mlg_gen_aggr_fwd_affine", x, nullptr, rowptr, col, n, H, epilogue, msg_scale_dev);
  if (rc) return rc;
  MLG_CHECK_ARG(epilogue == MLG_EPI_NONE || h, "mlg_gen_aggr_fwd_affine: epilogue needs h");
  MLG_CHECK_ARG(m, "mlg_gen_aggr_fwd_affine: m is required");
  if (n == 0) return MLG_OK;
  GenP P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.ea = edge_scalar; P.ep = edge_p; P.eq = edge_q; P.rowptr = rowptr; P.col = col; P.eid = eid;
  P.n = (int)n; P.H = (unsigned)H; P.mode = mode; P.epi = epilogue;
  P.t = t; P.p = p; P.eps = eps; P.t_dev = t_dev; P.p_dev = p_dev; P.y_dev = y_dev;
  P.scale_dev = msg_scale_dev; P.m = m; P.aux = aux; P.h = h;
  rc = dispatch<true>(P, (cudaStream_t)stream);
  if (rc) return rc;
  MLG_CHECK_LAUNCH("mlg_gen_aggr_fwd_affine");
  return MLG_OK;
}

// wcolsum.cu
int mlg_detail_colsum_partials(const float* partial, long long n_part, long long C, float* u, float* v, float* scratch,
                               cudaStream_t st);
long long mlg_detail_colsum_partials_scratch_floats(long long C);

namespace {
constexpr long long kPqSmemMax = 200 * 1024;   // [lane groups][2][H] staging of the fused edge-term gradient
inline bool pq_fused(long long H) { return (long long)(kThreads / pick_cfg(H).lanes) * 2 * H * 4 <= kPqSmemMax; }
}  // namespace

extern "C" int64_t mlg_gen_aggr_bwd_affine_workspace_bytes(int64_t n, int64_t n_edges, int64_t H) {
  if (n <= 0 || H <= 0) return 16;
  if (pq_fused(H)) return 4 * (grid_for(n, pick_cfg(H)) * 2 * H + mlg_detail_colsum_partials_scratch_floats(H));
  return mlg_wcolsum_workspace_bytes(n_edges, H);
}

static int gen_aggr_bwd_affine_impl(const float* g, const float* x, const float* edge_scalar, const float* edge_p,
                                    const float* edge_q, const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                                    int64_t n, int64_t n_edges, int64_t H, int mode, int learn, float t,
                                    const float* t_dev, float p, const float* p_dev, const float* y_dev, float eps,
                                    int epilogue, const float* msg_scale_dev, const float* m, const float* aux,
                                    float* g_edge, float* g_x, float* partials, float* g_p, float* g_q, void* workspace,
                                    int64_t workspace_bytes, int src_sum, void* stream) {
  const bool want_pq = g_p != nullptr || g_q != nullptr;
  MLG_CHECK_ARG(!want_pq || (workspace && workspace_bytes >= mlg_gen_aggr_bwd_affine_workspace_bytes(n, n_edges, H)),
                "mlg_gen_aggr_bwd_affine: g_p / g_q need a workspace of mlg_gen_aggr_bwd_affine_workspace_bytes()");
  if (want_pq && n == 0) {
    if (g_p) cudaMemsetAsync(g_p, 0, (size_t)H * 4, (cudaStream_t)stream);
    if (g_q) cudaMemsetAsync(g_q, 0, (size_t)H * 4, (cudaStream_t)stream);
  }
  MLG_CHECK_ARG(x && edge_scalar && edge_p && edge_q, "mlg_gen_aggr_bwd_affine: null x / edge_scalar / edge_p / edge_q");
  int rc = check_common("mlg_gen_aggr_bwd_affine", x, nullptr, rowptr, col, n, H, epilogue, msg_scale_dev);
  if (rc) return rc;
  MLG_CHECK_ARG(g && m && (g_edge || src_sum) && g_x && partials, "mlg_gen_aggr_bwd_affine: null g/m/g_edge/g_x/partials");
  MLG_CHECK_ARG(aux || (mode != MLG_AGGR_SOFTMAX && mode != MLG_AGGR_POWER),
                "mlg_gen_aggr_bwd_affine: softmax/power backward needs aux from the forward");
  if (n == 0) return MLG_OK;
  GenP P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.ea = edge_scalar; P.ep = edge_p; P.eq = edge_q; P.rowptr = rowptr; P.col = col; P.eid = eid;
  P.n = (int)n; P.H = (unsigned)H; P.mode = mode; P.epi = epilogue; P.learn = learn;
  P.t = t; P.p = p; P.eps = eps; P.t_dev = t_dev; P.p_dev = p_dev; P.y_dev = y_dev;
  P.scale_dev = msg_scale_dev; P.g = g; P.m_in = m; P.aux_in = aux;
  P.g_edge = g_edge; P.g_x = g_x; P.partials = partials; P.src_sum = src_sum;
  const bool fused = want_pq && pq_fused(H);
  if (fused) P.pq_part = (float*)workspace;
  long long parts = grid_for(n, pick_cfg(H));
  MLG_CHECK_ARG(!src_sum || (ring_bwd_ok(P) && (fused || !want_pq)),
                "mlg_gen_aggr_bwd_affine_src: shape / mode not supported (see mlg_gen_aggr_bwd_src_supported)");
#ifndef MLG_GEN_NO_RING
  if (ring_bwd_ok(P) && (fused || !want_pq)) {
    if (src_sum) MLG_CUDA(cudaMemsetAsync(g_x, 0, (size_t)n * H * 4, (cudaStream_t)stream));
    const long long used = ring_grid(n);
    if (parts > used) MLG_CUDA(cudaMemsetAsync(partials + used * 4, 0, (size_t)(parts - used) * 16, (cudaStream_t)stream));
    rc = H == 128 ? launch_ring_bwd<1, true>(P, (cudaStream_t)stream) : launch_ring_bwd<2, true>(P, (cudaStream_t)stream);
    if (rc) return rc;
    MLG_CHECK_LAUNCH("mlg_gen_aggr_bwd_affine(ring)");
    if (fused)   // the partial layout [blocks][2][H] is the same, there are just fewer blocks; the scratch stays where the
                 // workspace size function put it
      return mlg_detail_colsum_partials((const float*)workspace, used, H, g_p, g_q, (float*)workspace + parts * 2 * H,
                                        (cudaStream_t)stream);
    return MLG_OK;
  }
#endif
  rc = dispatch<false>(P, (cudaStream_t)stream);
  if (rc) return rc;
  MLG_CHECK_LAUNCH("mlg_gen_aggr_bwd_affine");
  if (fused) {
    return mlg_detail_colsum_partials((const float*)workspace, parts, H, g_p, g_q, (float*)workspace + parts * 2 * H,
                                      (cudaStream_t)stream);
  }
  if (want_pq)   // very wide rows: one extra streaming pass over g_edge
    return mlg_wcolsum(g_edge, H, edge_scalar, n_edges, H, g_p, g_q, workspace, workspace_bytes, stream);
  return MLG_OK;
}

extern "C" int mlg_gen_aggr_bwd_affine(const float* g, const float* x, const float* edge_scalar, const float* edge_p,
                                       const float* edge_q, const int32_t* rowptr, const int32_t* col, const int32_t* eid,
                                       int64_t n, int64_t n_edges, int64_t H, int mode, int learn, float t,
                                       const float* t_dev, float p, const float* p_dev, const float* y_dev, float eps,
                                       int epilogue, const float* msg_scale_dev, const float* m, const float* aux,
                                       float* g_edge, float* g_x, float* partials, float* g_p, float* g_q, void* workspace,
                                       int64_t workspace_bytes, void* stream) {
  return gen_aggr_bwd_affine_impl(g, x, edge_scalar, edge_p, edge_q, rowptr, col, eid, n, n_edges, H, mode, learn, t, t_dev, p,
                                  p_dev, y_dev, eps, epilogue, msg_scale_dev, m, aux, g_edge, g_x, partials, g_p, g_q,
                                  workspace, workspace_bytes, 0, stream);
}

extern "C" int mlg_gen_aggr_bwd_affine_src(const float* g, const float* x, const float* edge_scalar, const float* edge_p,
                                           const float* edge_q, const int32_t* rowptr, const int32_t* col,
                                           const int32_t* eid, int64_t n, int64_t n_edges, int64_t H, int mode, int learn,
                                           float t, const float* t_dev, float p, const float* p_dev, const float* y_dev,
                                           float eps, int epilogue, const float* msg_scale_dev, const float* m,
                                           const float* aux, float* g_edge, float* g_x, float* partials, float* g_p,
                                           float* g_q, void* workspace, int64_t workspace_bytes, void* stream) {
  return gen_aggr_bwd_affine_impl(g, x, edge_scalar, edge_p, edge_q, rowptr, col, eid, n, n_edges, H, mode, learn, t, t_dev, p,
                                  p_dev, y_dev, eps, epilogue, msg_scale_dev, m, aux, g_edge, g_x, partials, g_p, g_q,
                                  workspace, workspace_bytes, 1, stream);
}
