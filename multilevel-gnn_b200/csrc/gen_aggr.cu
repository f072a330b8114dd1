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
// shuffles; UN=4 edges' source rows and edge-feature rows are requested before any is consumed.
// Softmax is evaluated online (running max / sum / weighted sum, one exp2 per element).
// HBM-bound: algorithmic bytes fwd = 4H(E + 2N) + 4E + 4(N+1)   (SURVEY.md section 8d).
#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

struct GenP {
  const float* x;
  const float* e;
  const int* rowptr;
  const int* col;
  const int* eid;
  int n;
  int H;
  int mode, learn, epi, raw;
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
  float* g_edge;
  float* g_x;
  float* partials;
};

constexpr int kThreads = 256;
constexpr int UN = 4;
constexpr float kLo = 1e-7f, kHi = 1e1f;

template <int VEC>
struct Vec {
  float v[VEC];
};

template <int VEC>
__device__ __forceinline__ Vec<VEC> load_gather(const float* p, bool ok) {
  Vec<VEC> r;
  if (VEC == 4) {
    float4 t = ok ? ld_gather4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.v[0] = t.x; r.v[1 % VEC] = t.y; r.v[2 % VEC] = t.z; r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = ok ? __ldg(p) : 0.f;
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> load_stream(const float* p, bool ok) {
  Vec<VEC> r;
  if (VEC == 4) {
    float4 t = ok ? ld_stream4(p) : make_float4(0.f, 0.f, 0.f, 0.f);
    r.v[0] = t.x; r.v[1 % VEC] = t.y; r.v[2 % VEC] = t.z; r.v[3 % VEC] = t.w;
  } else {
    r.v[0] = ok ? __ldg(p) : 0.f;
  }
  return r;
}
template <int VEC>
__device__ __forceinline__ void store_vec(float* p, const Vec<VEC>& r, bool ok, bool stream) {
  if (!ok) return;
  if (VEC == 4) {
    float4 t = make_float4(r.v[0], r.v[1 % VEC], r.v[2 % VEC], r.v[3 % VEC]);
    if (stream) st_stream4(p, t); else st4(p, t);
  } else {
    *p = r.v[0];
  }
}

__device__ __forceinline__ float sigmoidf_(float y) { return 1.f / (1.f + expf(-y)); }

// ------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------
template <int LANES, int VEC, int MODE, bool HAS_E>
__global__ void __launch_bounds__(kThreads) gen_fwd_kernel(GenP P) {
  constexpr int RPW = 32 / LANES;
  constexpr int CW = LANES * VEC;  // channels per chunk
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, sl = lane % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long row = warp * RPW + sub;
  if (row >= P.n) return;  // group-uniform exit; no block-level sync in this kernel

  const int H = P.H;
  const bool raw = P.raw != 0;  // messages given directly in e (GenMessagePassing.aggregate drop-in)
  const int beg = __ldg(P.rowptr + row), end = __ldg(P.rowptr + row + 1);
  const int deg = end - beg;
  const float t = P.t_dev ? __ldg(P.t_dev) : P.t;
  const float pw = P.p_dev ? __ldg(P.p_dev) : P.p;
  const float tl2 = t * MLG_LOG2E;
  const float eps = P.eps;
  float degpow = 1.f;
  if (P.y_dev) degpow = powf((float)deg, sigmoidf_(__ldg(P.y_dev)));
  const int nchunks = (H + CW - 1) / CW;
  const float* xrow = P.x + (size_t)row * H;
  float* mrow = P.m + (size_t)row * H;

  float sx2 = 0.f, sm2 = 0.f;
  Vec<VEC> out, xi;
#pragma unroll
  for (int k = 0; k < VEC; ++k) out.v[k] = xi.v[k] = 0.f;

  for (int ch = 0; ch < nchunks; ++ch) {
    const int c = ch * CW + sl * VEC;
    const bool cok = c < H;
    float a0[VEC], a1[VEC], a2[VEC];
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      a0[k] = (MODE == MLG_AGGR_SOFTMAX || MODE == MLG_AGGR_MAX) ? -INFINITY : 0.f;
      a1[k] = 0.f;
      a2[k] = 0.f;
    }
    for (int base = beg; base < end; base += LANES) {
      const int q = min(base + sl, end - 1);
      const int my_col = __ldg(P.col + q);
      const int my_e = HAS_E ? (P.eid ? __ldg(P.eid + q) : q) : 0;
      const int cnt = min(LANES, end - base);
      for (int j = 0; j < cnt; j += UN) {
        Vec<VEC> xv[UN], ev[UN];
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          const int jj = min(j + u, cnt - 1);
          const int s = __shfl_sync(gmask, my_col, jj, LANES);
          xv[u] = load_gather<VEC>(P.x + (size_t)s * H + c, cok && !raw);
          if (HAS_E) {
            const int ee = __shfl_sync(gmask, my_e, jj, LANES);
            ev[u] = load_stream<VEC>(P.e + (size_t)ee * H + c, cok);
          }
        }
#pragma unroll
        for (int u = 0; u < UN; ++u) {
          if (j + u < cnt) {
#pragma unroll
            for (int k = 0; k < VEC; ++k) {
              const float pre = HAS_E ? xv[u].v[k] + ev[u].v[k] : xv[u].v[k];
              const float v = raw ? pre : fmaxf(pre, 0.f) + eps;
              if (MODE == MLG_AGGR_SOFTMAX) {
                const float z = v * tl2;
                const float d = z - a0[k];
                const float ex = exp2f(-fabsf(d));
                const bool up = d > 0.f;
                a1[k] = up ? fmaf(a1[k], ex, 1.f) : a1[k] + ex;
                a2[k] = up ? fmaf(a2[k], ex, v) : fmaf(v, ex, a2[k]);
                a0[k] = up ? z : a0[k];
              } else if (MODE == MLG_AGGR_POWER) {
                const float vc = fminf(fmaxf(v, kLo), kHi);
                a0[k] += (pw == 1.f) ? vc : powf(vc, pw);
              } else if (MODE == MLG_AGGR_MAX) {
                a0[k] = fmaxf(a0[k], v);
              } else {
                a0[k] += v;
              }
            }
          }
        }
      }
    }
    // finalise this chunk
    Vec<VEC> o, ax;
#pragma unroll
    for (int k = 0; k < VEC; ++k) {
      float r = 0.f, au = 0.f;
      if (MODE == MLG_AGGR_SOFTMAX) {
        if (deg > 0) {
          r = a2[k] / a1[k];
          au = a0[k] + log2f(a1[k]);
        }
      } else if (MODE == MLG_AGGR_POWER) {
        const float mean = a0[k] / (float)max(deg, 1);
        au = mean;
        const float cl = fminf(fmaxf(mean, kLo), kHi);
        r = (pw == 1.f) ? cl : powf(cl, 1.f / pw);
      } else if (MODE == MLG_AGGR_MAX) {
        r = deg > 0 ? a0[k] : 0.f;
      } else if (MODE == MLG_AGGR_MEAN) {
        r = a0[k] / (float)max(deg, 1);
      } else {
        r = a0[k];
      }
      if (P.y_dev) r *= degpow;
      o.v[k] = r;
      ax.v[k] = au;
    }
    store_vec<VEC>(mrow + c, o, cok, false);
    if (P.aux) store_vec<VEC>(P.aux + (size_t)row * H + c, ax, cok, true);
    if (P.epi != MLG_EPI_NONE) {
      Vec<VEC> xr = load_gather<VEC>(xrow + c, cok);
      if (P.epi == MLG_EPI_RESIDUAL) {
        Vec<VEC> hv;
#pragma unroll
        for (int k = 0; k < VEC; ++k) hv.v[k] = xr.v[k] + o.v[k];
        store_vec<VEC>(P.h + (size_t)row * H + c, hv, cok, false);
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
    for (int ch = 0; ch < nchunks; ++ch) {
      const int c = ch * CW + sl * VEC;
      const bool cok = c < H;
      Vec<VEC> mo, xr;
      if (nchunks == 1) {
        mo = out;
        xr = xi;
      } else {
        // re-read this lane's own writes of m (same thread wrote them: visible without a fence)
        mo.v[0] = 0.f;
        if (cok) {
#pragma unroll
          for (int k = 0; k < VEC; ++k) mo.v[k] = mrow[c + k];
        }
        xr = load_gather<VEC>(xrow + c, cok);
      }
      Vec<VEC> hv;
#pragma unroll
      for (int k = 0; k < VEC; ++k) hv.v[k] = fmaf(f, mo.v[k], xr.v[k]);
      store_vec<VEC>(P.h + (size_t)row * H + c, hv, cok, false);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------
template <int LANES, int VEC, int MODE, bool HAS_E>
__global__ void __launch_bounds__(kThreads) gen_bwd_kernel(GenP P) {
  constexpr int RPW = 32 / LANES;
  constexpr int CW = LANES * VEC;
  __shared__ float red[3 * 32];
  const int lane = threadIdx.x & 31;
  const int sub = lane / LANES, sl = lane % LANES;
  const unsigned gmask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << (sub * LANES));
  const long long warp = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const long long row = warp * RPW + sub;
  const bool active = row < P.n;

  float acc[3] = {0.f, 0.f, 0.f};  // d/dt (or d/dp), d/dy_raw, d/dmsg_scale

  if (active) {
    const int H = P.H;
    const bool raw = P.raw != 0;
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
    const float* xrow = P.x + (size_t)row * H;
    const float* mrow = P.m_in + (size_t)row * H;
    const float* grow = P.g + (size_t)row * H;

    // --- MsgNorm statistics of this row (one sweep over channels) ---
    float f_gm = 1.f, f_u = 0.f, f_x = 0.f;  // g_m = f_gm * g - f_u * m ; g_x = g + f_x * x
    if (P.epi == MLG_EPI_MSGNORM) {
      float sx2 = 0.f, sm2 = 0.f, sgm = 0.f;
      for (int ch = 0; ch < nchunks; ++ch) {
        const int c = ch * CW + sl * VEC;
        const bool cok = c < H;
        Vec<VEC> xr = load_gather<VEC>(xrow + c, cok), mr = load_gather<VEC>(mrow + c, cok),
                 gr = load_gather<VEC>(grow + c, cok);
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
      const int c = ch * CW + sl * VEC;
      const bool cok = c < H;
      Vec<VEC> gr = load_gather<VEC>(grow + c, cok);
      Vec<VEC> mr = load_gather<VEC>(mrow + c, cok);
      Vec<VEC> au;
      au.v[0] = 0.f;
      if (MODE == MLG_AGGR_SOFTMAX || MODE == MLG_AGGR_POWER)
        au = load_gather<VEC>(P.aux_in + (size_t)row * H + c, cok);
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
          Vec<VEC> xr = load_gather<VEC>(xrow + c, cok);
#pragma unroll
          for (int k = 0; k < VEC; ++k) gx.v[k] = fmaf(f_x, xr.v[k], gr.v[k]);
        }
        store_vec<VEC>(P.g_x + (size_t)row * H + c, gx, cok, false);
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
        if (MODE == MLG_AGGR_POWER) {
          const float mean = au.v[k];
          const bool inr = mean >= kLo && mean <= kHi;
          const float cl = fminf(fmaxf(mean, kLo), kHi);
          // d out / d v_e = inr * (oi/cl) * vc^(p-1) / deg
          k1[k] = inr ? gin[k] * oi[k] / cl * inv_deg : 0.f;
          if (P.learn && cok) {
            acc[0] = fmaf(gin[k] * oi[k], -logf(cl) / (pw * pw), acc[0]);
            k2[k] = inr ? gin[k] * oi[k] / (pw * cl) * inv_deg : 0.f;
          }
        } else if (MODE == MLG_AGGR_MEAN) {
          k1[k] = gin[k] * inv_deg;
        }
      }

      for (int base = beg; base < end; base += LANES) {
        const int q = min(base + sl, end - 1);
        const int my_col = __ldg(P.col + q);
        const int my_e = P.eid ? __ldg(P.eid + q) : q;
        const int cnt = min(LANES, end - base);
        for (int j = 0; j < cnt; j += UN) {
          Vec<VEC> xv[UN], ev[UN];
          int eo[UN];
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            const int jj = min(j + u, cnt - 1);
            const int s = __shfl_sync(gmask, my_col, jj, LANES);
            eo[u] = __shfl_sync(gmask, my_e, jj, LANES);
            xv[u] = load_gather<VEC>(P.x + (size_t)s * H + c, cok && !raw);
            if (HAS_E) ev[u] = load_stream<VEC>(P.e + (size_t)eo[u] * H + c, cok);
          }
#pragma unroll
          for (int u = 0; u < UN; ++u) {
            if (j + u < cnt) {
              Vec<VEC> ge;
#pragma unroll
              for (int k = 0; k < VEC; ++k) {
                const float pre = HAS_E ? xv[u].v[k] + ev[u].v[k] : xv[u].v[k];
                const float v = raw ? pre : fmaxf(pre, 0.f) + eps;
                float gv;
                if (MODE == MLG_AGGR_SOFTMAX) {
                  const float w = exp2f(fmaf(v, tl2, -au.v[k]));
                  const float gw = gin[k] * w;
                  if (P.learn) {
                    const float dv = v - oi[k];
                    gv = gw * fmaf(t, dv, 1.f);
                    if (cok) acc[0] = fmaf(gw * v, dv, acc[0]);
                  } else {
                    gv = gw;
                  }
                } else if (MODE == MLG_AGGR_POWER) {
                  const float vc = fminf(fmaxf(v, kLo), kHi);
                  const float vp1 = (pw == 1.f) ? 1.f : powf(vc, pw - 1.f);
                  gv = (v <= kHi) ? k1[k] * vp1 : 0.f;
                  if (P.learn && cok) acc[0] = fmaf(k2[k] * vp1 * vc, logf(vc), acc[0]);
                } else if (MODE == MLG_AGGR_MAX) {
                  const bool hit = (!taken[k]) && (v == oi[k]);
                  gv = hit ? gin[k] : 0.f;
                  taken[k] = taken[k] || hit;
                } else if (MODE == MLG_AGGR_MEAN) {
                  gv = k1[k];
                } else {
                  gv = gin[k];
                }
                ge.v[k] = (raw || pre > 0.f) ? gv : 0.f;
              }
              store_vec<VEC>(P.g_edge + (size_t)eo[u] * H + c, ge, cok, true);
            }
          }
        }
      }
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

struct Cfg {
  int lanes, vec;
};
inline Cfg pick_cfg(long long H) {
  if (H % 4 != 0) return {32, 1};
  if (H <= 32) return {8, 4};
  if (H <= 64) return {16, 4};
  return {32, 4};
}
inline long long grid_for(long long n, const Cfg& c) {
  const int rows_per_block = (kThreads / 32) * (32 / c.lanes);
  return (n + rows_per_block - 1) / rows_per_block;
}

template <int MODE, bool HAS_E>
void launch_fwd(const GenP& P, const Cfg& c, cudaStream_t st) {
  const unsigned grid = (unsigned)grid_for(P.n, c);
  if (c.vec == 1) gen_fwd_kernel<32, 1, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
  else if (c.lanes == 8) gen_fwd_kernel<8, 4, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
  else if (c.lanes == 16) gen_fwd_kernel<16, 4, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
  else gen_fwd_kernel<32, 4, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
}
template <int MODE, bool HAS_E>
void launch_bwd(const GenP& P, const Cfg& c, cudaStream_t st) {
  const unsigned grid = (unsigned)grid_for(P.n, c);
  if (c.vec == 1) gen_bwd_kernel<32, 1, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
  else if (c.lanes == 8) gen_bwd_kernel<8, 4, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
  else if (c.lanes == 16) gen_bwd_kernel<16, 4, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
  else gen_bwd_kernel<32, 4, MODE, HAS_E><<<grid, kThreads, 0, st>>>(P);
}

template <bool FWD>
int dispatch(const GenP& P, cudaStream_t st) {
  const Cfg c = pick_cfg(P.H);
  const bool he = P.e != nullptr;
#define MLG_GO(M)                                                        \
  if (FWD) { if (he) launch_fwd<M, true>(P, c, st); else launch_fwd<M, false>(P, c, st); } \
  else     { if (he) launch_bwd<M, true>(P, c, st); else launch_bwd<M, false>(P, c, st); }
  switch (P.mode) {
    case MLG_AGGR_SOFTMAX: MLG_GO(MLG_AGGR_SOFTMAX); break;
    case MLG_AGGR_POWER: MLG_GO(MLG_AGGR_POWER); break;
    case MLG_AGGR_ADD: MLG_GO(MLG_AGGR_ADD); break;
    case MLG_AGGR_MEAN: MLG_GO(MLG_AGGR_MEAN); break;
    case MLG_AGGR_MAX: MLG_GO(MLG_AGGR_MAX); break;
    default:
      mlg_set_error("mlg_gen_aggr: unknown mode %d", P.mode);
      return MLG_ERR_ARG;
  }
#undef MLG_GO
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
  MLG_CHECK_ARG(m, "mlg_gen_aggr_fwd: null m");
  MLG_CHECK_ARG(epilogue == MLG_EPI_NONE || h, "mlg_gen_aggr_fwd: epilogue needs h");
  if (n == 0) return MLG_OK;
  GenP P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.e = e; P.rowptr = rowptr; P.col = col; P.eid = eid;
  P.n = (int)n; P.H = (int)H; P.mode = mode; P.epi = epilogue; P.raw = x ? 0 : 1;
  P.t = t; P.p = p; P.eps = eps; P.t_dev = t_dev; P.p_dev = p_dev; P.y_dev = y_dev;
  P.scale_dev = msg_scale_dev; P.m = m; P.aux = aux; P.h = h;
  rc = dispatch<true>(P, (cudaStream_t)stream);
  if (rc) return rc;
  MLG_CHECK_LAUNCH("mlg_gen_aggr_fwd");
  return MLG_OK;
}

extern "C" int64_t mlg_gen_aggr_bwd_partial_rows(int64_t n, int64_t H) {
  if (n <= 0) return 1;
  return grid_for(n, pick_cfg(H));
}

extern "C" int mlg_gen_aggr_bwd(const float* g, const float* x, const float* e, const int32_t* rowptr,
                                const int32_t* col, const int32_t* eid, int64_t n, int64_t H, int mode,
                                int learn, float t, const float* t_dev, float p, const float* p_dev,
                                const float* y_dev, float eps, int epilogue, const float* msg_scale_dev,
                                const float* m, const float* aux, float* g_edge, float* g_x,
                                float* partials, void* stream) {
  int rc = check_common("mlg_gen_aggr_bwd", x, e, rowptr, col, n, H, epilogue, msg_scale_dev);
  if (rc) return rc;
  MLG_CHECK_ARG(g && m && g_edge && g_x && partials, "mlg_gen_aggr_bwd: null g/m/g_edge/g_x/partials");
  MLG_CHECK_ARG(aux || (mode != MLG_AGGR_SOFTMAX && mode != MLG_AGGR_POWER),
                "mlg_gen_aggr_bwd: softmax/power backward needs aux from the forward");
  if (n == 0) return MLG_OK;
  GenP P;
  memset(&P, 0, sizeof(P));
  P.x = x; P.e = e; P.rowptr = rowptr; P.col = col; P.eid = eid;
  P.n = (int)n; P.H = (int)H; P.mode = mode; P.epi = epilogue; P.learn = learn; P.raw = x ? 0 : 1;
  P.t = t; P.p = p; P.eps = eps; P.t_dev = t_dev; P.p_dev = p_dev; P.y_dev = y_dev;
  P.scale_dev = msg_scale_dev; P.g = g; P.m_in = m; P.aux_in = aux;
  P.g_edge = g_edge; P.g_x = g_x; P.partials = partials;
  rc = dispatch<false>(P, (cudaStream_t)stream);
  if (rc) return rc;
  MLG_CHECK_LAUNCH("mlg_gen_aggr_bwd");
  return MLG_OK;
}
