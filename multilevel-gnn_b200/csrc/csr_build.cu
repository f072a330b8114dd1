// Target-sorted (or source-sorted) CSR construction on the GPU (sm_100a).
//
// Replaces the per-forward remove_self_loops / add_self_loops rewrite
// (models/gcn_lib/sparse/torch_vertex.py:272-273) and the index expansion torch_scatter performs on
// every call.  Stable: entries of a row keep edge order, added self loops come last, so every
// downstream segment sum has a fixed order.  Plumbing, not a hot kernel: the sort is CUB's radix sort
// (header-only, part of the CUDA toolkit); everything else is three small kernels.
#include <cub/cub.cuh>

#include "common.cuh"
#include "../../include/mlg_b200.h"

namespace {

__global__ void csr_keys_kernel(const long long* __restrict__ ei, long long E, int n_rows, int by_source,
                                int drop_self, long long cap, unsigned* __restrict__ keys,
                                unsigned* __restrict__ vals) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= cap) return;
  unsigned key;
  if (i < E) {
    const long long s = ei[i], d = ei[E + i];
    const long long r = by_source ? s : d;
    const long long o = by_source ? d : s;
    const bool ok = r >= 0 && r < n_rows && o >= 0 && !(drop_self && s == d);
    key = ok ? (unsigned)r : (unsigned)n_rows;
  } else {
    key = (unsigned)(i - E);
  }
  keys[i] = key;
  vals[i] = (unsigned)i;
}

__global__ void csr_finalize_kernel(const long long* __restrict__ ei, long long E, int n_rows, int by_source,
                                    long long cap, const unsigned* __restrict__ keys,
                                    const unsigned* __restrict__ vals, int* __restrict__ rowptr,
                                    int* __restrict__ col, int* __restrict__ eid) {
  const long long pos = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (pos >= cap) return;
  const unsigned key = keys[pos];
  const long long prev = pos > 0 ? (long long)keys[pos - 1] : -1ll;
  for (long long r = prev + 1; r <= (long long)key; ++r) rowptr[r] = (int)pos;  // r <= n_rows
  if (pos == cap - 1) {
    for (long long r = (long long)key + 1; r <= n_rows; ++r) rowptr[r] = (int)cap;
  }
  if (key == (unsigned)n_rows) {
    col[pos] = 0;
    eid[pos] = -2;
  } else {
    const unsigned v = vals[pos];
    if ((long long)v < E) {
      col[pos] = (int)(by_source ? ei[E + v] : ei[v]);
      eid[pos] = (int)v;
    } else {
      col[pos] = (int)(v - E);
      eid[pos] = -1;
    }
  }
}

__global__ void csr_empty_kernel(int n_rows, int* rowptr) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i <= n_rows) rowptr[i] = 0;
}

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t cub_temp_bytes(long long cap, int bits) {
  size_t tb = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tb, (const unsigned*)nullptr, (unsigned*)nullptr,
                                  (const unsigned*)nullptr, (unsigned*)nullptr, cap, 0, bits);
  return tb;
}

inline int key_bits(long long n_rows) {
  int b = 1;
  while ((1ll << b) <= n_rows) ++b;  // keys in [0, n_rows]
  return b;
}

}  // namespace

extern "C" int64_t mlg_csr_build_workspace_bytes(int64_t n_edges, int64_t n_rows, int add_self) {
  const long long cap = n_edges + (add_self ? n_rows : 0);
  if (cap <= 0) return 256;
  return (int64_t)(4 * align256((size_t)cap * 4) + align256(cub_temp_bytes(cap, key_bits(n_rows))) + 256);
}

extern "C" int mlg_csr_build(const int64_t* edge_index, int64_t n_edges, int64_t n_rows, int by_source,
                             int drop_self, int add_self, int32_t* rowptr, int32_t* col, int32_t* eid,
                             void* workspace, int64_t workspace_bytes, void* stream) {
  MLG_CHECK_ARG(rowptr, "mlg_csr_build: null rowptr");
  MLG_CHECK_ARG(n_edges >= 0 && n_rows >= 0, "mlg_csr_build: negative size");
  const long long cap = n_edges + (add_self ? n_rows : 0);
  MLG_CHECK_ARG(cap < (1ll << 31) - 1 && n_rows < (1ll << 31) - 1, "mlg_csr_build: sizes exceed int32");
  cudaStream_t st = (cudaStream_t)stream;
  if (cap == 0) {
    csr_empty_kernel<<<mlg_ceil_div(n_rows + 1, 256), 256, 0, st>>>((int)n_rows, rowptr);
    MLG_CHECK_LAUNCH("mlg_csr_build(empty)");
    return MLG_OK;
  }
  MLG_CHECK_ARG(edge_index || n_edges == 0, "mlg_csr_build: null edge_index");
  MLG_CHECK_ARG(col && eid && workspace, "mlg_csr_build: null col/eid/workspace");
  const int bits = key_bits(n_rows);
  const size_t arr = align256((size_t)cap * 4);
  size_t temp = cub_temp_bytes(cap, bits);
  if ((int64_t)(4 * arr + align256(temp)) > workspace_bytes) {
    mlg_set_error("mlg_csr_build: workspace too small (%lld < %lld)", (long long)workspace_bytes,
                  (long long)(4 * arr + align256(temp)));
    return MLG_ERR_WORKSPACE;
  }
  char* ws = (char*)workspace;
  MLG_CHECK_ARG(((uintptr_t)ws & 255) == 0, "mlg_csr_build: workspace must be 256-byte aligned");
  unsigned* k_in = (unsigned*)(ws);
  unsigned* k_out = (unsigned*)(ws + arr);
  unsigned* v_in = (unsigned*)(ws + 2 * arr);
  unsigned* v_out = (unsigned*)(ws + 3 * arr);
  void* d_temp = ws + 4 * arr;
  const int grid = mlg_ceil_div(cap, 256);
  csr_keys_kernel<<<grid, 256, 0, st>>>((const long long*)edge_index, n_edges, (int)n_rows, by_source,
                                        drop_self, cap, k_in, v_in);
  MLG_CHECK_LAUNCH("mlg_csr_build(keys)");
  MLG_CUDA(cub::DeviceRadixSort::SortPairs(d_temp, temp, k_in, k_out, v_in, v_out, cap, 0, bits, st));
  csr_finalize_kernel<<<grid, 256, 0, st>>>((const long long*)edge_index, n_edges, (int)n_rows, by_source, cap,
                                            k_out, v_out, rowptr, col, eid);
  MLG_CHECK_LAUNCH("mlg_csr_build(finalize)");
  return MLG_OK;
}
