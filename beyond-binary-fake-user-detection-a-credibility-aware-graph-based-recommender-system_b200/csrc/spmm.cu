// CSR SpMM with a fused layer-mean / gradient-seed epilogue, and the K-layer propagation
// schedules (forward and adjoint) built on it.
//
// Replaces (reference, /root/reference): torch.sparse.mm at lightgcn_cu.py:431,434 and
// Version-2/lighgcn_cu_pop.py:483-484, the stack().mean() at lightgcn_cu.py:446-447 /
// lighgcn_cu_pop.py:488-489, and autograd's SparseAddmmBackward for the same calls.
//
// Kernel shape: the path is an HBM/L2-bound gather.  A group of G = d/4 lanes owns one output row
// (d=64: half a warp, two rows per warp; d=128: one warp); every lane keeps one float4 of the row
// in registers, column ids/values are fetched coalesced G at a time and broadcast by shuffle, and
// the embedding-row gathers are issued UNR at a time before any FMA so that each lane keeps UNR
// independent 16-byte loads in flight.  Rows longer than CGX_LONG_ROW are split into CGX_CHUNK-
// sized chunks handled by whole CTAs (partials in workspace, summed in chunk order => bitwise
// reproducible, no atomics).
#include "common.cuh"

namespace cgx {

constexpr int SP_THREADS = 256;
constexpr int SP_UNR = 8;

__device__ __forceinline__ float4 ld_f4(const float4* p) { return __ldg(p); }
__device__ __forceinline__ void fma4(float4& a, float v, const float4& x) {
  a.x = fmaf(v, x.x, a.x);
  a.y = fmaf(v, x.y, a.y);
  a.z = fmaf(v, x.z, a.z);
  a.w = fmaf(v, x.w, a.w);
}
__device__ __forceinline__ float4 add4(const float4& a, const float4& b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 scale4(const float4& a, float s) {
  return make_float4(a.x * s, a.y * s, a.z * s, a.w * s);
}

template <int G>
__device__ __forceinline__ unsigned group_mask() {
  if (G == 32) return 0xffffffffu;
  const unsigned lane = threadIdx.x & 31;
  return ((1u << (G & 31)) - 1u) << (lane & ~(G - 1));
}

// y[v] (v < V) += sum over nnz in [begin, end) of val * X[idx, :]; one group, lane = 0..G-1.
template <int G, int V>
__device__ __forceinline__ void gather_range(const int32_t* __restrict__ idx, const float* __restrict__ val,
                                             int64_t begin, int64_t end, const float4* __restrict__ X, int lane,
                                             unsigned mask, float4 (&acc)[V]) {
  constexpr int ROW4 = G * V;  // float4 per embedding row
  for (int64_t base = begin; base < end; base += G) {
    const int64_t p = base + lane;
    int32_t c = 0;
    float w = 0.f;
    if (p < end) {
      c = __ldg(idx + p);
      w = __ldg(val + p);
    }
    const int cnt = (end - base) < G ? int(end - base) : G;
    for (int j0 = 0; j0 < cnt; j0 += SP_UNR) {
      float4 x[SP_UNR][V];
      float ww[SP_UNR];
#pragma unroll
      for (int t = 0; t < SP_UNR; ++t) {
        const int j = j0 + t;                       // j < G always holds when G >= SP_UNR; guard otherwise
        const int src = (j < G) ? j : (G - 1);
        const int32_t cj = __shfl_sync(mask, c, src, G);
        ww[t] = __shfl_sync(mask, w, src, G);
        if (j < cnt) {
#pragma unroll
          for (int v = 0; v < V; ++v) x[t][v] = ld_f4(X + int64_t(cj) * ROW4 + v * G + lane);
        } else {
          ww[t] = 0.f;
#pragma unroll
          for (int v = 0; v < V; ++v) x[t][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int t = 0; t < SP_UNR; ++t) {
#pragma unroll
        for (int v = 0; v < V; ++v) fma4(acc[v], ww[t], x[t][v]);
      }
    }
  }
}

template <int G, int V>
__device__ __forceinline__ void epilogue(int64_t row, int lane, const float4 (&y)[V], float4* __restrict__ Y,
                                         const float4* ACC_IN, float4* ACC_OUT, float acc_scale) {
  constexpr int ROW4 = G * V;
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int64_t o = row * ROW4 + v * G + lane;
    if (Y) Y[o] = y[v];
    if (ACC_OUT) {
      float4 a = ACC_IN ? ACC_IN[o] : make_float4(0.f, 0.f, 0.f, 0.f);
      ACC_OUT[o] = scale4(add4(a, y[v]), acc_scale);
    }
  }
}

// one group per row; rows above CGX_LONG_ROW are left to the chunked path
template <int G, int V>
__global__ void __launch_bounds__(SP_THREADS) k_spmm_rows(const int64_t* __restrict__ indptr,
                                                          const int32_t* __restrict__ idx,
                                                          const float* __restrict__ val, int32_t n_rows,
                                                          const float4* __restrict__ X, float4* __restrict__ Y,
                                                          const float4* ACC_IN, float4* ACC_OUT, float acc_scale) {
  const int lane = threadIdx.x & (G - 1);
  const int64_t row = (int64_t(blockIdx.x) * SP_THREADS + threadIdx.x) / G;
  if (row >= n_rows) return;
  const unsigned mask = group_mask<G>();
  const int64_t begin = __ldg(indptr + row), end = __ldg(indptr + row + 1);
  if (end - begin > CGX_LONG_ROW) return;
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  gather_range<G, V>(idx, val, begin, end, X, lane, mask, acc);
  epilogue<G, V>(row, lane, acc, Y, ACC_IN, ACC_OUT, acc_scale);
}

// one CTA per chunk of a long row -> partial[chunk][d]
template <int G, int V>
__global__ void __launch_bounds__(SP_THREADS) k_spmm_long_partial(const int64_t* __restrict__ indptr,
                                                                  const int32_t* __restrict__ idx,
                                                                  const float* __restrict__ val,
                                                                  const int32_t* __restrict__ long_rows,
                                                                  const int32_t* __restrict__ chunk_ptr,
                                                                  int32_t n_long, const float4* __restrict__ X,
                                                                  float4* __restrict__ partial) {
  constexpr int GROUPS = SP_THREADS / G;
  constexpr int ROW4 = G * V;
  __shared__ float4 red[GROUPS][ROW4];
  const int chunk = blockIdx.x;
  int lo = 0, hi = n_long;  // largest k with chunk_ptr[k] <= chunk
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (__ldg(chunk_ptr + mid) <= chunk) lo = mid; else hi = mid;
  }
  const int32_t row = __ldg(long_rows + lo);
  const int64_t rbeg = __ldg(indptr + row), rend = __ldg(indptr + row + 1);
  const int64_t cbeg = rbeg + int64_t(chunk - __ldg(chunk_ptr + lo)) * CGX_CHUNK;
  const int64_t cend = (cbeg + CGX_CHUNK < rend) ? cbeg + CGX_CHUNK : rend;
  const int lane = threadIdx.x & (G - 1), grp = threadIdx.x / G;
  constexpr int PER = CGX_CHUNK / GROUPS;
  int64_t gbeg = cbeg + int64_t(grp) * PER;
  int64_t gend = gbeg + PER < cend ? gbeg + PER : cend;
  if (gbeg > cend) gbeg = cend;
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  gather_range<G, V>(idx, val, gbeg, gend, X, lane, group_mask<G>(), acc);
#pragma unroll
  for (int v = 0; v < V; ++v) red[grp][v * G + lane] = acc[v];
  __syncthreads();
  for (int o = threadIdx.x; o < ROW4; o += SP_THREADS) {
    float4 s = red[0][o];
#pragma unroll 4
    for (int g = 1; g < GROUPS; ++g) s = add4(s, red[g][o]);
    partial[int64_t(chunk) * ROW4 + o] = s;
  }
}

// one group per long row: sum chunk partials in order, then the epilogue
template <int G, int V>
__global__ void __launch_bounds__(SP_THREADS) k_spmm_long_finish(const int32_t* __restrict__ long_rows,
                                                                 const int32_t* __restrict__ chunk_ptr,
                                                                 int32_t n_long, const float4* __restrict__ partial,
                                                                 float4* __restrict__ Y, const float4* ACC_IN,
                                                                 float4* ACC_OUT, float acc_scale) {
  constexpr int ROW4 = G * V;
  const int lane = threadIdx.x & (G - 1);
  const int64_t k = (int64_t(blockIdx.x) * SP_THREADS + threadIdx.x) / G;
  if (k >= n_long) return;
  const int c0 = __ldg(chunk_ptr + k), c1 = __ldg(chunk_ptr + k + 1);
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = c0; c < c1; ++c) {
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], partial[int64_t(c) * ROW4 + v * G + lane]);
  }
  epilogue<G, V>(int64_t(__ldg(long_rows + k)), lane, acc, Y, ACC_IN, ACC_OUT, acc_scale);
}

__global__ void k_scale(const float4* __restrict__ in, float4* __restrict__ out, int64_t n4, float s) {
  int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n4) out[p] = scale4(in[p], s);
}

template <int G, int V>
static int launch_spmm(const cgx_csr* m, const float* val, const float* X, float* Y, const float* ACC_IN,
                       float* ACC_OUT, float acc_scale, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream) {
  constexpr int GROUPS = SP_THREADS / G;
  const unsigned grid = (unsigned)ceil_div(m->n_rows, GROUPS);
  k_spmm_rows<G, V><<<grid, SP_THREADS, 0, stream>>>(
      m->indptr, m->idx, val, m->n_rows, reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(Y),
      reinterpret_cast<const float4*>(ACC_IN), reinterpret_cast<float4*>(ACC_OUT), acc_scale);
  CGX_LAUNCH_CHECK();
  if (m->n_long > 0) {
    size_t need = size_t(m->n_chunks) * G * V * sizeof(float4);
    CGX_REQUIRE(workspace != nullptr && workspace_bytes >= need, CGX_ERR_WORKSPACE,
                "spmm: workspace too small for %d long-row chunks", m->n_chunks);
    float4* partial = static_cast<float4*>(workspace);
    k_spmm_long_partial<G, V><<<(unsigned)m->n_chunks, SP_THREADS, 0, stream>>>(
        m->indptr, m->idx, val, m->long_rows, m->chunk_ptr, m->n_long, reinterpret_cast<const float4*>(X),
        partial);
    CGX_LAUNCH_CHECK();
    k_spmm_long_finish<G, V><<<(unsigned)ceil_div(m->n_long, GROUPS), SP_THREADS, 0, stream>>>(
        m->long_rows, m->chunk_ptr, m->n_long, partial, reinterpret_cast<float4*>(Y),
        reinterpret_cast<const float4*>(ACC_IN), reinterpret_cast<float4*>(ACC_OUT), acc_scale);
    CGX_LAUNCH_CHECK();
  }
  return CGX_OK;
}

static int spmm_dispatch(const cgx_csr* m, int use_bwd, int32_t d, const float* X, float* Y, const float* ACC_IN,
                         float* ACC_OUT, float acc_scale, void* ws, size_t ws_bytes, cudaStream_t stream) {
  CGX_REQUIRE(m && m->indptr && (m->nnz == 0 || (m->idx && m->val_fwd && m->val_bwd)) && X, CGX_ERR_ARG,
              "spmm: NULL pointer");
  CGX_REQUIRE(m->n_long == 0 || (m->long_rows && m->chunk_ptr), CGX_ERR_ARG, "spmm: long-row lists missing");
  CGX_REQUIRE(Y || ACC_OUT, CGX_ERR_ARG, "spmm: no output requested");
  const float* val = use_bwd ? m->val_bwd : m->val_fwd;
  if (m->n_rows == 0) return CGX_OK;
  switch (d) {
    case 16: return launch_spmm<4, 1>(m, val, X, Y, ACC_IN, ACC_OUT, acc_scale, ws, ws_bytes, stream);
    case 32: return launch_spmm<8, 1>(m, val, X, Y, ACC_IN, ACC_OUT, acc_scale, ws, ws_bytes, stream);
    case 64: return launch_spmm<16, 1>(m, val, X, Y, ACC_IN, ACC_OUT, acc_scale, ws, ws_bytes, stream);
    case 128: return launch_spmm<32, 1>(m, val, X, Y, ACC_IN, ACC_OUT, acc_scale, ws, ws_bytes, stream);
    case 256: return launch_spmm<32, 2>(m, val, X, Y, ACC_IN, ACC_OUT, acc_scale, ws, ws_bytes, stream);
    default:
      set_error("spmm: emb_dim %d unsupported (16, 32, 64, 128, 256)", d);
      return CGX_ERR_UNSUPPORTED;
  }
}

static size_t spmm_ws(const cgx_csr* m, int32_t d) { return align_up(size_t(m ? m->n_chunks : 0) * d * 4); }

}  // namespace cgx

using namespace cgx;

extern "C" size_t cgx_spmm_workspace_bytes(const cgx_csr* m, int32_t d) { return spmm_ws(m, d); }

extern "C" int cgx_spmm(const cgx_csr* m, int use_bwd_values, int32_t d, const float* X, float* Y,
                        const float* ACC_IN, float* ACC_OUT, float acc_scale, void* workspace,
                        size_t workspace_bytes, void* stream) {
  return spmm_dispatch(m, use_bwd_values, d, X, Y, ACC_IN, ACC_OUT, acc_scale, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream));
}

extern "C" size_t cgx_propagate_workspace_bytes(const cgx_csr* by_user, const cgx_csr* by_item, int32_t d) {
  if (!by_user || !by_item) return 0;
  size_t rows = size_t(by_user->n_rows) + size_t(by_item->n_rows);
  size_t long_ws = spmm_ws(by_user, d) > spmm_ws(by_item, d) ? spmm_ws(by_user, d) : spmm_ws(by_item, d);
  return 2 * (align_up(size_t(by_user->n_rows) * d * 4) + align_up(size_t(by_item->n_rows) * d * 4)) + long_ws +
         256 + 0 * rows;
}

extern "C" int cgx_propagate_fwd(const cgx_csr* by_user, const cgx_csr* by_item, int order, int32_t K, int32_t d,
                                 const float* e0_u, const float* e0_i, float* out_u, float* out_i,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(by_user && by_item && e0_u && e0_i && out_u && out_i, CGX_ERR_ARG, "propagate_fwd: NULL pointer");
  CGX_REQUIRE(K >= 1, CGX_ERR_ARG, "propagate_fwd: num_layers must be >= 1");
  CGX_REQUIRE(order == CGX_ORDER_JACOBI || order == CGX_ORDER_GS, CGX_ERR_ARG, "propagate_fwd: bad order");
  CGX_REQUIRE(by_user->n_rows == by_item->n_cols && by_user->n_cols == by_item->n_rows, CGX_ERR_ARG,
              "propagate_fwd: operator shapes disagree");
  CGX_REQUIRE(workspace_bytes >= cgx_propagate_workspace_bytes(by_user, by_item, d), CGX_ERR_WORKSPACE,
              "propagate_fwd: workspace too small");
  const int64_t U = by_user->n_rows, I = by_item->n_rows;
  Arena ws(workspace, workspace_bytes);
  float* ub[2] = {ws.take<float>(U * d), ws.take<float>(U * d)};
  float* ib[2] = {ws.take<float>(I * d), ws.take<float>(I * d)};
  size_t lws = spmm_ws(by_user, d) > spmm_ws(by_item, d) ? spmm_ws(by_user, d) : spmm_ws(by_item, d);
  void* lw = ws.take<char>(lws);
  CGX_REQUIRE(ws.ok, CGX_ERR_WORKSPACE, "propagate_fwd: workspace too small");
  const float s = 1.0f / float(K + 1);
  const float* cu = e0_u;  // layer k tables
  const float* ci = e0_i;
  for (int k = 0; k < K; ++k) {
    const bool first = k == 0, last = k == K - 1;
    float* nu = ub[k & 1];
    float* ni = ib[k & 1];
    const float sc = last ? s : 1.0f;
    // item side: i_{k+1} = C u_k
    const bool need_ni = !last || order == CGX_ORDER_GS;
    CGX_TRY(spmm_dispatch(by_item, 0, d, cu, need_ni ? ni : nullptr, first ? e0_i : out_i, out_i, sc, lw, lws,
                          stream));
    // user side: u_{k+1} = A i_k (Jacobi) or A i_{k+1} (Gauss-Seidel)
    const float* src = order == CGX_ORDER_JACOBI ? ci : ni;
    CGX_TRY(spmm_dispatch(by_user, 0, d, src, last ? nullptr : nu, first ? e0_u : out_u, out_u, sc, lw, lws,
                          stream));
    cu = nu;
    ci = ni;
  }
  return CGX_OK;
}

extern "C" int cgx_propagate_bwd(const cgx_csr* by_user, const cgx_csr* by_item, int order, int32_t K, int32_t d,
                                 const float* g_u, const float* g_i, float* d_e0_u, float* d_e0_i,
                                 void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(by_user && by_item && g_u && g_i && d_e0_u && d_e0_i, CGX_ERR_ARG, "propagate_bwd: NULL pointer");
  CGX_REQUIRE(K >= 1, CGX_ERR_ARG, "propagate_bwd: num_layers must be >= 1");
  CGX_REQUIRE(order == CGX_ORDER_JACOBI || order == CGX_ORDER_GS, CGX_ERR_ARG, "propagate_bwd: bad order");
  CGX_REQUIRE(workspace_bytes >= cgx_propagate_workspace_bytes(by_user, by_item, d), CGX_ERR_WORKSPACE,
              "propagate_bwd: workspace too small");
  const int64_t U = by_user->n_rows, I = by_item->n_rows;
  Arena ws(workspace, workspace_bytes);
  float* ub[2] = {ws.take<float>(U * d), ws.take<float>(U * d)};
  float* ib[2] = {ws.take<float>(I * d), ws.take<float>(I * d)};
  size_t lws = spmm_ws(by_user, d) > spmm_ws(by_item, d) ? spmm_ws(by_user, d) : spmm_ws(by_item, d);
  void* lw = ws.take<char>(lws);
  CGX_REQUIRE(ws.ok, CGX_ERR_WORKSPACE, "propagate_bwd: workspace too small");
  const float s = 1.0f / float(K + 1);
  // Work with the unscaled adjoints bu' = bu / s, bi' = bi / s (the recurrences are linear):
  //   Jacobi: (bu', bi') <- (g_u + C^T bi', g_i + A^T bu');  Gauss-Seidel: bi' = g_i + A^T bu'; bu' = g_u + C^T bi'
  // and fold s into the last product's epilogue.
  if (order == CGX_ORDER_JACOBI) {
    const float* bu = g_u;
    const float* bi = g_i;
    for (int k = 0; k < K; ++k) {
      const bool last = k == K - 1;
      float* nu = last ? d_e0_u : ub[k & 1];
      float* ni = last ? d_e0_i : ib[k & 1];
      const float sc = last ? s : 1.0f;
      CGX_TRY(spmm_dispatch(by_user, 1, d, bi, nullptr, g_u, nu, sc, lw, lws, stream));  // g_u + C^T bi
      CGX_TRY(spmm_dispatch(by_item, 1, d, bu, nullptr, g_i, ni, sc, lw, lws, stream));  // g_i + A^T bu
      bu = nu;
      bi = ni;
    }
  } else {
    const float* bu = g_u;
    for (int k = 0; k < K; ++k) {
      const bool last = k == K - 1;
      float* ni = ib[0];
      float* nu = last ? d_e0_u : ub[k & 1];
      CGX_TRY(spmm_dispatch(by_item, 1, d, bu, nullptr, g_i, ni, 1.0f, lw, lws, stream));            // bi'
      CGX_TRY(spmm_dispatch(by_user, 1, d, ni, nullptr, g_u, nu, last ? s : 1.0f, lw, lws, stream));  // bu'
      bu = nu;
    }
    const int64_t n4 = I * d / 4;
    k_scale<<<(unsigned)ceil_div(n4, 256), 256, 0, stream>>>(reinterpret_cast<const float4*>(g_i),
                                                             reinterpret_cast<float4*>(d_e0_i), n4, s);
    CGX_LAUNCH_CHECK();
  }
  return CGX_OK;
}
