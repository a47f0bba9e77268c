// Graph build on device: degree vectors, the duplicate-keeping user-row CSR, and both coalesced,
// credibility-weighted operators in both row orders.
//
// Replaces (reference, /root/reference):
//   edges_to_user_csr            lightgcn_cu.py:259-276
//   build_cred_weighted_mats     lightgcn_cu.py:368-399
//   build_message_passing_mats   Version-2/lighgcn_cu_pop.py:429-452,
//                                version_1/lightgcn_cu_pop_Degree-Aware Message.py:349-403
// Bit-exactness: the reference does this arithmetic with NumPy float32 ufuncs, which are
// correctly rounded; every float op below is an explicit round-to-nearest intrinsic
// (__fmul_rn / __fdiv_rn / __fsqrt_rn / __fadd_rn) so that nvcc can neither contract to FMA
// nor substitute an approximate division.
#include "common.cuh"

namespace cgx {

constexpr int GB_THREADS = 256;

__global__ void k_pack_and_count(const int32_t* __restrict__ eu, const int32_t* __restrict__ ei, int64_t E,
                                 int32_t U, int32_t I, int bits_u, int bits_i,
                                 uint64_t* __restrict__ key_ui, uint64_t* __restrict__ key_iu,
                                 int32_t* deg_u, int32_t* deg_i, unsigned long long* bad) {
  int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (e >= E) return;
  int32_t u = eu[e], i = ei[e];
  if (u < 0 || u >= U || i < 0 || i >= I) {
    atomicAdd(bad, 1ull);
    u = 0;
    i = 0;  // keep the arrays well formed; the caller rejects the build when bad != 0
  }
  if (key_ui) key_ui[e] = (uint64_t(uint32_t(u)) << bits_i) | uint32_t(i);
  if (key_iu) key_iu[e] = (uint64_t(uint32_t(i)) << bits_u) | uint32_t(u);
  if (deg_u) atomicAdd(&deg_u[u], 1);
  if (deg_i) atomicAdd(&deg_i[i], 1);
}

__global__ void k_widen(const uint32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n) {
  int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n) out[p] = int64_t(in[p]);
}

__global__ void k_unpack_minor(const uint64_t* __restrict__ keys, int64_t n, int minor_bits,
                               int32_t* __restrict__ minor) {
  int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n) minor[p] = int32_t(keys[p] & ((uint64_t(1) << minor_bits) - 1));
}

__global__ void k_head_flags(const uint64_t* __restrict__ keys, int64_t n, uint32_t* __restrict__ flags) {
  int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n) flags[p] = (p == 0 || keys[p] != keys[p - 1]) ? 1u : 0u;
}

struct WeightArgs {
  const int32_t* deg_u;
  const int32_t* deg_i;
  const float* cred;
  const float* alpha;
  int variant;
};

// (w_A, w_C) of one (u, i) pair; mirrors oracle.edge_weights op for op.
__device__ __forceinline__ void edge_weight(const WeightArgs& a, int32_t u, int32_t i, float* w_a, float* w_c) {
  float du = float(a.deg_u[u]), di = float(a.deg_i[i]);  // int -> float, round to nearest even (== astype)
  float c = a.cred[u];
  if (a.variant == CGX_VARIANT_CU) {  // lightgcn_cu.py:386-389
    float denom = __fsqrt_rn(fmaxf(__fmul_rn(du, di), 1e-12f));
    *w_c = __fdiv_rn(c, denom);
    *w_a = __fdiv_rn(1.0f, denom);
  } else {  // lighgcn_cu_pop.py:436-446
    float isu = __fdiv_rn(1.0f, __fsqrt_rn(fmaxf(du, 1.0f)));
    float isi = __fdiv_rn(1.0f, __fsqrt_rn(fmaxf(di, 1.0f)));
    float w = __fmul_rn(isu, isi);
    if (a.variant == CGX_VARIANT_DA) w = __fmul_rn(w, a.alpha[i]);  // Degree-Aware Message.py:383
    *w_a = w;
    *w_c = __fmul_rn(c, w);
  }
}

// One thread per sorted entry; run heads emit the coalesced entry.  Duplicates of one (u, i) pair
// carry identical weights and are added left to right, as torch's CPU coalesce() does.
__global__ void k_emit(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ pos, int64_t n,
                       int minor_bits, int major_is_user, WeightArgs wa, int32_t* __restrict__ idx,
                       float* __restrict__ val_fwd, float* __restrict__ val_bwd) {
  int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  uint64_t k = keys[p];
  if (p > 0 && keys[p - 1] == k) return;
  int mult = 1;
  while (p + mult < n && keys[p + mult] == k) ++mult;
  int32_t minor = int32_t(k & ((uint64_t(1) << minor_bits) - 1));
  int32_t major = int32_t(k >> minor_bits);
  int32_t u = major_is_user ? major : minor;
  int32_t i = major_is_user ? minor : major;
  float w_a, w_c;
  edge_weight(wa, u, i, &w_a, &w_c);
  float s_a = w_a, s_c = w_c;
  for (int m = 1; m < mult; ++m) {
    s_a = __fadd_rn(s_a, w_a);
    s_c = __fadd_rn(s_c, w_c);
  }
  uint32_t q = pos[p];
  idx[q] = minor;
  val_fwd[q] = major_is_user ? s_a : s_c;
  val_bwd[q] = major_is_user ? s_c : s_a;
}

// indptr of the coalesced pattern from the duplicate-keeping row starts
__global__ void k_coalesced_indptr(const uint32_t* __restrict__ row_start, const uint32_t* __restrict__ pos,
                                   const uint32_t* __restrict__ nnz, int64_t n_entries, int32_t n_rows,
                                   int64_t* __restrict__ indptr) {
  int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r > n_rows) return;
  uint32_t s = row_start[r];
  indptr[r] = (int64_t(s) < n_entries) ? int64_t(pos[s]) : int64_t(*nnz);
}

__global__ void k_finish_counts(const uint32_t* nnz, const unsigned long long* bad, int64_t* out) {
  out[0] = int64_t(*nnz);
  out[1] = int64_t(*bad);
}

static inline unsigned grid_for(int64_t n) { return (unsigned)ceil_div(n > 0 ? n : 1, GB_THREADS); }

static size_t build_ws_bytes(int64_t E, int32_t U, int32_t I) {
  size_t b = 0;
  b += 3 * align_up(size_t(E) * 8);                  // key_ui, key_iu, alt
  b += align_up(size_t(E + 1) * 4);                  // flags / pos
  b += align_up(size_t(int64_t(U) + 1) * 4);         // user row starts (u32)
  b += align_up(size_t(int64_t(I) + 1) * 4);         // item row starts (u32)
  b += radix_sort_temp_bytes(E);
  int64_t m = E + 1;
  if (int64_t(U) + 1 > m) m = int64_t(U) + 1;
  if (int64_t(I) + 1 > m) m = int64_t(I) + 1;
  b += scan_temp_bytes(m);
  b += 1024;
  return b;
}

}  // namespace cgx

using namespace cgx;

extern "C" size_t cgx_graph_build_workspace_bytes(int64_t num_edges, int32_t num_users, int32_t num_items) {
  return build_ws_bytes(num_edges, num_users, num_items);
}

extern "C" int cgx_graph_build(const int32_t* edges_u, const int32_t* edges_i, int64_t E, int32_t U, int32_t I,
                               const float* cred, int variant, const float* alpha,
                               const int32_t* deg_i_weights, int32_t* deg_u,
                               int32_t* deg_i, int64_t* samp_indptr, int32_t* samp_idx,
                               int64_t* user_indptr, int32_t* user_idx, float* user_val_fwd,
                               float* user_val_bwd, int64_t* item_indptr, int32_t* item_idx,
                               float* item_val_fwd, float* item_val_bwd, int64_t* nnz_out, int deg_only,
                               void* workspace, size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(E >= 0 && U > 0 && I > 0, CGX_ERR_ARG, "graph_build: bad sizes E=%lld U=%d I=%d", (long long)E, U, I);
  CGX_REQUIRE(E < (int64_t(1) << 31), CGX_ERR_ARG, "graph_build: E must be < 2^31");
  CGX_REQUIRE((E == 0 || (edges_u && edges_i)) && deg_u && deg_i && nnz_out, CGX_ERR_ARG,
              "graph_build: NULL pointer");
  CGX_REQUIRE(variant >= CGX_VARIANT_CU && variant <= CGX_VARIANT_DA, CGX_ERR_ARG, "graph_build: bad variant %d",
              variant);
  CGX_REQUIRE(workspace_bytes >= build_ws_bytes(E, U, I), CGX_ERR_WORKSPACE, "graph_build: workspace too small");
  if (!deg_only) {
    CGX_REQUIRE(cred && samp_indptr && samp_idx && user_indptr && user_idx && user_val_fwd && user_val_bwd &&
                    item_indptr && item_idx && item_val_fwd && item_val_bwd,
                CGX_ERR_ARG, "graph_build: NULL pointer");
    CGX_REQUIRE(variant != CGX_VARIANT_DA || alpha != nullptr, CGX_ERR_ARG,
                "graph_build: the degree-aware variant needs alpha");
  }
  const int bits_u = bits_for(U), bits_i = bits_for(I);

  Arena ws(workspace, workspace_bytes);
  uint64_t* key_ui = ws.take<uint64_t>(E);
  uint64_t* key_iu = ws.take<uint64_t>(E);
  uint64_t* alt = ws.take<uint64_t>(E);
  uint32_t* flags = ws.take<uint32_t>(E + 1);
  uint32_t* ustart = ws.take<uint32_t>(int64_t(U) + 1);
  uint32_t* istart = ws.take<uint32_t>(int64_t(I) + 1);
  size_t sort_bytes = radix_sort_temp_bytes(E);
  void* sort_tmp = ws.take<char>(sort_bytes);
  int64_t m = E + 1;
  if (int64_t(U) + 1 > m) m = int64_t(U) + 1;
  if (int64_t(I) + 1 > m) m = int64_t(I) + 1;
  size_t scan_bytes = scan_temp_bytes(m);
  void* scan_tmp = ws.take<char>(scan_bytes);
  uint32_t* nnz_dev = ws.take<uint32_t>(2);
  unsigned long long* bad = ws.take<unsigned long long>(1);
  CGX_REQUIRE(ws.ok, CGX_ERR_WORKSPACE, "graph_build: workspace too small");

  CGX_CUDA(cudaMemsetAsync(deg_u, 0, size_t(U) * 4, stream));
  CGX_CUDA(cudaMemsetAsync(deg_i, 0, size_t(I) * 4, stream));
  CGX_CUDA(cudaMemsetAsync(bad, 0, 8, stream));
  CGX_CUDA(cudaMemsetAsync(nnz_dev, 0, 8, stream));
  if (E > 0) {
    k_pack_and_count<<<grid_for(E), GB_THREADS, 0, stream>>>(edges_u, edges_i, E, U, I, bits_u, bits_i,
                                                            deg_only ? nullptr : key_ui,
                                                            deg_only ? nullptr : key_iu, deg_u, deg_i, bad);
    CGX_LAUNCH_CHECK();
  }
  if (deg_only) {
    k_finish_counts<<<1, 1, 0, stream>>>(nnz_dev, bad, nnz_out);
    CGX_LAUNCH_CHECK();
    return CGX_OK;
  }

  // duplicate-keeping row starts = exclusive scan of the degrees (+ total at [n_rows])
  CGX_TRY(exclusive_scan_u32(reinterpret_cast<const uint32_t*>(deg_u), ustart, U, ustart + U, scan_tmp,
                             scan_bytes, stream));
  CGX_TRY(exclusive_scan_u32(reinterpret_cast<const uint32_t*>(deg_i), istart, I, istart + I, scan_tmp,
                             scan_bytes, stream));
  k_widen<<<grid_for(int64_t(U) + 1), GB_THREADS, 0, stream>>>(ustart, samp_indptr, int64_t(U) + 1);
  CGX_LAUNCH_CHECK();

  // item degrees entering the weight formulas: the local histogram, or (user-sharded builds) the
  // degrees over ALL shards supplied by the caller
  WeightArgs wa{deg_u, deg_i_weights ? deg_i_weights : deg_i, cred, alpha, variant};
  for (int pass = 0; pass < 2; ++pass) {
    const bool by_user = pass == 0;
    uint64_t* keys = by_user ? key_ui : key_iu;
    const int minor_bits = by_user ? bits_i : bits_u;
    const int32_t n_rows = by_user ? U : I;
    uint64_t* sorted = keys;
    CGX_TRY(radix_sort_u64(keys, alt, E, bits_u + bits_i, sort_tmp, sort_bytes, stream, &sorted));
    if (by_user && E > 0) {
      k_unpack_minor<<<grid_for(E), GB_THREADS, 0, stream>>>(sorted, E, minor_bits, samp_idx);
      CGX_LAUNCH_CHECK();
    }
    if (E > 0) {
      k_head_flags<<<grid_for(E), GB_THREADS, 0, stream>>>(sorted, E, flags);
      CGX_LAUNCH_CHECK();
    }
    CGX_TRY(exclusive_scan_u32(flags, flags, E, nnz_dev, scan_tmp, scan_bytes, stream));
    if (E > 0) {
      k_emit<<<grid_for(E), GB_THREADS, 0, stream>>>(sorted, flags, E, minor_bits, by_user ? 1 : 0, wa,
                                                    by_user ? user_idx : item_idx,
                                                    by_user ? user_val_fwd : item_val_fwd,
                                                    by_user ? user_val_bwd : item_val_bwd);
      CGX_LAUNCH_CHECK();
    }
    k_coalesced_indptr<<<grid_for(int64_t(n_rows) + 1), GB_THREADS, 0, stream>>>(
        by_user ? ustart : istart, flags, nnz_dev, E, n_rows, by_user ? user_indptr : item_indptr);
    CGX_LAUNCH_CHECK();
    // if the sort left its result in `alt`, the next pass must not clobber it before use: it does
    // not -- pass 1 sorts key_iu with `alt` as scratch only after pass 0 has consumed `sorted`.
  }
  k_finish_counts<<<1, 1, 0, stream>>>(nnz_dev, bad, nnz_out);
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}

extern "C" int cgx_user_csr(const int32_t* edges_u, const int32_t* edges_i, int64_t E, int32_t U, int32_t I,
                            int64_t* indptr, int32_t* idx, void* workspace, size_t workspace_bytes,
                            void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(E >= 0 && U > 0 && I > 0 && E < (int64_t(1) << 31), CGX_ERR_ARG, "user_csr: bad sizes");
  CGX_REQUIRE(indptr && (E == 0 || (edges_u && edges_i && idx)), CGX_ERR_ARG, "user_csr: NULL pointer");
  CGX_REQUIRE(workspace_bytes >= build_ws_bytes(E, U, I), CGX_ERR_WORKSPACE, "user_csr: workspace too small");
  const int bits_u = bits_for(U), bits_i = bits_for(I);
  Arena ws(workspace, workspace_bytes);
  uint64_t* key_ui = ws.take<uint64_t>(E);
  uint64_t* alt = ws.take<uint64_t>(E);
  int32_t* deg = ws.take<int32_t>(int64_t(U) + 1);
  uint32_t* ustart = ws.take<uint32_t>(int64_t(U) + 1);
  size_t sort_bytes = radix_sort_temp_bytes(E);
  void* sort_tmp = ws.take<char>(sort_bytes);
  size_t scan_bytes = scan_temp_bytes(int64_t(U) + 1);
  void* scan_tmp = ws.take<char>(scan_bytes);
  unsigned long long* bad = ws.take<unsigned long long>(1);
  CGX_REQUIRE(ws.ok, CGX_ERR_WORKSPACE, "user_csr: workspace too small");
  CGX_CUDA(cudaMemsetAsync(deg, 0, size_t(U) * 4, stream));
  CGX_CUDA(cudaMemsetAsync(bad, 0, 8, stream));
  if (E > 0) {
    k_pack_and_count<<<grid_for(E), GB_THREADS, 0, stream>>>(edges_u, edges_i, E, U, I, bits_u, bits_i, key_ui,
                                                            nullptr, deg, nullptr, bad);
    CGX_LAUNCH_CHECK();
  }
  CGX_TRY(exclusive_scan_u32(reinterpret_cast<const uint32_t*>(deg), ustart, U, ustart + U, scan_tmp, scan_bytes,
                             stream));
  k_widen<<<grid_for(int64_t(U) + 1), GB_THREADS, 0, stream>>>(ustart, indptr, int64_t(U) + 1);
  CGX_LAUNCH_CHECK();
  uint64_t* sorted = key_ui;
  CGX_TRY(radix_sort_u64(key_ui, alt, E, bits_u + bits_i, sort_tmp, sort_bytes, stream, &sorted));
  if (E > 0) {
    k_unpack_minor<<<grid_for(E), GB_THREADS, 0, stream>>>(sorted, E, bits_i, idx);
    CGX_LAUNCH_CHECK();
  }
  return CGX_OK;
}

// ---- row schedule -------------------------------------------------------------------------------
// The SpMM walks rows in DESCENDING degree order (perm): rows of similar length share a warp and a
// CTA (no intra-block idling), and the expensive rows start first (no tail).  Rows longer than
// CGX_LONG_ROW come first in perm and are cut into CGX_CHUNK-sized chunks, each a work item of its
// own; their partial sums are combined in chunk order by a small finishing kernel.
namespace cgx {
__global__ void k_sched_keys(const int64_t* __restrict__ indptr, int32_t n_rows, int bits_r,
                             uint64_t* __restrict__ keys) {
  int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= n_rows) return;
  const uint64_t deg = uint64_t(indptr[r + 1] - indptr[r]);
  keys[r] = ((uint64_t(0x7fffffff) - deg) << bits_r) | uint64_t(r);   // ascending key = descending degree
}
__global__ void k_sched_perm(const uint64_t* __restrict__ keys, const int64_t* __restrict__ indptr,
                             int32_t n_rows, int bits_r, int32_t* __restrict__ perm,
                             uint32_t* __restrict__ is_long, uint32_t* __restrict__ n_chunks,
                             unsigned int* __restrict__ n_huge) {
  int64_t k = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (k >= n_rows) return;
  const int32_t r = int32_t(keys[k] & ((uint64_t(1) << bits_r) - 1));
  perm[k] = r;
  const int64_t len = indptr[r + 1] - indptr[r];
  const bool lg = len > CGX_LONG_ROW;
  is_long[k] = lg ? 1u : 0u;
  n_chunks[k] = lg ? uint32_t((len + CGX_CHUNK - 1) / CGX_CHUNK) : 0u;
  if (len > CGX_HUGE_ROW) atomicAdd(n_huge, 1u);   // integer count: order independent
}
__global__ void k_sched_chunks(const uint32_t* __restrict__ cpos, const uint32_t* __restrict__ totals,
                               int32_t n_long, int32_t* __restrict__ chunk_ptr, int32_t* __restrict__ chunk_row) {
  // cpos = exclusive scan of chunk counts in perm order; long rows are exactly perm[0 .. n_long)
  int64_t c = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n_ch = totals[1];
  if (c <= n_long) chunk_ptr[c] = (c < n_long) ? int32_t(cpos[c]) : int32_t(n_ch);
  if (c >= n_ch) return;
  int lo = 0, hi = n_long;   // largest k with cpos[k] <= c
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (int64_t(cpos[mid]) <= c) lo = mid; else hi = mid;
  }
  chunk_row[c] = lo;
}
struct SchedWs {
  uint64_t *keys, *alt;
  uint32_t *is_long, *cpos, *totals;
  void *sort_tmp, *scan_tmp;
  size_t sort_bytes, scan_bytes;
};
static size_t sched_ws_bytes(int32_t n_rows) {
  return 2 * align_up(size_t(n_rows) * 8) + 2 * align_up(size_t(n_rows) * 4) + 256 + radix_sort_temp_bytes(n_rows) +
         scan_temp_bytes(n_rows) + 1024;
}
static int sched_carve(void* workspace, size_t workspace_bytes, int32_t n_rows, SchedWs* w) {
  Arena ws(workspace, workspace_bytes);
  w->keys = ws.take<uint64_t>(n_rows);
  w->alt = ws.take<uint64_t>(n_rows);
  w->is_long = ws.take<uint32_t>(n_rows);
  w->cpos = ws.take<uint32_t>(n_rows);
  w->totals = ws.take<uint32_t>(4);
  w->sort_bytes = radix_sort_temp_bytes(n_rows);
  w->sort_tmp = ws.take<char>(w->sort_bytes);
  w->scan_bytes = scan_temp_bytes(n_rows);
  w->scan_tmp = ws.take<char>(w->scan_bytes);
  CGX_REQUIRE(ws.ok, CGX_ERR_WORKSPACE, "row_schedule: workspace too small");
  return CGX_OK;
}
}  // namespace cgx

extern "C" size_t cgx_row_schedule_workspace_bytes(int32_t n_rows) { return sched_ws_bytes(n_rows); }

extern "C" int cgx_row_schedule(const int64_t* indptr, int32_t n_rows, int32_t* perm, int32_t* n_long_host,
                                int32_t* n_chunks_host, int32_t* n_huge_host, void* workspace,
                                size_t workspace_bytes, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(indptr && n_rows > 0 && perm && n_long_host && n_chunks_host && n_huge_host, CGX_ERR_ARG,
              "row_schedule: bad argument");
  SchedWs w;
  CGX_TRY(sched_carve(workspace, workspace_bytes, n_rows, &w));
  const int bits_r = bits_for(n_rows);
  k_sched_keys<<<grid_for(n_rows), GB_THREADS, 0, stream>>>(indptr, n_rows, bits_r, w.keys);
  CGX_LAUNCH_CHECK();
  uint64_t* sorted = w.keys;
  CGX_TRY(radix_sort_u64(w.keys, w.alt, n_rows, bits_r + 31, w.sort_tmp, w.sort_bytes, stream, &sorted));
  CGX_CUDA(cudaMemsetAsync(w.totals, 0, 16, stream));
  k_sched_perm<<<grid_for(n_rows), GB_THREADS, 0, stream>>>(sorted, indptr, n_rows, bits_r, perm, w.is_long, w.cpos,
                                                           w.totals + 2);
  CGX_LAUNCH_CHECK();
  CGX_TRY(exclusive_scan_u32(w.is_long, w.is_long, n_rows, w.totals, w.scan_tmp, w.scan_bytes, stream));
  CGX_TRY(exclusive_scan_u32(w.cpos, w.cpos, n_rows, w.totals + 1, w.scan_tmp, w.scan_bytes, stream));
  uint32_t h[4];
  CGX_CUDA(cudaMemcpyAsync(h, w.totals, 16, cudaMemcpyDeviceToHost, stream));
  CGX_CUDA(cudaStreamSynchronize(stream));
  *n_long_host = int32_t(h[0]);
  *n_chunks_host = int32_t(h[1]);
  *n_huge_host = int32_t(h[2]);
  return CGX_OK;
}

extern "C" int cgx_row_schedule_chunks(int32_t n_rows, int32_t n_long, int32_t n_chunks, int32_t* chunk_ptr,
                                       int32_t* chunk_row, void* workspace, size_t workspace_bytes,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(n_rows > 0 && n_long > 0 && n_chunks > 0 && chunk_ptr && chunk_row, CGX_ERR_ARG,
              "row_schedule_chunks: bad argument");
  SchedWs w;   // same workspace, untouched since cgx_row_schedule: cpos / totals are still in place
  CGX_TRY(sched_carve(workspace, workspace_bytes, n_rows, &w));
  const int64_t n = n_chunks > n_long + 1 ? n_chunks : n_long + 1;
  k_sched_chunks<<<grid_for(n), GB_THREADS, 0, stream>>>(w.cpos, w.totals, n_long, chunk_ptr, chunk_row);
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}

namespace cgx {
__global__ void k_sched_work(const int64_t* __restrict__ indptr, const int32_t* __restrict__ perm, int32_t n_rows,
                             int32_t n_long, int32_t n_chunks, const int32_t* __restrict__ chunk_ptr,
                             const int32_t* __restrict__ chunk_row, int4* __restrict__ work) {
  const int64_t item = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int64_t n_items = int64_t(n_chunks) + (n_rows - n_long);
  if (item >= n_items) return;
  int64_t begin;
  int32_t len, tag;
  if (item < n_chunks) {
    const int32_t k = chunk_row[item];
    const int32_t row = perm[k];
    begin = indptr[row] + int64_t(int32_t(item) - chunk_ptr[k]) * CGX_CHUNK;
    const int64_t left = indptr[row + 1] - begin;
    len = int32_t(left < CGX_CHUNK ? left : CGX_CHUNK);
    tag = k;
  } else {
    const int32_t row = perm[item - n_chunks + n_long];
    begin = indptr[row];
    len = int32_t(indptr[row + 1] - begin);
    tag = row;
  }
  work[item] = make_int4(int32_t(uint32_t(begin)), int32_t(begin >> 32), len, tag);
}
}  // namespace cgx

extern "C" int cgx_row_schedule_work(const int64_t* indptr, const int32_t* perm, int32_t n_rows, int32_t n_long,
                                     int32_t n_chunks, const int32_t* chunk_ptr, const int32_t* chunk_row,
                                     void* work, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(indptr && perm && work && n_rows > 0 && n_long >= 0 && n_chunks >= 0, CGX_ERR_ARG,
              "row_schedule_work: bad argument");
  CGX_REQUIRE(n_long == 0 || (chunk_ptr && chunk_row), CGX_ERR_ARG, "row_schedule_work: chunk tables missing");
  const int64_t n_items = int64_t(n_chunks) + (n_rows - n_long);
  k_sched_work<<<grid_for(n_items), GB_THREADS, 0, stream>>>(indptr, perm, n_rows, n_long, n_chunks, chunk_ptr,
                                                            chunk_row, static_cast<int4*>(work));
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}
