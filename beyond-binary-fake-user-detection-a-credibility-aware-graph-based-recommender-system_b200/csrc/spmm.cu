// CSR SpMM with a fused layer-mean / gradient-seed epilogue, and the K-layer propagation
// schedules (forward and adjoint) built on it.
//
// Replaces (reference, /root/reference): torch.sparse.mm at lightgcn_cu.py:431,434 and
// Version-2/lighgcn_cu_pop.py:483-484, the stack().mean() at lightgcn_cu.py:446-447 /
// lighgcn_cu_pop.py:488-489, and autograd's SparseAddmmBackward for the same calls.
//
// Kernel shape: the path is an HBM/L2-bound gather.  A group of G = d/4 lanes owns one work item
// (d=64: half a warp, two items per warp; d=128: one warp); every lane keeps one float4 of the
// output row in registers, column ids/values are fetched coalesced G at a time (the next batch is
// prefetched while the current one is consumed) and broadcast by shuffle, and the embedding-row
// gathers are issued UNR at a time before any FMA so that each lane keeps UNR independent 16-byte
// loads in flight.
// Work items follow the schedule built by cgx_row_schedule: first the CGX_CHUNK-sized chunks of the
// rows longer than CGX_LONG_ROW (partials in workspace, summed IN CHUNK ORDER by the chunk that
// arrives last, or by a finishing kernel for rows above CGX_HUGE_ROW), then all other rows in
// descending degree order.  Neighbouring groups therefore carry equal work (no idle lanes inside a
// warp or CTA) and the heavy items start first (no tail).
// No floating-point atomics anywhere (the only atomic is an integer arrival counter): results are
// bitwise reproducible.
// Tuning (profiles/r1_spmm_variants.txt): 8 gathers in flight per lane with the register budget
// capped for 4 CTAs/SM is within 3 % of the best variant on both the L2-resident C2 shape and the
// HBM-bound 64M-edge shape; higher unrolls lose occupancy (d=128: 86 registers -> 2 CTAs/SM).
#include <stdlib.h>

#include "common.cuh"

namespace cgx {

constexpr int SP_THREADS = 256;

// Tuning knobs of the gather loop (selected per width in spmm_dispatch; CGX_SPMM_VARIANT overrides
// them for experiments): UNR = embedding-row gathers in flight per lane, HINT = cache policy of
// those gathers, MINB = CTAs per SM the register allocation must allow.
enum { HINT_NC = 0, HINT_CG = 1, HINT_NC_NOALLOC = 2 };

template <int HINT>
__device__ __forceinline__ float4 ld_row(const float4* p) {
  if (HINT == HINT_CG) return __ldcg(p);
  if (HINT == HINT_NC_NOALLOC) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
  }
  return __ldg(p);
}
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

// acc[v] (v < V) += sum over nnz in [begin, end) of val * X[idx, :]; one group, lane = 0..G-1.
// NZ: `nz[c] == 0` marks row c of X as all-zero (the loss gradient touches <= 3 * batch rows): its weight is
// forced to 0 and rows of weight 0 are not loaded -- same sum, a fraction of the gather traffic.
template <int G, int V, int UNR, int HINT, bool NZ = false>
__device__ __forceinline__ void gather_batches(const int32_t* __restrict__ idx, const float* __restrict__ val,
                                               int64_t begin, int64_t end, const float4* __restrict__ X, int lane,
                                               unsigned mask, int32_t c_nxt, float w_nxt, float4 (&acc)[V],
                                               const uint8_t* __restrict__ nz = nullptr) {
  // (c_nxt, w_nxt) = this lane's column id / value of the first batch, already loaded by the caller
  constexpr int ROW4 = G * V;  // float4 per embedding row
  for (int64_t base = begin; base < end; base += G) {
    const int32_t c = c_nxt;
    float w = w_nxt;
    if (NZ) {
      if (w != 0.f && __ldg(nz + c) == 0) w = 0.f;
    }
    const int64_t pn = base + G + lane;  // prefetch the next batch of column ids / values
    if (pn < end) {
      c_nxt = __ldg(idx + pn);
      w_nxt = __ldg(val + pn);
    }
    const int cnt = (end - base) < G ? int(end - base) : G;
    for (int j0 = 0; j0 < cnt; j0 += UNR) {
      float4 x[UNR][V];
      float ww[UNR];
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
        const int j = j0 + t;
        const int src = (j < G) ? j : (G - 1);
        const int32_t cj = __shfl_sync(mask, c, src, G);
        ww[t] = __shfl_sync(mask, w, src, G);
        if (j < cnt && (!NZ || ww[t] != 0.f)) {
#pragma unroll
          for (int v = 0; v < V; ++v) x[t][v] = ld_row<HINT>(X + int64_t(cj) * ROW4 + v * G + lane);
        } else {
          ww[t] = 0.f;
#pragma unroll
          for (int v = 0; v < V; ++v) x[t][v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int t = 0; t < UNR; ++t) {
#pragma unroll
        for (int v = 0; v < V; ++v) fma4(acc[v], ww[t], x[t][v]);
      }
    }
  }
}

template <int G, int V, int UNR, int HINT>
__device__ __forceinline__ void gather_range(const int32_t* __restrict__ idx, const float* __restrict__ val,
                                             int64_t begin, int64_t end, const float4* __restrict__ X, int lane,
                                             unsigned mask, float4 (&acc)[V]) {
  int32_t c = 0;
  float w = 0.f;
  if (begin + lane < end) {
    c = __ldg(idx + begin + lane);
    w = __ldg(val + begin + lane);
  }
  gather_batches<G, V, UNR, HINT>(idx, val, begin, end, X, lane, mask, c, w, acc);
}

// Push mode of the user-sharded propagation (cgx_spmm_push): output row r is not stored to Y but straight into the
// staging area of the rank that OWNS row r (rows_per consecutive rows per rank), slot `rank` -- a posted store over
// NVLink, issued row by row while the rest of the product is still being gathered.  The exchange kernel
// (comm.cu, pushed form) then sums the slots with local reads only.
__constant__ char* c_push_base[16];        // peer-mapped communication buffers, index = rank (cgx_spmm_set_push_peers)
struct SpmmPush {
  unsigned long long off;                  // byte offset of the staging area inside every communication buffer
  int32_t rows_per;                        // rows owned by each rank; 0 = push mode off
  int32_t rank;
};

template <int G, int V>
__device__ __forceinline__ void epilogue(int64_t row, int lane, const float4 (&y)[V], float4* __restrict__ Y,
                                         const float4* ACC_IN, float4* ACC_OUT, float acc_scale,
                                         const SpmmPush ps = SpmmPush{0ull, 0, 0}) {
  constexpr int ROW4 = G * V;
  float4* ypush = nullptr;
  if (ps.rows_per > 0) {
    const int owner = int(row / ps.rows_per);
    ypush = reinterpret_cast<float4*>(c_push_base[owner] + ps.off) +
            (int64_t(ps.rank) * ps.rows_per + (row - int64_t(owner) * ps.rows_per)) * ROW4;
  }
#pragma unroll
  for (int v = 0; v < V; ++v) {
    const int64_t o = row * ROW4 + v * G + lane;
    if (ypush) ypush[v * G + lane] = y[v];
    if (Y) Y[o] = y[v];
    if (ACC_OUT) {
      float4 a = ACC_IN ? ACC_IN[o] : make_float4(0.f, 0.f, 0.f, 0.f);
      ACC_OUT[o] = scale4(add4(a, y[v]), acc_scale);
    }
  }
}

struct SpmmSched {
  const int32_t* perm;
  const int32_t* chunk_ptr;
  const int32_t* chunk_row;
  int32_t* arrive;
  int32_t n_long, n_chunks, n_huge;
};

// One group per work item: chunk items first, then rows in descending degree.  A chunk stores its
// partial sum; for rows up to CGX_HUGE_ROW the chunk that arrives last (per-row counter, release /
// acquire through __threadfence) adds the partials IN CHUNK ORDER and runs the epilogue, so the
// result does not depend on which chunk happened to be last.
template <int G, int V, int UNR, int HINT, int MINB, bool NZ = false, bool PUSH = false>
__global__ void __launch_bounds__(SP_THREADS, MINB) k_spmm(const int64_t* __restrict__ indptr,
                                                           const int32_t* __restrict__ idx,
                                                           const float* __restrict__ val, int32_t n_rows,
                                                           SpmmSched sc, const float4* __restrict__ X,
                                                           float4* __restrict__ Y, const float4* ACC_IN,
                                                           float4* ACC_OUT, float acc_scale, float4* partial,
                                                           const uint8_t* __restrict__ nz,
                                                           const int4* __restrict__ work, const SpmmPush ps_in) {
  // PUSH = false: a compile-time "off", so that the ordinary instantiations carry no trace of the push path
  const SpmmPush ps = PUSH ? ps_in : SpmmPush{0ull, 0, 0};
  constexpr int ROW4 = G * V;
  // Programmatic dependent launch: let the next kernel of the stream start its own prologue now, and run
  // THIS kernel's prologue (schedule lookups, first batch of column ids / values -- graph constants) while the
  // previous kernel is still draining.  Nothing the previous kernel wrote is read, and nothing is written,
  // before griddepcontrol.wait returns.
  asm volatile("griddepcontrol.launch_dependents;");
  const int lane = threadIdx.x & (G - 1);
  const int64_t item = (int64_t(blockIdx.x) * SP_THREADS + threadIdx.x) / G;
  const unsigned mask = group_mask<G>();
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (work != nullptr) {
    // Flattened schedule (cgx_row_schedule_work): ONE 16-byte descriptor {begin, length, row | long-row index}
    // replaces the perm -> indptr chain (two dependent loads) in front of every row's first gather.
    const int64_t n_items = int64_t(sc.n_chunks) + (n_rows - sc.n_long);
    if (item >= n_items) return;
    const int4 wd = __ldg(work + item);
    const int64_t begin = (int64_t(wd.y) << 32) | uint32_t(wd.x);
    const int64_t end = begin + wd.z;
    int32_t cf = 0;
    float wf = 0.f;
    if (lane < wd.z) {
      cf = __ldg(idx + begin + lane);
      wf = __ldg(val + begin + lane);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    gather_batches<G, V, UNR, HINT, NZ>(idx, val, begin, end, X, lane, mask, cf, wf, acc, nz);
    if (item >= sc.n_chunks) {
      // (loading ACC_IN[row] before the gathers, to take it off the dependency chain, was measured SLOWER: the
      // four extra live registers spill at the 64-register cap -- C2 0.514 -> 0.553 ms, C4 155 -> 160 ms)
      epilogue<G, V>(wd.w, lane, acc, Y, ACC_IN, ACC_OUT, acc_scale, ps);
      return;
    }
    const int32_t k = wd.w;
#pragma unroll
    for (int v = 0; v < V; ++v) __stcg(partial + item * ROW4 + v * G + lane, acc[v]);
    if (k < sc.n_huge) return;            // combined by k_spmm_finish
    const int32_t c0 = __ldg(sc.chunk_ptr + k), c1 = __ldg(sc.chunk_ptr + k + 1);
    __threadfence();                       // release: this group's partial is visible device-wide
    __syncwarp(mask);
    int prev = 0;
    if (lane == 0) prev = atomicAdd(sc.arrive + k, 1);
    prev = __shfl_sync(mask, prev, 0, G);
    if (prev != c1 - c0 - 1) return;
    __threadfence();                       // acquire: every other chunk's partial is visible
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c0; c < c1; ++c) {
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], __ldcg(partial + int64_t(c) * ROW4 + v * G + lane));
    }
    epilogue<G, V>(int64_t(__ldg(sc.perm + k)), lane, acc, Y, ACC_IN, ACC_OUT, acc_scale, ps);
    if (lane == 0) sc.arrive[k] = 0;       // self-resetting for the next launch
    return;
  }
  if (item < sc.n_chunks) {
    const int32_t k = __ldg(sc.chunk_row + item);
    const int32_t row = __ldg(sc.perm + k);
    const int32_t c0 = __ldg(sc.chunk_ptr + k);
    const int64_t rbeg = __ldg(indptr + row), rend = __ldg(indptr + row + 1);
    const int64_t begin = rbeg + int64_t(int32_t(item) - c0) * CGX_CHUNK;
    const int64_t end = begin + CGX_CHUNK < rend ? begin + CGX_CHUNK : rend;
    int32_t cf = 0;
    float wf = 0.f;
    if (begin + lane < end) {
      cf = __ldg(idx + begin + lane);
      wf = __ldg(val + begin + lane);
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
    gather_batches<G, V, UNR, HINT, NZ>(idx, val, begin, end, X, lane, mask, cf, wf, acc, nz);
#pragma unroll
    for (int v = 0; v < V; ++v) __stcg(partial + item * ROW4 + v * G + lane, acc[v]);
    if (k < sc.n_huge) return;            // combined by k_spmm_finish
    const int32_t c1 = __ldg(sc.chunk_ptr + k + 1);
    __threadfence();                       // release: this group's partial is visible device-wide
    __syncwarp(mask);
    int prev = 0;
    if (lane == 0) prev = atomicAdd(sc.arrive + k, 1);
    prev = __shfl_sync(mask, prev, 0, G);
    if (prev != c1 - c0 - 1) return;
    __threadfence();                       // acquire: every other chunk's partial is visible
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int c = c0; c < c1; ++c) {
#pragma unroll
      for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], __ldcg(partial + int64_t(c) * ROW4 + v * G + lane));
    }
    epilogue<G, V>(row, lane, acc, Y, ACC_IN, ACC_OUT, acc_scale, ps);
    if (lane == 0) sc.arrive[k] = 0;       // self-resetting for the next launch
    return;
  }
  const int64_t r = item - sc.n_chunks + sc.n_long;
  if (r >= n_rows) return;
  const int32_t row = __ldg(sc.perm + r);
  const int64_t begin = __ldg(indptr + row), end = __ldg(indptr + row + 1);
  int32_t cf = 0;
  float wf = 0.f;
  if (begin + lane < end) {
    cf = __ldg(idx + begin + lane);
    wf = __ldg(val + begin + lane);
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  gather_batches<G, V, UNR, HINT, NZ>(idx, val, begin, end, X, lane, mask, cf, wf, acc, nz);
  epilogue<G, V>(row, lane, acc, Y, ACC_IN, ACC_OUT, acc_scale, ps);
}

// Persistent, software-pipelined form of k_spmm: a fixed grid of groups walks the flattened work list
// (cgx_row_schedule_work) round-robin -- group g takes items g, g + n_groups, ... so every round hands
// neighbouring groups items of equal weight, heaviest rounds first -- and while an item's rows are
// being gathered the NEXT item's 16-byte descriptor and its first batch of column ids / values are
// already in flight.  That removes the per-row pointer chase (perm -> indptr -> idx -> rows) from the
// critical path, which is what bounds rows of 10-40 non-zeros.
template <int G, int V, int UNR, int HINT, int MINB>
__global__ void __launch_bounds__(SP_THREADS, MINB) k_spmm_p(const int4* __restrict__ work, int64_t n_items,
                                                             const int32_t* __restrict__ idx,
                                                             const float* __restrict__ val, SpmmSched sc,
                                                             const float4* __restrict__ X, float4* __restrict__ Y,
                                                             const float4* ACC_IN, float4* ACC_OUT, float acc_scale,
                                                             float4* partial) {
  constexpr int ROW4 = G * V;
  const int lane = threadIdx.x & (G - 1);
  const unsigned mask = group_mask<G>();
  const int64_t n_groups = int64_t(gridDim.x) * (SP_THREADS / G);
  int64_t item = (int64_t(blockIdx.x) * SP_THREADS + threadIdx.x) / G;
  if (item >= n_items) return;
  int4 cur = __ldg(work + item);
  int64_t begin = (int64_t(cur.y) << 32) | uint32_t(cur.x);
  int32_t c0v = 0;
  float w0v = 0.f;
  if (lane < cur.z) {
    c0v = __ldg(idx + begin + lane);
    w0v = __ldg(val + begin + lane);
  }
  while (true) {
    const int64_t nitem = item + n_groups;
    const bool more = nitem < n_items;
    int4 nxt = make_int4(0, 0, 0, 0);
    if (more) nxt = __ldg(work + nitem);                 // descriptor of the next item: in flight during the gather
    float4 acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_batches<G, V, UNR, HINT>(idx, val, begin, begin + cur.z, X, lane, mask, c0v, w0v, acc);
    const int64_t nbegin = (int64_t(nxt.y) << 32) | uint32_t(nxt.x);
    c0v = 0;
    w0v = 0.f;
    if (more && lane < nxt.z) {                          // first batch of the next item: in flight during the epilogue
      c0v = __ldg(idx + nbegin + lane);
      w0v = __ldg(val + nbegin + lane);
    }
    if (item < sc.n_chunks) {
      const int32_t k = cur.w;
#pragma unroll
      for (int v = 0; v < V; ++v) __stcg(partial + item * ROW4 + v * G + lane, acc[v]);
      if (k >= sc.n_huge) {                              // rows above CGX_HUGE_ROW are combined by k_spmm_finish
        const int32_t ch0 = __ldg(sc.chunk_ptr + k), ch1 = __ldg(sc.chunk_ptr + k + 1);
        __threadfence();                                 // release: this group's partial is visible device-wide
        __syncwarp(mask);
        int prev = 0;
        if (lane == 0) prev = atomicAdd(sc.arrive + k, 1);
        prev = __shfl_sync(mask, prev, 0, G);
        if (prev == ch1 - ch0 - 1) {                     // last chunk of the row: add the partials in chunk order
          __threadfence();
#pragma unroll
          for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
          for (int c = ch0; c < ch1; ++c) {
#pragma unroll
            for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], __ldcg(partial + int64_t(c) * ROW4 + v * G + lane));
          }
          epilogue<G, V>(int64_t(__ldg(sc.perm + k)), lane, acc, Y, ACC_IN, ACC_OUT, acc_scale);
          if (lane == 0) sc.arrive[k] = 0;               // self-resetting for the next launch
        }
      }
    } else {
      epilogue<G, V>(int64_t(cur.w), lane, acc, Y, ACC_IN, ACC_OUT, acc_scale);
    }
    if (!more) break;
    item = nitem;
    cur = nxt;
    begin = nbegin;
  }
}

// one CTA per huge row: groups sum interleaved chunk partials, fixed-order reduction, epilogue
template <int G, int V>
__global__ void __launch_bounds__(SP_THREADS) k_spmm_finish(SpmmSched sc, const float4* __restrict__ partial,
                                                            float4* __restrict__ Y, const float4* ACC_IN,
                                                            float4* ACC_OUT, float acc_scale, const SpmmPush ps) {
  constexpr int GROUPS = SP_THREADS / G;
  constexpr int ROW4 = G * V;
  __shared__ float4 red[GROUPS][ROW4];
  const int k = blockIdx.x;
  const int lane = threadIdx.x & (G - 1), grp = threadIdx.x / G;
  const int c0 = __ldg(sc.chunk_ptr + k), c1 = __ldg(sc.chunk_ptr + k + 1);
  float4 acc[V];
#pragma unroll
  for (int v = 0; v < V; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int c = c0 + grp; c < c1; c += GROUPS) {
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = add4(acc[v], __ldg(partial + int64_t(c) * ROW4 + v * G + lane));
  }
#pragma unroll
  for (int v = 0; v < V; ++v) red[grp][v * G + lane] = acc[v];
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int v = 0; v < V; ++v) {
      float4 s = red[0][v * G + lane];
      for (int g = 1; g < GROUPS; ++g) s = add4(s, red[g][v * G + lane]);
      acc[v] = s;
    }
    epilogue<G, V>(int64_t(__ldg(sc.perm + k)), lane, acc, Y, ACC_IN, ACC_OUT, acc_scale, ps);
  }
}

__global__ void k_scale(const float4* __restrict__ in, float4* __restrict__ out, int64_t n4, float s) {
  int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p < n4) out[p] = scale4(in[p], s);
}

// flags[r] = 1 when row r of X (row4 float4 per row) has a non-zero entry (-0 counts as zero); flags pre-zeroed
__global__ void k_row_flags(const uint4* __restrict__ X, int64_t n4, int32_t row4, uint8_t* __restrict__ flags) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n4) return;
  const uint4 v = X[p];
  if ((v.x | v.y | v.z | v.w) & 0x7fffffffu) flags[p / row4] = 1;
}

static int row_flags(const float* X, int64_t n_rows, int32_t d, uint8_t* flags, cudaStream_t stream) {
  CGX_CUDA(cudaMemsetAsync(flags, 0, size_t(n_rows), stream));
  const int64_t n4 = n_rows * d / 4;
  if (n4 == 0) return CGX_OK;
  k_row_flags<<<(unsigned)ceil_div(n4, 256), 256, 0, stream>>>(reinterpret_cast<const uint4*>(X), n4, d / 4, flags);
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}

static int spmm_env(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}
// The persistent, software-pipelined kernel is kept as an experiment: measured SLOWER than one group per
// item with hardware CTA scheduling (C2 d=64: 0.62 vs 0.43 ms fwd+bwd; 64M-edge d=128: 65.4 vs 61.2 ms;
// profiles/r1_spmm_variants.txt), so it is off unless CGX_SPMM_PERSISTENT=1.
// CGX_SPMM_WORK=0: k_spmm looks rows up through perm -> indptr instead of the flattened descriptors
static bool spmm_use_work() {
  static const bool on = spmm_env("CGX_SPMM_WORK", 1) != 0;
  return on;
}
static bool spmm_persistent() {
  static const bool v = spmm_env("CGX_SPMM_PERSISTENT", 0) != 0;
  return v;
}
static bool spmm_pdl() {   // programmatic dependent launch of k_spmm (CGX_SPMM_PDL=0 disables)
  static const bool v = spmm_env("CGX_SPMM_PDL", 1) != 0;
  return v;
}
static int spmm_waves() {
  static const int v = spmm_env("CGX_SPMM_WAVES", 1);
  return v < 1 ? 1 : v;
}

template <int G, int V, int UNR, int HINT, int MINB, bool NZ = false, bool PUSH = false>
static int launch_spmm(const cgx_csr* m, const float* val, const float* X, float* Y, const float* ACC_IN,
                       float* ACC_OUT, float acc_scale, void* workspace, size_t workspace_bytes,
                       cudaStream_t stream, const uint8_t* nz = nullptr, const SpmmPush ps = SpmmPush{0ull, 0, 0}) {
  constexpr int GROUPS = SP_THREADS / G;
  float4* partial = nullptr;
  if (m->n_long > 0) {
    size_t need = size_t(m->n_chunks) * G * V * sizeof(float4);
    CGX_REQUIRE(workspace != nullptr && workspace_bytes >= need, CGX_ERR_WORKSPACE,
                "spmm: workspace too small for %d long-row chunks", m->n_chunks);
    partial = static_cast<float4*>(workspace);
  }
  SpmmSched sc{m->perm, m->chunk_ptr, m->chunk_row, m->arrive, m->n_long, m->n_chunks, m->n_huge};
  const int64_t items = int64_t(m->n_chunks) + (m->n_rows - m->n_long);
  if (!NZ && ps.rows_per == 0 && m->work != nullptr && spmm_persistent()) {
    int64_t blocks = ceil_div(items, GROUPS);
    const int64_t resident = int64_t(148) * MINB * spmm_waves();     // CTAs that fit the chip at once (x waves)
    if (blocks > resident) blocks = resident;
    k_spmm_p<G, V, UNR, HINT, MINB><<<(unsigned)blocks, SP_THREADS, 0, stream>>>(
        static_cast<const int4*>(m->work), items, m->idx, val, sc, reinterpret_cast<const float4*>(X),
        reinterpret_cast<float4*>(Y), reinterpret_cast<const float4*>(ACC_IN), reinterpret_cast<float4*>(ACC_OUT),
        acc_scale, partial);
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)ceil_div(items, GROUPS));
    cfg.blockDim = dim3(SP_THREADS);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = spmm_pdl() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CGX_CUDA(cudaLaunchKernelEx(&cfg, k_spmm<G, V, UNR, HINT, MINB, NZ, PUSH>, m->indptr, m->idx, val, m->n_rows, sc,
                                reinterpret_cast<const float4*>(X), reinterpret_cast<float4*>(Y),
                                reinterpret_cast<const float4*>(ACC_IN), reinterpret_cast<float4*>(ACC_OUT), acc_scale,
                                partial, nz, spmm_use_work() ? static_cast<const int4*>(m->work) : nullptr, ps));
  }
  CGX_LAUNCH_CHECK();
  if (m->n_huge > 0) {
    k_spmm_finish<G, V><<<(unsigned)m->n_huge, SP_THREADS, 0, stream>>>(
        sc, partial, reinterpret_cast<float4*>(Y), reinterpret_cast<const float4*>(ACC_IN),
        reinterpret_cast<float4*>(ACC_OUT), acc_scale, ps);
    CGX_LAUNCH_CHECK();
  }
  return CGX_OK;
}

// gathered tables above this many bytes count as "beyond L2" (cgx_spmm_set_l2_table_bytes; default 96 MiB of 126)
static int64_t g_l2_table_bytes = int64_t(96) << 20;

static int spmm_variant() {
  static int v = [] {
    const char* e = getenv("CGX_SPMM_VARIANT");
    return e ? atoi(e) : 0;
  }();
  return v;
}

#define CGX_SPMM_ARGS m, val, X, Y, ACC_IN, ACC_OUT, acc_scale, ws, ws_bytes, stream, nz, ps

static int spmm_dispatch(const cgx_csr* m, int use_bwd, int32_t d, const float* X, float* Y, const float* ACC_IN,
                         float* ACC_OUT, float acc_scale, void* ws, size_t ws_bytes, cudaStream_t stream,
                         const uint8_t* nz = nullptr, const SpmmPush ps = SpmmPush{0ull, 0, 0}) {
  CGX_REQUIRE(m && m->indptr && m->perm && (m->nnz == 0 || (m->idx && m->val_fwd && m->val_bwd)) && X, CGX_ERR_ARG,
              "spmm: NULL pointer");
  CGX_REQUIRE(m->n_long == 0 || (m->chunk_ptr && m->chunk_row && m->arrive), CGX_ERR_ARG,
              "spmm: chunk tables missing");
  CGX_REQUIRE(Y || ACC_OUT || ps.rows_per > 0, CGX_ERR_ARG, "spmm: no output requested");
  const float* val = use_bwd ? m->val_bwd : m->val_fwd;
  if (m->n_rows == 0) return CGX_OK;
  // Group geometry by regime (profiles/r1_spmm_variants.txt).  Gathered table in L2 (C2/C3: latency-bound): d/4 lanes
  // per row, one float4 per lane, 8 gathers in flight.  Gathered table beyond L2 (HBM-bound): d/8 lanes per row, two
  // float4 per lane, 4 gathers in flight -- twice the rows in flight per SM at the same bytes in flight per lane
  // (d = 64: 27.3 vs 30.1 ms, d = 128: 53.8 vs 58.6 ms on the 64M-edge shape; on C2 the same split LOSES, 0.65 vs
  // 0.43 ms at d = 64 and 0.71 vs 0.61 ms at d = 128).
  const bool beyond_l2 = int64_t(m->n_cols) * int64_t(d) * 4 > g_l2_table_bytes;
  if (ps.rows_per > 0) {   // push mode (user-sharded item products): default geometries, with or without row flags
#define CGX_PUSH_CASE(GG, VV, UU)                                                              \
    return nz ? launch_spmm<GG, VV, UU, HINT_NC, 4, true, true>(CGX_SPMM_ARGS)                \
              : launch_spmm<GG, VV, UU, HINT_NC, 4, false, true>(CGX_SPMM_ARGS)
    switch (d) {
      case 16: CGX_PUSH_CASE(4, 1, 4);
      case 32: CGX_PUSH_CASE(8, 1, 8);
      case 64:
        if (beyond_l2) { CGX_PUSH_CASE(8, 2, 4); }
        CGX_PUSH_CASE(16, 1, 8);
      case 128:
        if (beyond_l2) { CGX_PUSH_CASE(16, 2, 4); }
        CGX_PUSH_CASE(32, 1, 8);
      case 256: CGX_PUSH_CASE(32, 2, 4);
      default:
        set_error("spmm: emb_dim %d unsupported (16, 32, 64, 128, 256)", d);
        return CGX_ERR_UNSUPPORTED;
    }
#undef CGX_PUSH_CASE
  }
  if (nz != nullptr) {   // sparse input rows: the default geometry of every width, row loads predicated on nz
    switch (d) {
      case 16: return launch_spmm<4, 1, 4, HINT_NC, 4, true>(CGX_SPMM_ARGS);
      case 32: return launch_spmm<8, 1, 8, HINT_NC, 4, true>(CGX_SPMM_ARGS);
      case 64:
        if (beyond_l2) return launch_spmm<8, 2, 4, HINT_NC, 4, true>(CGX_SPMM_ARGS);
        return launch_spmm<16, 1, 8, HINT_NC, 4, true>(CGX_SPMM_ARGS);
      case 128:
        if (beyond_l2) return launch_spmm<16, 2, 4, HINT_NC, 4, true>(CGX_SPMM_ARGS);
        return launch_spmm<32, 1, 8, HINT_NC, 4, true>(CGX_SPMM_ARGS);
      case 256: return launch_spmm<32, 2, 4, HINT_NC, 4, true>(CGX_SPMM_ARGS);
      default: break;
    }
  }
  const int variant = spmm_variant();
  switch (d) {
    case 16: return launch_spmm<4, 1, 4, HINT_NC, 4>(CGX_SPMM_ARGS);
    case 32: return launch_spmm<8, 1, 8, HINT_NC, 4>(CGX_SPMM_ARGS);
  // The gather-loop / geometry variants measured in profiles/r1_spmm_variants.txt (CGX_SPMM_VARIANT=<id>) are only
  // compiled with `make EXTRA=-DCGX_SPMM_TUNING` (profiles/tune_spmm.py): 40 extra instantiations of k_spmm.
#ifndef CGX_SPMM_TUNING
#define CGX_VARIANTS(GG) return launch_spmm<GG, 1, 8, HINT_NC, 4>(CGX_SPMM_ARGS);
#else
#define CGX_VARIANTS(GG)                                                                   \
      switch (variant) {                                                                   \
        case 1: return launch_spmm<GG, 1, 16, HINT_NC, 1>(CGX_SPMM_ARGS);                  \
        case 2: return launch_spmm<GG, 1, 8, HINT_CG, 1>(CGX_SPMM_ARGS);                   \
        case 3: return launch_spmm<GG, 1, 8, HINT_NC_NOALLOC, 1>(CGX_SPMM_ARGS);           \
        case 4: return launch_spmm<GG, 1, 4, HINT_NC, 6>(CGX_SPMM_ARGS);                   \
        case 5: return launch_spmm<GG, 1, 8, HINT_NC, 5>(CGX_SPMM_ARGS);                   \
        case 6: return launch_spmm<GG, 1, 16, HINT_NC_NOALLOC, 1>(CGX_SPMM_ARGS);          \
        case 7: return launch_spmm<GG, 1, 4, HINT_NC, 8>(CGX_SPMM_ARGS);                   \
        case 8: return launch_spmm<GG, 1, 2, HINT_NC, 8>(CGX_SPMM_ARGS);                   \
        case 9: return launch_spmm<GG, 1, 4, HINT_NC_NOALLOC, 6>(CGX_SPMM_ARGS);           \
        case 10: return launch_spmm<GG, 1, 4, HINT_CG, 6>(CGX_SPMM_ARGS);                  \
        case 11: return launch_spmm<GG, 1, 2, HINT_NC, 6>(CGX_SPMM_ARGS);                  \
        case 12: return launch_spmm<GG, 1, 8, HINT_NC, 3>(CGX_SPMM_ARGS);                  \
        case 13: return launch_spmm<GG, 1, 16, HINT_NC, 2>(CGX_SPMM_ARGS);                 \
        case 14: return launch_spmm<GG, 1, 4, HINT_NC, 4>(CGX_SPMM_ARGS);                  \
        case 15: return launch_spmm<GG, 1, 8, HINT_NC, 4>(CGX_SPMM_ARGS);                  \
        case 16: return launch_spmm<GG, 1, 8, HINT_NC, 1>(CGX_SPMM_ARGS);                  \
        case 17: return launch_spmm<GG / 2, 2, 4, HINT_NC, 4>(CGX_SPMM_ARGS);              \
        case 18: return launch_spmm<GG / 2, 2, 8, HINT_NC, 2>(CGX_SPMM_ARGS);              \
        case 19: return launch_spmm<GG / 4, 4, 2, HINT_NC, 4>(CGX_SPMM_ARGS);              \
        case 20: return launch_spmm<GG / 2, 2, 2, HINT_NC, 6>(CGX_SPMM_ARGS);              \
        case 21: return launch_spmm<GG / 2, 2, 4, HINT_NC, 3>(CGX_SPMM_ARGS);              \
        default: return launch_spmm<GG, 1, 8, HINT_NC, 4>(CGX_SPMM_ARGS);                  \
      }
#endif
    case 64:
      if (variant == 0 && beyond_l2) return launch_spmm<8, 2, 4, HINT_NC, 4>(CGX_SPMM_ARGS);
      CGX_VARIANTS(16)
    case 128:
      if (variant == 0 && beyond_l2) return launch_spmm<16, 2, 4, HINT_NC, 4>(CGX_SPMM_ARGS);
      CGX_VARIANTS(32)
    case 256:
#ifdef CGX_SPMM_TUNING
      if (variant == 17) return launch_spmm<16, 4, 2, HINT_NC, 4>(CGX_SPMM_ARGS);
#endif
      return launch_spmm<32, 2, 4, HINT_NC, 4>(CGX_SPMM_ARGS);
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

extern "C" int cgx_spmm_sparse_rows(const cgx_csr* m, int use_bwd_values, int32_t d, const float* X,
                                    const uint8_t* x_row_nonzero, float* Y, const float* ACC_IN, float* ACC_OUT,
                                    float acc_scale, void* workspace, size_t workspace_bytes, void* stream) {
  CGX_REQUIRE(x_row_nonzero != nullptr, CGX_ERR_ARG, "spmm_sparse_rows: NULL flags");
  return spmm_dispatch(m, use_bwd_values, d, X, Y, ACC_IN, ACC_OUT, acc_scale, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream), x_row_nonzero);
}

extern "C" int64_t cgx_spmm_set_l2_table_bytes(int64_t bytes) {
  const int64_t old = g_l2_table_bytes;
  g_l2_table_bytes = bytes < 0 ? (int64_t(96) << 20) : bytes;
  return old;
}

extern "C" int cgx_spmm_set_push_peers(void* const* peer_bases, int world) {
  CGX_REQUIRE(peer_bases != nullptr && world >= 1 && world <= 16, CGX_ERR_ARG, "spmm_set_push_peers: bad argument");
  char* h[16] = {nullptr};
  for (int p = 0; p < world; ++p) h[p] = static_cast<char*>(peer_bases[p]);
  CGX_CUDA(cudaMemcpyToSymbol(c_push_base, h, sizeof(h)));
  return CGX_OK;
}

extern "C" int cgx_spmm_push(const cgx_csr* m, int use_bwd_values, int32_t d, const float* X,
                             const uint8_t* x_row_nonzero, size_t stage_off, int rank, int world, int32_t rows_per,
                             void* workspace, size_t workspace_bytes, void* stream) {
  CGX_REQUIRE(m != nullptr && world >= 1 && world <= 16 && rank >= 0 && rank < world && rows_per > 0 &&
                  int64_t(rows_per) * world >= m->n_rows && stage_off % 16 == 0,
              CGX_ERR_ARG, "spmm_push: bad rank / world / rows_per");
  const SpmmPush ps{(unsigned long long)stage_off, rows_per, rank};
  return spmm_dispatch(m, use_bwd_values, d, X, nullptr, nullptr, nullptr, 1.0f, workspace, workspace_bytes,
                       static_cast<cudaStream_t>(stream), x_row_nonzero, ps);
}

extern "C" int cgx_row_flags(const float* X, int64_t n_rows, int32_t d, uint8_t* flags, void* stream) {
  CGX_REQUIRE(X && flags && n_rows >= 0 && d > 0 && d % 4 == 0, CGX_ERR_ARG, "row_flags: bad argument");
  return row_flags(X, n_rows, d, flags, static_cast<cudaStream_t>(stream));
}

extern "C" size_t cgx_propagate_workspace_bytes(const cgx_csr* by_user, const cgx_csr* by_item, int32_t d) {
  if (!by_user || !by_item) return 0;
  size_t long_ws = spmm_ws(by_user, d) > spmm_ws(by_item, d) ? spmm_ws(by_user, d) : spmm_ws(by_item, d);
  return 2 * (align_up(size_t(by_user->n_rows) * d * 4) + align_up(size_t(by_item->n_rows) * d * 4)) + long_ws +
         align_up(size_t(by_user->n_rows)) + align_up(size_t(by_item->n_rows)) + 256;
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
  // The incoming gradient has non-zero rows only where the batch touched the tables (<= batch users,
  // <= 2 * batch items): the products that gather g_u / g_i directly skip the zero rows (row flags, 1 byte per
  // row).  After one product the adjoint is dense.  CGX_SPARSE_BWD=0 switches this off.
  static const bool sparse = spmm_env("CGX_SPARSE_BWD", 1) != 0;
  uint8_t* nz_u = ws.take<uint8_t>(U);
  uint8_t* nz_i = ws.take<uint8_t>(I);
  CGX_REQUIRE(ws.ok, CGX_ERR_WORKSPACE, "propagate_bwd: workspace too small");
  // Work with the unscaled adjoints bu' = bu / s, bi' = bi / s (the recurrences are linear):
  //   Jacobi: (bu', bi') <- (g_u + C^T bi', g_i + A^T bu');  Gauss-Seidel: bi' = g_i + A^T bu'; bu' = g_u + C^T bi'
  // and fold s into the last product's epilogue.
  if (order == CGX_ORDER_JACOBI) {
    const float* bu = g_u;
    const float* bi = g_i;
    if (sparse) {
      CGX_TRY(row_flags(g_u, U, d, nz_u, stream));
      CGX_TRY(row_flags(g_i, I, d, nz_i, stream));
    }
    for (int k = 0; k < K; ++k) {
      const bool last = k == K - 1;
      float* nu = last ? d_e0_u : ub[k & 1];
      float* ni = last ? d_e0_i : ib[k & 1];
      const float sc = last ? s : 1.0f;
      const bool sp = sparse && k == 0;
      CGX_TRY(spmm_dispatch(by_user, 1, d, bi, nullptr, g_u, nu, sc, lw, lws, stream, sp ? nz_i : nullptr));  // g_u + C^T bi
      CGX_TRY(spmm_dispatch(by_item, 1, d, bu, nullptr, g_i, ni, sc, lw, lws, stream, sp ? nz_u : nullptr));  // g_i + A^T bu
      bu = nu;
      bi = ni;
    }
  } else {
    const float* bu = g_u;
    if (sparse) CGX_TRY(row_flags(g_u, U, d, nz_u, stream));
    for (int k = 0; k < K; ++k) {
      const bool last = k == K - 1;
      float* ni = ib[0];
      float* nu = last ? d_e0_u : ub[k & 1];
      CGX_TRY(spmm_dispatch(by_item, 1, d, bu, nullptr, g_i, ni, 1.0f, lw, lws, stream,
                            sparse && k == 0 ? nz_u : nullptr));                                      // bi'
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
