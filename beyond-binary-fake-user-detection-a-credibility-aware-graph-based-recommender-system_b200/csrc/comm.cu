// Item-table exchange of the user-sharded propagation over NVLink peer memory.
//
// The reference is single-device (no collective to replace); SURVEY.md section 8e defines the exchange: every
// rank holds a PARTIAL item table (the product over its own users) that must be summed over ranks once
// per layer.  NCCL's all-reduce is latency-bound at these sizes (10-20 MB: ~56 us at 2 GPUs, measured),
// so the sum is done by one kernel over peer-mapped buffers instead:
//
//   * every rank owns one cudaMalloc'ed communication buffer, exported with cudaIpcGetMemHandle and
//     mapped by all peers (cudaIpcOpenMemHandle): two `in` regions, two `out` regions (parity of the
//     collective's epoch), and a flag page;
//   * the SpMM writes its partial straight into in[parity] (no staging copy);
//   * k_p2p_allreduce, two-shot and pull-based:
//       A. signal "my partial is complete" into every peer's flag page (st.release.sys), wait for all
//          peers' signals (ld.acquire.sys);
//       1. rank r sums slice r of all partials IN RANK ORDER (peer loads over NVLink; deterministic,
//          identical bits on every rank) and stores the reduced slice into every peer's out[parity];
//       B. the last CTA to finish signals "my slice is delivered" to every peer and waits for theirs, so
//          the kernel completes only when out[parity] is whole.
//     Barrier B also means nobody still reads a rank's `in` region when its kernel completes, and barrier A
//     means nobody writes a rank's `out` region before that rank's earlier kernels (the consumers of the
//     previous result) have finished: regions can be re-used immediately.  Two regions are kept only so that
//     the result of the previous exchange stays readable (the Jacobi order needs it).
//   * small worlds (<= CGX_OPT_P2P_ONESHOT_MAX ranks, default 2) use the ONE-SHOT form instead: after barrier A every
//     rank pulls ALL partials and sums the whole table locally, in rank order (same bits on every rank), into its
//     own out[parity]; no remote store, no barrier B.  It moves (R-1) x the table per rank instead of
//     2 (R-1)/R x, but saves a system-scope fence, a barrier and an NVLink round trip -- at 10 MB and 2 ranks the
//     exchange is latency-bound, not wire-bound.  A peer's in[parity] is next overwritten two exchanges later,
//     i.e. after that peer has passed barrier A of the exchange in between, which this rank only signals once
//     this kernel has completed.
//   * above 2 ranks the propagation uses the PUSHED form (k_p2p_reduce_pushed, below): the SpMM epilogue has already
//     stored every partial row into its owner's staging slot, so the reduce needs local loads only.
//   * large tables whose item rows are short (C5 shards) take the NVLS form (k_nvls_allreduce): slice r is summed
//     INSIDE the NVSwitch by multimem.ld_reduce on a multicast mapping of the `in` regions and replicated into every
//     `out` region by multimem.st -- about half the NVLink bytes of the pull form, level with NCCL at 2.56 GB.
//   * the loss gradient of the sharded step travels as one small all-gather (k_p2p_allgather: <= 2 * batch item rows
//     per rank, posted stores into every rank's gather region + one barrier).
//   * every cross-GPU wait is bounded (wait_flag, CGX_OPT_P2P_TIMEOUT_MS): a dead peer turns into an error word
//     the host can read (cgx_comm_status), not into eight GPUs spinning inside a replayed graph.
//   * the epoch is a device-side counter (cgx_tick before every exchange) that advances identically on all
//     ranks (the propagation schedule is deterministic), so a whole step can be replayed as a CUDA graph.
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace cgx {

constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_THREADS = 256;

struct P2PPeers {
  char* base[P2P_MAX_RANKS];   // peer-mapped communication buffers, index = rank
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_peer(const float4* p) {   // written by another GPU: never from a stale L1 line
  float4 r;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

__device__ __forceinline__ unsigned long long p2p_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// Bounded wait for `*p >= epoch` (a flag another GPU writes).  A peer that died or lost step with this rank would
// otherwise hang every GPU of the job inside a replayed CUDA graph: after timeout_ns the waiter records
// (barrier << 8 | peer + 1) in the flag page's error word (first failure wins; read by cgx_comm_status) and goes
// on -- the exchange's result is then unusable, but every kernel terminates and the host can report it.
__device__ __forceinline__ void wait_flag(const uint32_t* p, uint32_t epoch, unsigned long long timeout_ns,
                                          uint32_t* err, uint32_t code) {
  unsigned long long t0 = 0;
  unsigned spins = 0;
  while (ld_acquire_sys(p) < epoch) {
    __nanosleep(64);
    if ((++spins & 255u) == 0) {                 // look at the clock every ~16 us of waiting
      const unsigned long long now = p2p_now();
      if (t0 == 0) t0 = now;
      else if (now - t0 > timeout_ns) {
        atomicCAS(err, 0u, code);
        return;
      }
    }
  }
}

// flag page layout (uint32): [0 .. R) barrier A slots, [R .. 2R) barrier B slots, [2R] CTA arrival counter,
// [2R + 1] error word (0 = no barrier has timed out)
__global__ void __launch_bounds__(P2P_THREADS) k_p2p_allreduce(P2PPeers peers, int rank, int world, size_t in_off,
                                                               size_t out_off, size_t flag_off, int64_t n4,
                                                               const unsigned long long* __restrict__ epoch_dev,
                                                               int one_shot, unsigned long long timeout_ns,
                                                               unsigned long long* dbg) {
  // dbg (CGX_OPT_P2P_TIMING): ns spent in [0] barrier A, [1] reduce + delivery, [2] barrier B; [3] launches; [4] scratch
  unsigned long long t0 = 0;
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) t0 = p2p_now();
  const uint32_t epoch = uint32_t(*epoch_dev);   // device-side counter: the launch is replayable in a CUDA graph
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(peers.base[rank] + flag_off);
  // ---- barrier A: all partials complete ----
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + rank, epoch);
  }
  if (threadIdx.x < world)
    wait_flag(my_flags + threadIdx.x, epoch, timeout_ns, my_flags + 2 * world + 1, (1u << 8) | (threadIdx.x + 1));
  __syncthreads();
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long t1 = p2p_now();
    dbg[0] += t1 - t0;
    dbg[3] += 1;
    dbg[4] = t1;
  }
  // ---- reduce-scatter + all-gather of my slice (one-shot: the whole table, stored locally only) ----
  const int64_t per = one_shot ? n4 : (n4 + world - 1) / world;
  const int64_t lo = one_shot ? 0 : int64_t(rank) * per;
  const int64_t hi = lo + per < n4 ? lo + per : n4;
  // P2P_UNROLL x world peer loads are issued before the first add: NVLink round trips (~2 us) overlap
  constexpr int P2P_UNROLL = 4;
  const int64_t stride = int64_t(gridDim.x) * P2P_THREADS;
  for (int64_t i0 = lo + int64_t(blockIdx.x) * P2P_THREADS + threadIdx.x; i0 < hi; i0 += stride * P2P_UNROLL) {
    float4 acc[P2P_UNROLL];
#pragma unroll
    for (int t = 0; t < P2P_UNROLL; ++t) {
      const int64_t i = i0 + t * stride;
      acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < hi) acc[t] = ld_peer(reinterpret_cast<const float4*>(peers.base[0] + in_off) + i);
    }
    for (int p = 1; p < world; ++p) {
      float4 v[P2P_UNROLL];
#pragma unroll
      for (int t = 0; t < P2P_UNROLL; ++t) {
        const int64_t i = i0 + t * stride;
        v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < hi) v[t] = ld_peer(reinterpret_cast<const float4*>(peers.base[p] + in_off) + i);
      }
#pragma unroll
      for (int t = 0; t < P2P_UNROLL; ++t) {
        acc[t].x += v[t].x; acc[t].y += v[t].y; acc[t].z += v[t].z; acc[t].w += v[t].w;
      }
    }
#pragma unroll
    for (int t = 0; t < P2P_UNROLL; ++t) {
      const int64_t i = i0 + t * stride;
      if (i < hi) {
        if (one_shot) {
          reinterpret_cast<float4*>(peers.base[rank] + out_off)[i] = acc[t];
        } else {
          for (int p = 0; p < world; ++p) reinterpret_cast<float4*>(peers.base[p] + out_off)[i] = acc[t];
        }
      }
    }
  }
  if (one_shot) {
    if (dbg && blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) dbg[1] += p2p_now() - dbg[4];   // one CTA's view
    return;
  }
  // ---- barrier B: every slice delivered ----
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(my_flags + 2 * world, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  unsigned long long t2 = 0;
  if (dbg && threadIdx.x == 0) {
    t2 = p2p_now();
    dbg[1] += t2 - dbg[4];
  }
  if (threadIdx.x == 0) my_flags[2 * world] = 0;   // self-resetting
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + world + rank, epoch);
    wait_flag(my_flags + world + threadIdx.x, epoch, timeout_ns, my_flags + 2 * world + 1,
              (2u << 8) | (threadIdx.x + 1));
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[2] += p2p_now() - t2;
}

// Pushed form (cgx_spmm_push + cgx_comm_allreduce_pushed): the SpMM already stored every partial row into the
// staging area of the row's OWNER, slot = source rank, so after barrier A the owner sums its rows_per rows over the
// world slots with LOCAL loads (rank order, same bits as the pull form), stores the reduced rows into every rank's
// out region, and barrier B closes the exchange.  No remote load is left on the path.
__global__ void __launch_bounds__(P2P_THREADS) k_p2p_reduce_pushed(P2PPeers peers, int rank, int world,
                                                                   size_t stage_off, size_t out_off, size_t flag_off,
                                                                   int64_t n_rows, int32_t row4, int32_t rows_per,
                                                                   const unsigned long long* __restrict__ epoch_dev,
                                                                   unsigned long long timeout_ns,
                                                                   unsigned long long* dbg) {
  unsigned long long t0 = 0;
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) t0 = p2p_now();
  const uint32_t epoch = uint32_t(*epoch_dev);
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(peers.base[rank] + flag_off);
  // ---- barrier A: every peer's SpMM (the kernel before its exchange kernel) has completed its pushes ----
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + rank, epoch);
  }
  if (threadIdx.x < world)
    wait_flag(my_flags + threadIdx.x, epoch, timeout_ns, my_flags + 2 * world + 1, (1u << 8) | (threadIdx.x + 1));
  __syncthreads();
  if (dbg && blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long t1 = p2p_now();
    dbg[0] += t1 - t0;
    dbg[3] += 1;
    dbg[4] = t1;
  }
  const int64_t lo_row = int64_t(rank) * rows_per;
  const int64_t hi_row = lo_row + rows_per < n_rows ? lo_row + rows_per : n_rows;
  const int64_t n4 = hi_row > lo_row ? (hi_row - lo_row) * row4 : 0;
  const int64_t slot4 = int64_t(rows_per) * row4;
  const float4* stage = reinterpret_cast<const float4*>(peers.base[rank] + stage_off);
  const int64_t stride = int64_t(gridDim.x) * P2P_THREADS;
  // PU float4 per thread and pass: all PU x world staged values are requested before the first add, and the PU x world
  // posted NVLink stores of a pass leave back to back
  constexpr int PU = 4;
  for (int64_t i0 = int64_t(blockIdx.x) * P2P_THREADS + threadIdx.x; i0 < n4; i0 += stride * PU) {
    float4 acc[PU];
#pragma unroll
    for (int t = 0; t < PU; ++t) {
      const int64_t i = i0 + t * stride;
      acc[t] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (i < n4) acc[t] = ld_peer(stage + i);             // written by rank 0's SpMM (maybe a remote GPU): no L1
    }
    for (int p = 1; p < world; ++p) {
      float4 v[PU];
#pragma unroll
      for (int t = 0; t < PU; ++t) {
        const int64_t i = i0 + t * stride;
        v[t] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (i < n4) v[t] = ld_peer(stage + int64_t(p) * slot4 + i);
      }
#pragma unroll
      for (int t = 0; t < PU; ++t) {
        acc[t].x += v[t].x; acc[t].y += v[t].y; acc[t].z += v[t].z; acc[t].w += v[t].w;
      }
    }
    for (int p = 0; p < world; ++p) {
      float4* dst = reinterpret_cast<float4*>(peers.base[p] + out_off) + lo_row * row4;
#pragma unroll
      for (int t = 0; t < PU; ++t) {
        const int64_t i = i0 + t * stride;
        if (i < n4) dst[i] = acc[t];
      }
    }
  }
  // ---- barrier B: every slice delivered ----
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(my_flags + 2 * world, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  unsigned long long t2 = 0;
  if (dbg && threadIdx.x == 0) {
    t2 = p2p_now();
    dbg[1] += t2 - dbg[4];
  }
  if (threadIdx.x == 0) my_flags[2 * world] = 0;   // self-resetting
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + world + rank, epoch);
    wait_flag(my_flags + world + threadIdx.x, epoch, timeout_ns, my_flags + 2 * world + 1,
              (2u << 8) | (threadIdx.x + 1));
  }
  __syncthreads();
  if (dbg && threadIdx.x == 0) dbg[2] += p2p_now() - t2;
}

// NVLS form of the exchange: when the communication buffers are bound to one NVSwitch multicast object (mc_base;
// torch symmetric memory sets that up), rank r reduces slice r with multimem.ld_reduce -- ONE load whose value is the
// sum over all ranks' `in` regions, added inside the switch -- and delivers it with multimem.st -- ONE store the switch
// replicates into every rank's `out` region.  Per exchange and direction a rank then moves ~S bytes over NVLink
// (its partial out, the reduced table in) instead of 2 (R-1)/R S with peer loads + peer stores.  The order of the
// in-switch sum is the switch's, not rank order: every rank receives the same bits, but they can differ in the last
// place from the pull / pushed forms (tests compare this form by tolerance).
__device__ __forceinline__ float4 multimem_ld_reduce(const float4* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ void multimem_st(float4* p, const float4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(P2P_THREADS) k_nvls_allreduce(P2PPeers peers, char* mc_base, int rank, int world,
                                                                size_t in_off, size_t out_off, size_t flag_off,
                                                                int64_t n4,
                                                                const unsigned long long* __restrict__ epoch_dev,
                                                                unsigned long long timeout_ns) {
  const uint32_t epoch = uint32_t(*epoch_dev);
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(peers.base[rank] + flag_off);
  // ---- barrier A: all partials complete and visible ----
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + rank, epoch);
  }
  if (threadIdx.x < world)
    wait_flag(my_flags + threadIdx.x, epoch, timeout_ns, my_flags + 2 * world + 1, (1u << 8) | (threadIdx.x + 1));
  __syncthreads();
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = int64_t(rank) * per;
  const int64_t hi = lo + per < n4 ? lo + per : n4;
  const float4* src = reinterpret_cast<const float4*>(mc_base + in_off);
  float4* dst = reinterpret_cast<float4*>(mc_base + out_off);
  constexpr int NU = 4;
  const int64_t stride = int64_t(gridDim.x) * P2P_THREADS;
  for (int64_t i0 = lo + int64_t(blockIdx.x) * P2P_THREADS + threadIdx.x; i0 < hi; i0 += stride * NU) {
    float4 v[NU];
#pragma unroll
    for (int t = 0; t < NU; ++t) {
      const int64_t i = i0 + t * stride;
      if (i < hi) v[t] = multimem_ld_reduce(src + i);
    }
#pragma unroll
    for (int t = 0; t < NU; ++t) {
      const int64_t i = i0 + t * stride;
      if (i < hi) multimem_st(dst + i, v[t]);
    }
  }
  // ---- barrier B: every slice delivered ----
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(my_flags + 2 * world, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) my_flags[2 * world] = 0;   // self-resetting
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + world + rank, epoch);
    wait_flag(my_flags + world + threadIdx.x, epoch, timeout_ns, my_flags + 2 * world + 1,
              (2u << 8) | (threadIdx.x + 1));
  }
  __syncthreads();
}

// All-gather of one small block per rank (the compact loss gradient of the user-sharded step: <= 2 * batch item rows
// per rank instead of two dense item tables).  Every rank stores its block into slot `rank` of the gather region of
// EVERY rank (posted NVLink stores), the last CTA to finish signals the peers (the barrier-A flag slots: "rank p has
// delivered epoch e") and waits for theirs, so the kernel completes only when all `world` blocks are local.
__global__ void __launch_bounds__(P2P_THREADS) k_p2p_allgather(P2PPeers peers, int rank, int world,
                                                               const float4* __restrict__ src, size_t dst_off,
                                                               size_t flag_off, int64_t n4,
                                                               const unsigned long long* __restrict__ epoch_dev,
                                                               unsigned long long timeout_ns) {
  const uint32_t epoch = uint32_t(*epoch_dev);
  uint32_t* my_flags = reinterpret_cast<uint32_t*>(peers.base[rank] + flag_off);
  const int64_t stride = int64_t(gridDim.x) * P2P_THREADS;
  for (int64_t i = int64_t(blockIdx.x) * P2P_THREADS + threadIdx.x; i < n4; i += stride) {
    const float4 v = src[i];
    for (int p = 0; p < world; ++p)
      reinterpret_cast<float4*>(peers.base[p] + dst_off)[int64_t(rank) * n4 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  __shared__ bool last;
  if (threadIdx.x == 0) last = atomicAdd(my_flags + 2 * world, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  if (threadIdx.x == 0) my_flags[2 * world] = 0;   // self-resetting
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint32_t*>(peers.base[threadIdx.x] + flag_off) + rank, epoch);
    wait_flag(my_flags + threadIdx.x, epoch, timeout_ns, my_flags + 2 * world + 1, (3u << 8) | (threadIdx.x + 1));
  }
  __syncthreads();
}

}  // namespace cgx

using namespace cgx;

extern "C" int cgx_comm_alloc(size_t bytes, void** base_out) {
  CGX_REQUIRE(base_out != nullptr && bytes > 0, CGX_ERR_ARG, "comm_alloc: bad argument");
  CGX_CUDA(cudaMalloc(base_out, bytes));
  CGX_CUDA(cudaMemset(*base_out, 0, bytes));
  CGX_CUDA(cudaDeviceSynchronize());
  return CGX_OK;
}

extern "C" int cgx_comm_free(void* base) {
  if (base) CGX_CUDA(cudaFree(base));
  return CGX_OK;
}

extern "C" int cgx_comm_ipc_handle(void* base, void* handle_out_64) {
  CGX_REQUIRE(base && handle_out_64, CGX_ERR_ARG, "comm_ipc_handle: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CGX_CUDA(cudaIpcGetMemHandle(static_cast<cudaIpcMemHandle_t*>(handle_out_64), base));
  return CGX_OK;
}

extern "C" int cgx_comm_ipc_open(const void* handle_64, void** peer_base_out) {
  CGX_REQUIRE(handle_64 && peer_base_out, CGX_ERR_ARG, "comm_ipc_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle_64, sizeof(h));
  CGX_CUDA(cudaIpcOpenMemHandle(peer_base_out, h, cudaIpcMemLazyEnablePeerAccess));
  return CGX_OK;
}

extern "C" int cgx_comm_ipc_close(void* peer_base) {
  if (peer_base) CGX_CUDA(cudaIpcCloseMemHandle(peer_base));
  return CGX_OK;
}

static unsigned long long* g_p2p_dbg = nullptr;

/* CGX_OPT_P2P_TIMING diagnostics: ns accumulated in {barrier A, reduce + delivery, barrier B} and the number of
 * exchanges since the last call (synchronises the device; zeros when timing is off). */
extern "C" int cgx_comm_timing(uint64_t* out4) {
  CGX_REQUIRE(out4 != nullptr, CGX_ERR_ARG, "comm_timing: NULL pointer");
  for (int i = 0; i < 4; ++i) out4[i] = 0;
  if (g_p2p_dbg == nullptr) return CGX_OK;
  CGX_CUDA(cudaDeviceSynchronize());
  CGX_CUDA(cudaMemcpy(out4, g_p2p_dbg, 32, cudaMemcpyDeviceToHost));
  CGX_CUDA(cudaMemset(g_p2p_dbg, 0, 64));
  return CGX_OK;
}

static unsigned long long p2p_timeout_ns() { return (unsigned long long)option(CGX_OPT_P2P_TIMEOUT_MS) * 1000000ull; }

static int p2p_dbg_buffer(unsigned long long** out) {
  static unsigned long long* dbg = nullptr;
  if (dbg == nullptr && option(CGX_OPT_P2P_TIMING) != 0) {
    CGX_CUDA(cudaMalloc(&dbg, 64));
    CGX_CUDA(cudaMemset(dbg, 0, 64));
    g_p2p_dbg = dbg;
  }
  *out = dbg;
  return CGX_OK;
}

extern "C" int cgx_comm_allreduce_pushed(int rank, int world, void* const* peer_bases, size_t stage_off, size_t out_off,
                                         size_t flag_off, int64_t n_rows, int32_t d, int32_t rows_per,
                                         const uint64_t* epoch_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world && peer_bases, CGX_ERR_ARG,
              "comm_allreduce_pushed: bad rank/world");
  CGX_REQUIRE(n_rows > 0 && d > 0 && d % 4 == 0 && rows_per > 0 && int64_t(rows_per) * world >= n_rows &&
                  stage_off % 16 == 0 && out_off % 16 == 0 && flag_off % 16 == 0 && epoch_dev != nullptr,
              CGX_ERR_ARG, "comm_allreduce_pushed: bad sizes/offsets");
  P2PPeers peers;
  for (int p = 0; p < world; ++p) {
    CGX_REQUIRE(peer_bases[p] != nullptr, CGX_ERR_ARG, "comm_allreduce_pushed: NULL peer buffer");
    peers.base[p] = static_cast<char*>(peer_bases[p]);
  }
  unsigned long long* dbg = nullptr;
  CGX_TRY(p2p_dbg_buffer(&dbg));
  int64_t blocks = ceil_div(int64_t(rows_per) * (d / 4), P2P_THREADS);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  k_p2p_reduce_pushed<<<(unsigned)blocks, P2P_THREADS, 0, stream>>>(
      peers, rank, world, stage_off, out_off, flag_off, n_rows, d / 4, rows_per,
      reinterpret_cast<const unsigned long long*>(epoch_dev), p2p_timeout_ns(), dbg);
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}

extern "C" int cgx_comm_allreduce_nvls(int rank, int world, void* const* peer_bases, void* mc_base, size_t in_off,
                                       size_t out_off, size_t flag_off, int64_t n_floats, const uint64_t* epoch_dev,
                                       void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world && peer_bases && mc_base, CGX_ERR_ARG,
              "comm_allreduce_nvls: bad rank/world or no multicast mapping");
  CGX_REQUIRE(n_floats > 0 && n_floats % 4 == 0 && in_off % 16 == 0 && out_off % 16 == 0 && flag_off % 16 == 0 &&
                  epoch_dev != nullptr,
              CGX_ERR_ARG, "comm_allreduce_nvls: bad sizes/offsets");
  P2PPeers peers;
  for (int p = 0; p < world; ++p) {
    CGX_REQUIRE(peer_bases[p] != nullptr, CGX_ERR_ARG, "comm_allreduce_nvls: NULL peer buffer");
    peers.base[p] = static_cast<char*>(peer_bases[p]);
  }
  const int64_t n4 = n_floats / 4;
  int64_t blocks = ceil_div(ceil_div(n4, world), P2P_THREADS * 4);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  k_nvls_allreduce<<<(unsigned)blocks, P2P_THREADS, 0, stream>>>(
      peers, static_cast<char*>(mc_base), rank, world, in_off, out_off, flag_off, n4,
      reinterpret_cast<const unsigned long long*>(epoch_dev), p2p_timeout_ns());
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}

extern "C" int cgx_comm_allgather(int rank, int world, void* const* peer_bases, const void* src, size_t dst_off,
                                  size_t flag_off, int64_t bytes_per_rank, const uint64_t* epoch_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world && peer_bases && src, CGX_ERR_ARG,
              "comm_allgather: bad rank/world");
  CGX_REQUIRE(bytes_per_rank > 0 && bytes_per_rank % 16 == 0 && dst_off % 16 == 0 && flag_off % 16 == 0 &&
                  (reinterpret_cast<uintptr_t>(src) & 15) == 0 && epoch_dev != nullptr,
              CGX_ERR_ARG, "comm_allgather: bad sizes/offsets (16-byte granularity)");
  P2PPeers peers;
  for (int p = 0; p < world; ++p) {
    CGX_REQUIRE(peer_bases[p] != nullptr, CGX_ERR_ARG, "comm_allgather: NULL peer buffer");
    peers.base[p] = static_cast<char*>(peer_bases[p]);
  }
  const int64_t n4 = bytes_per_rank / 16;
  int64_t blocks = ceil_div(n4, P2P_THREADS * 2);
  if (blocks > 148 * 2) blocks = 148 * 2;
  if (blocks < 1) blocks = 1;
  k_p2p_allgather<<<(unsigned)blocks, P2P_THREADS, 0, stream>>>(
      peers, rank, world, static_cast<const float4*>(src), dst_off, flag_off, n4,
      reinterpret_cast<const unsigned long long*>(epoch_dev), p2p_timeout_ns());
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}

extern "C" int cgx_comm_status(const void* base, size_t flag_off, int world, uint32_t* error_out) {
  CGX_REQUIRE(base && error_out && world >= 1 && world <= P2P_MAX_RANKS && flag_off % 16 == 0, CGX_ERR_ARG,
              "comm_status: bad argument");
  CGX_CUDA(cudaDeviceSynchronize());
  CGX_CUDA(cudaMemcpy(error_out, static_cast<const char*>(base) + flag_off + 4 * (2 * size_t(world) + 1), 4,
                      cudaMemcpyDeviceToHost));
  return CGX_OK;
}

extern "C" int cgx_comm_allreduce(int rank, int world, void* const* peer_bases, size_t in_off, size_t out_off,
                                  size_t flag_off, int64_t n_floats, const uint64_t* epoch_dev, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CGX_REQUIRE(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world && peer_bases, CGX_ERR_ARG,
              "comm_allreduce: bad rank/world");
  CGX_REQUIRE(n_floats > 0 && n_floats % 4 == 0 && in_off % 16 == 0 && out_off % 16 == 0 && flag_off % 16 == 0 &&
                  epoch_dev != nullptr,
              CGX_ERR_ARG, "comm_allreduce: bad sizes/offsets");
  P2PPeers peers;
  for (int p = 0; p < world; ++p) {
    CGX_REQUIRE(peer_bases[p] != nullptr, CGX_ERR_ARG, "comm_allreduce: NULL peer buffer");
    peers.base[p] = static_cast<char*>(peer_bases[p]);
  }
  const int64_t n4 = n_floats / 4;
  const int one_shot = world <= option(CGX_OPT_P2P_ONESHOT_MAX) ? 1 : 0;
  unsigned long long* dbg = nullptr;
  CGX_TRY(p2p_dbg_buffer(&dbg));
  const int64_t per = one_shot ? n4 : ceil_div(n4, world);
  int64_t blocks = ceil_div(per, P2P_THREADS * 4);
  if (blocks > 148 * 4) blocks = 148 * 4;
  if (blocks < 1) blocks = 1;
  k_p2p_allreduce<<<(unsigned)blocks, P2P_THREADS, 0, stream>>>(peers, rank, world, in_off, out_off, flag_off, n4,
                                                             reinterpret_cast<const unsigned long long*>(epoch_dev),
                                                             one_shot, p2p_timeout_ns(), dbg);
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}
