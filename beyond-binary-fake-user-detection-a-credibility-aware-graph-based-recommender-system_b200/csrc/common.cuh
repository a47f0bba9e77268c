// Shared host/device helpers for libcredgcn.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "credgcn.h"

namespace cgx {

void set_error(const char* fmt, ...);
int64_t option(int which);   // current value of a cgx_option (cgx_set_option)

#define CGX_CUDA(expr)                                                                   \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess) {                                                             \
      ::cgx::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return CGX_ERR_CUDA;                                                               \
    }                                                                                    \
  } while (0)

// every kernel launch of the library goes through this: error check + process-wide launch tally
// (read by cgx_launch_count(); bench.py reports the delta over its timed region as gpu_launches)
void count_launch();
#define CGX_LAUNCH_CHECK()          \
  do {                              \
    ::cgx::count_launch();          \
    CGX_CUDA(cudaGetLastError());   \
  } while (0)

#define CGX_REQUIRE(cond, code, ...)   \
  do {                                 \
    if (!(cond)) {                     \
      ::cgx::set_error(__VA_ARGS__);   \
      return (code);                   \
    }                                  \
  } while (0)

#define CGX_TRY(expr)          \
  do {                         \
    int _s = (expr);           \
    if (_s != CGX_OK) return _s; \
  } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t size;
  size_t used = 0;
  bool ok = true;
  Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align_up(count * sizeof(T));
    if (base == nullptr || used + bytes > size) {
      ok = false;
      used += bytes;
      return nullptr;
    }
    T* p = reinterpret_cast<T*>(base + used);
    used += bytes;
    return p;
  }
};

static inline int bits_for(int64_t n) {  // bits needed to hold values 0 .. n-1 (at least 1)
  int b = 1;
  while ((int64_t(1) << b) < n) ++b;
  return b;
}

// ---- scan / sort primitives (scan_sort.cu) ----
size_t scan_temp_bytes(int64_t n);
// out[k] = sum_{j<k} in[j]; if total_out != nullptr, *total_out = sum of all (device pointer).
// in == out allowed.
int exclusive_scan_u32(const uint32_t* in, uint32_t* out, int64_t n, uint32_t* total_out, void* temp,
                       size_t temp_bytes, cudaStream_t stream);

size_t radix_sort_temp_bytes(int64_t n);
// LSD radix sort of the low `bits` bits, 8 bits per pass, stable.  keys/alt are ping-pong buffers;
// *sorted points at whichever holds the result.
int radix_sort_u64(uint64_t* keys, uint64_t* alt, int64_t n, int bits, void* temp, size_t temp_bytes,
                   cudaStream_t stream, uint64_t** sorted);

}  // namespace cgx
