// Dense Adam over the two embedding tables in one launch, plus the device-side step counter that
// makes a whole training step replayable as a CUDA graph.
//
// Replaces (reference, /root/reference): torch.optim.Adam(model.parameters(), lr) and its
// opt.step() at lightgcn_cu.py:587,652 / Version-2/lighgcn_cu_pop.py:793,863 (defaults: betas
// (0.9, 0.999), eps 1e-8, no weight decay, no amsgrad).  Same update rule as torch:
//     m = b1 m + (1-b1) g;  v = b2 v + (1-b2) g^2
//     p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// Streaming kernel: 16 B per lane loads/stores of p, g, m, v (7 x 4 B per parameter of HBM traffic).
#include "common.cuh"

namespace cgx {

struct AdamSeg {
  float4* p;
  const float4* g;
  float4* m;
  float4* v;
  int64_t n4;
};

__global__ void k_tick(unsigned long long* counter) { *counter += 1ull; }

__global__ void __launch_bounds__(256) k_adam(AdamSeg s0, AdamSeg s1, float lr, float b1, float b2, float eps,
                                              const unsigned long long* __restrict__ step_dev, int64_t step_host) {
  const float t = float(step_dev ? (long long)(*step_dev) + step_host : step_host);
  const float bc1 = 1.0f - powf(b1, t);
  const float bc2_sqrt = sqrtf(1.0f - powf(b2, t));
  const float step_size = lr / bc1;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int seg = 0; seg < 2; ++seg) {
    const AdamSeg& s = seg == 0 ? s0 : s1;
    for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < s.n4; i += stride) {
      const float4 g = __ldg(s.g + i);
      float4 p = s.p[i], m = s.m[i], v = s.v[i];
#define CGX_ADAM1(c)                                 \
  m.c = b1 * m.c + (1.0f - b1) * g.c;                \
  v.c = b2 * v.c + (1.0f - b2) * g.c * g.c;          \
  p.c -= step_size * (m.c / (sqrtf(v.c) / bc2_sqrt + eps));
      CGX_ADAM1(x) CGX_ADAM1(y) CGX_ADAM1(z) CGX_ADAM1(w)
#undef CGX_ADAM1
      s.p[i] = p;
      s.m[i] = m;
      s.v[i] = v;
    }
  }
}

}  // namespace cgx

using namespace cgx;

extern "C" int cgx_tick(uint64_t* counter, void* stream_) {
  CGX_REQUIRE(counter != nullptr, CGX_ERR_ARG, "tick: NULL counter");
  k_tick<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<unsigned long long*>(counter));
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}

extern "C" int cgx_adam_step(float* p0, const float* g0, float* m0, float* v0, int64_t n0, float* p1, const float* g1,
                             float* m1, float* v1, int64_t n1, float lr, float beta1, float beta2, float eps,
                             const uint64_t* step_dev, int64_t step_host, void* stream_) {
  CGX_REQUIRE(p0 && g0 && m0 && v0 && n0 > 0 && n0 % 4 == 0 && n1 >= 0 && n1 % 4 == 0, CGX_ERR_ARG,
              "adam_step: bad argument");
  CGX_REQUIRE(n1 == 0 || (p1 && g1 && m1 && v1), CGX_ERR_ARG, "adam_step: NULL pointer");
  AdamSeg s0{reinterpret_cast<float4*>(p0), reinterpret_cast<const float4*>(g0), reinterpret_cast<float4*>(m0),
             reinterpret_cast<float4*>(v0), n0 / 4};
  AdamSeg s1{reinterpret_cast<float4*>(p1), reinterpret_cast<const float4*>(g1), reinterpret_cast<float4*>(m1),
             reinterpret_cast<float4*>(v1), n1 / 4};
  const int64_t n4 = (n0 + n1) / 4;
  int64_t blocks = ceil_div(n4, 256);
  if (blocks > 148 * 16) blocks = 148 * 16;   // grid-stride: 16 CTAs of 256 threads per SM
  k_adam<<<(unsigned)blocks, 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      s0, s1, lr, beta1, beta2, eps, reinterpret_cast<const unsigned long long*>(step_dev), step_host);
  CGX_LAUNCH_CHECK();
  return CGX_OK;
}
