// Tensor-core (tcgen05) score path for full-rank evaluation -- placeholder until the UMMA kernel
// lands; the fp32 path in eval.cu is complete and exact.
#include "common.cuh"

namespace cgx {

size_t eval_topk_tc_workspace(int64_t, int32_t, int32_t, int32_t, int) { return 256; }

int eval_topk_tc(const int64_t*, int64_t, const float*, const float*, int32_t, int32_t, const int64_t*,
                 const int32_t*, int32_t, int precision, int32_t*, float*, void*, size_t, cudaStream_t) {
  set_error("eval_topk: score precision %d is not implemented yet (use CGX_SCORE_FP32)", precision);
  return CGX_ERR_UNSUPPORTED;
}

}  // namespace cgx
