/* Plain-C caller of libcredgcn.so: what a non-Python host (cgo / JNI / Rust FFI ...) would link against.
 * Only host-side entry points are called, so it runs without a GPU (tests/test_cabi_cpu.py builds and runs it). */
#include <stdio.h>
#include <string.h>

#include "credgcn.h"

int main(void) {
  int failures = 0;
  if (cgx_emb_dim_supported(64) != 1 || cgx_emb_dim_supported(48) != 0) { puts("emb_dim table"); ++failures; }
  /* the tensor-core evaluation covers BF16X3 up to d = 128 and BF16 for every supported width */
  if (cgx_eval_topk_uses_tensor_cores(64, 20, CGX_SCORE_BF16X3) != 1) { puts("tc d=64"); ++failures; }
  if (cgx_eval_topk_uses_tensor_cores(128, 20, CGX_SCORE_BF16X3) != 1) { puts("tc d=128"); ++failures; }
  if (cgx_eval_topk_uses_tensor_cores(256, 20, CGX_SCORE_BF16X3) != 0) { puts("tc d=256 x3"); ++failures; }
  if (cgx_eval_topk_uses_tensor_cores(256, 20, CGX_SCORE_BF16) != 1) { puts("tc d=256 bf16"); ++failures; }
  if (cgx_eval_topk_uses_tensor_cores(64, 20, CGX_SCORE_FP32) != 0) { puts("tc fp32"); ++failures; }
  if (cgx_eval_topk_uses_tensor_cores(64, 60, CGX_SCORE_BF16X3) != 0) { puts("tc k=60"); ++failures; }
  /* workspace queries are pure arithmetic and grow with the problem */
  if (cgx_eval_topk_workspace_bytes(1000, 5000, 64, 20, CGX_SCORE_BF16X3) <=
      cgx_eval_topk_workspace_bytes(1000, 5000, 64, 20, CGX_SCORE_FP32)) { puts("eval ws"); ++failures; }
  if (cgx_bpr_workspace_bytes(4096, 31668, 38048) <= cgx_bpr_workspace_bytes(64, 31668, 38048)) { puts("bpr ws"); ++failures; }
  if (cgx_eval_metrics_workspace_bytes(100000, 2) <= cgx_eval_metrics_workspace_bytes(100, 2)) { puts("metrics ws"); ++failures; }
  /* argument errors come back as codes + a message, never as a crash */
  if (cgx_tick(NULL, NULL) == CGX_OK) { puts("tick(NULL) accepted"); ++failures; }
  if (strstr(cgx_last_error(), "tick") == NULL) { printf("last_error: %s\n", cgx_last_error()); ++failures; }
  if (cgx_eval_coverage(NULL, 10, 1, NULL, NULL) == CGX_OK) { puts("coverage(NULL) accepted"); ++failures; }
  /* the geometry threshold is a plain setter / getter */
  {
    const int64_t old = cgx_spmm_set_l2_table_bytes(123);
    if (cgx_spmm_set_l2_table_bytes(-1) != 123 || cgx_spmm_set_l2_table_bytes(old) != ((int64_t)96 << 20)) {
      puts("l2 threshold"); ++failures;
    }
  }
  /* tuning options: set returns the previous value, a negative value restores the default, unknown keys are errors */
  {
    int64_t prev = -7;
    if (cgx_set_option(CGX_OPT_P2P_TIMEOUT_MS, 5, &prev) != CGX_OK || prev != 20000 ||
        cgx_get_option(CGX_OPT_P2P_TIMEOUT_MS) != 5 || cgx_set_option(CGX_OPT_P2P_TIMEOUT_MS, -1, NULL) != CGX_OK ||
        cgx_get_option(CGX_OPT_P2P_TIMEOUT_MS) != 20000 || cgx_set_option(CGX_OPT_COUNT_, 1, NULL) == CGX_OK ||
        cgx_get_option(-1) != -1) {
      puts("options"); ++failures;
    }
    if (sizeof(cgx_csr) != 112) { puts("cgx_csr layout"); ++failures; }
  }
  printf("abi_smoke: %d failure(s)\n", failures);
  return failures;
}
