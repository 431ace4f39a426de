// gemm_tc.cu — tcgen05 split-bf16 GEMM (B2CTR_GEMM_BF16X3).  Placeholder until the tensor-core
// path lands: fails loudly, never falls back.
#include "common.cuh"
namespace b2ctr {
size_t gemm_bf16x3_workspace_bytes(const b2ctr_gemm_t*) { return 0; }
b2ctr_status_t gemm_bf16x3(const b2ctr_gemm_t*, void*, size_t, cudaStream_t) {
  set_error("gemm: precision BF16X3 (tcgen05) is not built into this library yet");
  return B2CTR_ERR_UNSUPPORTED;
}
}  // namespace b2ctr
