// tcgen05 / TMA fused dequant-GEMM (large M).  Placeholder until the kernel lands: nothing routes here.
#include "common.cuh"
namespace hqq {
bool gemm_route_ok(int64_t, int64_t, int64_t, int, int, int, int) { return false; }
size_t gemm_workspace_bytes(int64_t, int64_t, int64_t, int, int, int) { return 0; }
int linear_gemm(const void*, const void*, const void*, const void*, const void*, void*, int64_t, int64_t, int64_t, int, int, int, void*,
                size_t, cudaStream_t) {
  set_error("hqq_b200_linear_fwd: tcgen05 GEMM path not built");
  return HQQ_E_UNSUPPORTED;
}
}  // namespace hqq
