// hqq_b200_linear_fwd: routing between the fused forward kernels.
#include "common.cuh"

namespace hqq {
bool small_route_ok(int64_t M, int64_t N, int64_t K, int gs, int nbits, int axis, int dtype);
int linear_small(const void* x, const void* Wq, const void* scale, const void* zero, const void* bias, void* y, int64_t M,
                 int64_t N, int64_t K, int gs, int nbits, int dtype, cudaStream_t st);
bool gemm_route_ok(int64_t M, int64_t N, int64_t K, int gs, int nbits, int axis, int dtype);
size_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int gs, int nbits, int dtype);
int linear_gemm(const void* x, const void* Wq, const void* scale, const void* zero, const void* bias, void* y, int64_t M,
                int64_t N, int64_t K, int gs, int nbits, int dtype, void* ws, size_t ws_bytes, cudaStream_t st);
}  // namespace hqq

using namespace hqq;

extern "C" int hqq_b200_linear_fwd_route(int64_t M, int64_t N, int64_t K, int group_size, int nbits, int axis, int dtype) {
  if (small_route_ok(M, N, K, group_size, nbits, axis, dtype)) return 1;
  if (gemm_route_ok(M, N, K, group_size, nbits, axis, dtype)) return 2;
  return 0;
}

extern "C" size_t hqq_b200_linear_fwd_workspace_bytes(int64_t M, int64_t N, int64_t K, int group_size, int nbits, int dtype) {
  if (small_route_ok(M, N, K, group_size, nbits, 1, dtype)) return 0;
  if (gemm_route_ok(M, N, K, group_size, nbits, 1, dtype)) return gemm_workspace_bytes(M, N, K, group_size, nbits, dtype);
  return 0;
}

extern "C" int hqq_b200_linear_fwd(const void* x, const void* W_q, const void* scale, const void* zero, const void* bias, void* y,
                                   int64_t M, int64_t N, int64_t K, int group_size, int nbits, int axis, int dtype, void* workspace,
                                   size_t workspace_bytes, void* stream) {
  HQQ_REQUIRE(x && W_q && scale && zero && y, HQQ_E_INVALID, "hqq_b200_linear_fwd: null pointer");
  HQQ_REQUIRE(M > 0 && N > 0 && K > 0 && group_size > 0, HQQ_E_INVALID, "hqq_b200_linear_fwd: bad shape M=%lld N=%lld K=%lld gs=%d",
              (long long)M, (long long)N, (long long)K, group_size);
  HQQ_REQUIRE(valid_nbits(nbits), HQQ_E_INVALID, "nbits=%d not supported.", nbits);
  HQQ_REQUIRE(axis == 0 || axis == 1, HQQ_E_INVALID, "axis should be either 0 or 1");
  cudaStream_t st = (cudaStream_t)stream;
  const int route = hqq_b200_linear_fwd_route(M, N, K, group_size, nbits, axis, dtype);
  if (route == 1) return linear_small(x, W_q, scale, zero, bias, y, M, N, K, group_size, nbits, dtype, st);
  if (route == 2) return linear_gemm(x, W_q, scale, zero, bias, y, M, N, K, group_size, nbits, dtype, workspace, workspace_bytes, st);
  set_error("hqq_b200_linear_fwd: no fused kernel for M=%lld N=%lld K=%lld gs=%d nbits=%d axis=%d dtype=%d", (long long)M, (long long)N,
            (long long)K, group_size, nbits, axis, dtype);
  return HQQ_E_UNSUPPORTED;
}
