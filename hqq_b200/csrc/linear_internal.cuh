// Declarations shared by the routing layer (linear.cu) and the fused forward kernels (linear_small.cu, linear_gemm.cu).
#pragma once
#include "common.cuh"

namespace hqq {

// host-side view of the optional tagged-word exchange of the M = 1 decode kernel (see SKArgs in linear_small.cu and
// hqq_b200_decode_desc in include/hqq_b200.h)
struct TpExchange {
  int tp, rank;
  void* const* peer_data;
  const void* red_data;
  void* const* y_tagged;
  const void* x_tagged;
  const void* x2_tagged;
  const int* step_ctr;
  int x_index, x_per_step;
};

}  // namespace hqq
