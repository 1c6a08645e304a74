// HQQLinear.forward for ONE token and 3-bit weights (3bit_32 packing), fused unpack -> dequant -> dot, HBM-bound.
//
//   y[N] = x[K] @ dequantize(W_q)^T (+bias)        reference: hqq/core/quantize.py:184-199, 880-898; packing bitpack.py:69-91
//
// EXPERIMENTAL (HQQ_B200_FUSED_3BIT=1): written after round 1's GPU budget was spent, never run.  Without the knob 3-bit layers
// take the dequantize kernel + library GEMM, as before.
//
// Layout.  The [R, 64] group view (R = N*K/64 rows) is cut into ten slabs of step = ceil(R/10) rows; field f (bits 27-3f..29-3f)
// of packed word (i, c) is level (i + f*step, c).  step is not a multiple of Gk = K/64 (groups per output row), so the ten slabs
// cut output rows at ten different offsets and a 16-row MMA tile has no compact image in the packed tensor.  This kernel
// therefore walks the PACKED words, once, in chunks of Gk packed rows (= K words, one output row's worth per field): inside a
// chunk, field f holds the tail of output row nA_f (its groups kbA_f .. Gk-1) followed by the head of row nA_f + 1.  A CTA reduces
// both pieces of all ten fields over its 8 warps in a fixed order and stores each into a private fp32 slot of the row:
//     slot 0: the piece that holds the row's first group, slot 1: the piece that holds its last, slot 2: a piece cut on both
//     sides (only where a slab boundary falls inside the row)
// -- at most three pieces per row, one writer per slot, no atomics; a second tiny kernel adds the slots in slot order and rounds.
// tests/test_fused3_layout_cpu.py executes exactly this decomposition on the oracle's packed layout.
//
// Arithmetic: like the 4/2/1-bit one-token kernel the levels are not dequantised per element:
//     sum_k x_k (q_k - z) s = s * (sum_k q_k x_k) - s z * (sum_k x_k),  fp32 throughout, per-group sums of x precomputed.
// Algorithmic bytes: N*K*0.4 (packed) + 2*R*2 (scale, zero) + (K + N)*2.
#include <stdlib.h>

#include "common.cuh"

namespace hqq {

struct L3Args {
  const int32_t* Wq;
  const void* scale;
  const void* zero;
  const void* x;
  float* y3;  // [N][3] slots, zero on entry
  int N, K, Gk;
  int R, step;        // grouped rows, packed rows (both < 2^31 - 16: host check)
  int step_q, step_r; // step / Gk, step % Gk
  int nchunks;
};

constexpr int kL3Threads = 256;
constexpr int kL3Fields = 10;

template <typename T>
__global__ void __launch_bounds__(kL3Threads) linear3_decode1_kernel(const L3Args a) {
  extern __shared__ __align__(16) float l3_smem[];
  float* xs = l3_smem;            // [K]  activations in fp32
  float* X = xs + a.K;            // [Gk] per-group sums of x
  float* red = X + a.Gk;          // [8][20] per-warp piece sums of the current chunk
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Gk = a.Gk;

  const T* x = reinterpret_cast<const T*>(a.x);
  for (int k = tid; k < a.K; k += kL3Threads) xs[k] = to_f32<T>(x[k]);
  __syncthreads();
  for (int g = warp; g < Gk; g += kL3Threads / 32) {
    float s = xs[g * 64 + lane] + xs[g * 64 + 32 + lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) X[g] = s;
  }
  __syncthreads();

  // Field f of a chunk starts at grouped row i0 + f*step with i0 = chunk*Gk: its first group index inside the output row,
  // kbA_f = (f*step) % Gk, does not depend on the chunk (its output row is chunk + (f*step) / Gk, recomputed by the writers below).
  int kbA[kL3Fields];
  {
    int kb = 0;
#pragma unroll
    for (int f = 0; f < kL3Fields; ++f) {
      kbA[f] = kb;
      kb += a.step_r;
      if (kb >= Gk) kb -= Gk;
    }
  }
  const T* scale = reinterpret_cast<const T*>(a.scale);
  const T* zero = reinterpret_cast<const T*>(a.zero);
  const int rows_per_warp = (Gk + 7) / 8;

  for (int chunk = blockIdx.x; chunk < a.nchunks; chunk += gridDim.x) {
    const int i0 = chunk * Gk;
    const int rows_c = min(Gk, a.step - i0);  // packed rows of this chunk (the last chunk may be short)
    const int j0 = warp * rows_per_warp, j1 = min(j0 + rows_per_warp, rows_c);
    float acc0[kL3Fields], acc1[kL3Fields];
    int nvalid[kL3Fields];
#pragma unroll
    for (int f = 0; f < kL3Fields; ++f) {
      acc0[f] = 0.0f; acc1[f] = 0.0f;
      nvalid[f] = min(rows_c, a.R - (i0 + f * a.step));  // rows past R exist only as zero padding of the last slab
    }
#pragma unroll 4
    for (int j = j0; j < j1; ++j) {
      const uint2 w = __ldg(reinterpret_cast<const uint2*>(a.Wq + ((long long)(i0 + j) * 64 + 2 * lane)));
#pragma unroll
      for (int f = 0; f < kL3Fields; ++f) {
        if (j < nvalid[f]) {  // warp-uniform
          const int b = Gk - kbA[f];          // rows j >= b belong to the next output row
          const bool p = j >= b;
          const int kb = p ? j - b : kbA[f] + j;
          const int r = i0 + f * a.step + j;
          const float s = to_f32<T>(scale[r]), z = to_f32<T>(zero[r]);  // one address per warp: broadcast
          const float q0 = __uint_as_float(0x4B000000u | ((w.x >> (27 - 3 * f)) & 7u)) - 8388608.0f;  // 2^23 + q, exact
          const float q1 = __uint_as_float(0x4B000000u | ((w.y >> (27 - 3 * f)) & 7u)) - 8388608.0f;
          const float2 xv = *reinterpret_cast<const float2*>(xs + kb * 64 + 2 * lane);
          float d = s * fmaf(q1, xv.y, q0 * xv.x);
          if (lane == 0) d -= s * z * X[kb];
          if (p) acc1[f] += d; else acc0[f] += d;
        }
      }
    }
#pragma unroll
    for (int f = 0; f < kL3Fields; ++f) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        acc0[f] += __shfl_xor_sync(0xffffffffu, acc0[f], o);
        acc1[f] += __shfl_xor_sync(0xffffffffu, acc1[f], o);
      }
      if (lane == 0) { red[warp * 20 + 2 * f] = acc0[f]; red[warp * 20 + 2 * f + 1] = acc1[f]; }
    }
    __syncthreads();
    if (tid < 2 * kL3Fields) {
      const int f = tid >> 1, p = tid & 1;
      // same constants as above, for a runtime f
      int kb = 0, n = 0;
      for (int g = 0; g < f; ++g) {
        kb += a.step_r; n += a.step_q;
        if (kb >= Gk) { kb -= Gk; ++n; }
      }
      const int nv = min(rows_c, a.R - (i0 + f * a.step));
      if (nv > 0) {
        const int cnt0 = min(nv, Gk - kb), cnt = p ? nv - cnt0 : cnt0;
        if (cnt > 0) {
          float v = 0.0f;
#pragma unroll
          for (int wv = 0; wv < kL3Threads / 32; ++wv) v += red[wv * 20 + tid];  // warp order: deterministic
          const int row = chunk + n + p;
          const int slot = (p == 1 || kb == 0) ? 0 : ((kb + cnt0 == Gk) ? 1 : 2);
          a.y3[(long long)row * 3 + slot] = v;
        }
      }
    }
    __syncthreads();  // red is reused by the next chunk
  }
}

template <typename T> __device__ __forceinline__ T l3_round(float v);
template <> __device__ __forceinline__ __half l3_round<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 l3_round<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <typename T>
__global__ void __launch_bounds__(256) linear3_finish_kernel(const float* __restrict__ y3, const T* __restrict__ bias, T* __restrict__ y, int N) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float v = (y3[(long long)n * 3] + y3[(long long)n * 3 + 1]) + y3[(long long)n * 3 + 2];  // slot order
  T o = l3_round<T>(v);
  if (bias) o = __hadd(o, bias[n]);  // out += bias: second rounding, as in the reference
  y[n] = o;
}

static bool fused3_enabled() {
  const char* e = getenv("HQQ_B200_FUSED_3BIT");
  return e && e[0] == '1';
}

bool fused3_route_ok(int64_t M, int64_t N, int64_t K, int gs, int nbits, int axis, int dtype) {
  if (!fused3_enabled()) return false;
  if (M != 1 || nbits != 3 || axis != 1 || gs != 64) return false;
  if (dtype != HQQ_F16 && dtype != HQQ_BF16) return false;
  if (N < 1 || K < 64 || K % 64 != 0 || K > 16384) return false;
  if (N * (K / 64) >= (1LL << 31) - 16) return false;
  return true;
}

size_t fused3_workspace_bytes(int64_t N) { return (size_t)((N * 3 * sizeof(float) + 255) & ~(int64_t)255); }

template <typename T>
static int fused3_typed(const L3Args& a, const void* bias, void* y, cudaStream_t st) {
  const size_t smem = (size_t)(a.K + a.Gk + (kL3Threads / 32) * 2 * kL3Fields) * sizeof(float);
  auto k = linear3_decode1_kernel<T>;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: cannot reserve %zu bytes of shared memory: %s", smem, cudaGetErrorString(e));
    smem_set = smem;
  }
  cudaError_t e = cudaMemsetAsync(a.y3, 0, (size_t)a.N * 3 * sizeof(float), st);
  HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: slot memset failed: %s", cudaGetErrorString(e));
  const int grid = a.nchunks < 3 * kNumSMs ? a.nchunks : 3 * kNumSMs;
  k<<<grid, kL3Threads, smem, st>>>(a);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/3bit");
  linear3_finish_kernel<T><<<(unsigned)cdiv(a.N, 256), 256, 0, st>>>(a.y3, reinterpret_cast<const T*>(bias), reinterpret_cast<T*>(y), a.N);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/3bit-finish");
  return HQQ_OK;
}

int linear_fused3(const void* x, const void* Wq, const void* scale, const void* zero, const void* bias, void* y, int64_t N, int64_t K,
                  int dtype, void* ws, size_t ws_bytes, cudaStream_t st) {
  HQQ_REQUIRE(ws != nullptr && ws_bytes >= fused3_workspace_bytes(N) && aligned(ws, 16), HQQ_E_WORKSPACE,
              "hqq_b200_linear_fwd: the fused 3-bit kernel needs a workspace of %zu bytes (got %zu)", fused3_workspace_bytes(N), ws_bytes);
  HQQ_REQUIRE(aligned(Wq, 8), HQQ_E_INVALID, "hqq_b200_linear_fwd: W_q must be 8-byte aligned");
  L3Args a;
  a.Wq = reinterpret_cast<const int32_t*>(Wq); a.scale = scale; a.zero = zero; a.x = x; a.y3 = reinterpret_cast<float*>(ws);
  a.N = (int)N; a.K = (int)K; a.Gk = (int)(K / 64);
  a.R = (int)(N * (K / 64));
  a.step = (int)cdiv(a.R, 10);
  a.step_q = a.step / a.Gk; a.step_r = a.step % a.Gk;
  a.nchunks = (int)cdiv(a.step, a.Gk);
  if (dtype == HQQ_F16) return fused3_typed<__half>(a, bias, y, st);
  return fused3_typed<__nv_bfloat16>(a, bias, y, st);
}

}  // namespace hqq
