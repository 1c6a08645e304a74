// mma.sync m16n8k16 f16 latency / throughput on sm_100a (legacy tensor path), per SM sub-partition.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int CHAINS>
__global__ void k(float* out, long long* cyc, int iters, uint32_t seed) {
  float d[CHAINS][4];
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) d[c][i] = 0.f;
  uint32_t a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, b0 = seed + 4, b1 = seed + 5;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) mma(d[c], a0, a1, a2, a3, b0, b1);
  }
  long long t1 = clock64();
  float s = 0;
  for (int c = 0; c < CHAINS; ++c) for (int i = 0; i < 4; ++i) s += d[c][i];
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  if (s == 123.456f) out[0] = s;
}

template <int CHAINS>
void run(int warps_per_cta) {
  float* out; long long* cyc; cudaMalloc(&out, 4); cudaMalloc(&cyc, 8);
  const int iters = 2048;
  k<CHAINS><<<148, warps_per_cta * 32>>>(out, cyc, iters, 0);
  k<CHAINS><<<148, warps_per_cta * 32>>>(out, cyc, iters, 0);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / (iters * CHAINS);
  printf("chains=%d warps/CTA=%2d : %.2f cycles per HMMA per warp ; per-SMSP issue interval %.2f cycles (%s)\n", CHAINS, warps_per_cta, per,
         per / ((warps_per_cta + 3) / 4), cudaGetErrorString(cudaGetLastError()));
}

int main() {
  run<1>(1); run<2>(1); run<4>(1); run<8>(1);
  run<1>(4); run<4>(4); run<8>(4);
  run<1>(16); run<2>(16); run<4>(16); run<8>(16);
  return 0;
}
