// Streaming-read micro-benchmark: which per-warp access pattern / staging mechanism reaches HBM speed on B200 when
// reading a [rows x pitch] byte matrix as 8-row x 256-byte "units" (the small-M fused forward's weight traffic).
// Build on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/membench tools/membench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint4 ldg_na(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ void cp16(void* s, const void* g) {
  uint32_t a = (uint32_t)__cvta_generic_to_shared(s);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g) : "memory");
}

// MODE 0: LDG, lane (r=lane>>2, c=lane&3): 8 rows x 64 B per instruction (the kernel's mapping), DIST units prefetched in registers
// MODE 1: same mapping through a 4-stage cp.async ring
// MODE 2: LDG, lane (r=lane>>4, c=lane&15): 2 rows x 256 B per instruction
// MODE 3: LDG, fully linear 512 B per instruction (upper bound, ignores the row structure)
template <int MODE, int DIST>
__global__ void __launch_bounds__(256, 2) reader(const uint8_t* __restrict__ base, long long pitch, int KB, long long total_units, uint32_t* out,
                                                   int cta_split) {
  extern __shared__ __align__(16) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  long long u0, u1;
  if (cta_split) {
    const long long U0 = total_units * blockIdx.x / gridDim.x, U1 = total_units * (blockIdx.x + 1) / gridDim.x;
    u0 = U0 + (U1 - U0) * warp / 8; u1 = U0 + (U1 - U0) * (warp + 1) / 8;
  } else {
    const long long TW = (long long)gridDim.x * 8, gw = (long long)blockIdx.x * 8 + warp;
    u0 = total_units * gw / TW; u1 = total_units * (gw + 1) / TW;
  }
  uint32_t acc = 0;
  auto addr = [&](long long u, int i) -> const uint8_t* {
    const long long t = u / KB; const int kb = (int)(u % KB);
    if (MODE == 0 || MODE == 1) return base + (t * 8 + (lane >> 2)) * pitch + kb * 256 + i * 64 + (lane & 3) * 16;
    if (MODE == 2) return base + (t * 8 + i * 2 + (lane >> 4)) * pitch + kb * 256 + (lane & 15) * 16;
    return base + u * 2048 + i * 512 + lane * 16;
  };
  if (MODE == 1) {
    uint4* ring = reinterpret_cast<uint4*>(smem);
    int issued = 0; const int n = (int)(u1 - u0);
    for (int s = 0; s < 3; ++s) {
      if (issued < n) { for (int i = 0; i < 4; ++i) cp16(&ring[(s * 4 + i) * 256 + tid], addr(u0 + issued, i)); ++issued; }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
    int stage = 0;
    for (int k = 0; k < n; ++k) {
      int is = (stage + 3) & 3;
      if (issued < n) { for (int i = 0; i < 4; ++i) cp16(&ring[(is * 4 + i) * 256 + tid], addr(u0 + issued, i)); ++issued; }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_group 3;" ::: "memory");
      for (int i = 0; i < 4; ++i) { uint4 v = ring[(stage * 4 + i) * 256 + tid]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
      stage = (stage + 1) & 3;
    }
  } else {
    uint4 buf[DIST][4];
    const int n = (int)(u1 - u0);
#pragma unroll
    for (int d = 0; d < DIST; ++d)
      if (d < n)
#pragma unroll
        for (int i = 0; i < 4; ++i) buf[d][i] = ldg_na(addr(u0 + d, i));
    for (int k = 0; k < n; k += DIST) {
#pragma unroll
      for (int d = 0; d < DIST; ++d) {
        if (k + d < n) {
          uint4 cur[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) cur[i] = buf[d][i];
          if (k + d + DIST < n)
#pragma unroll
            for (int i = 0; i < 4; ++i) buf[d][i] = ldg_na(addr(u0 + k + d + DIST, i));
#pragma unroll
          for (int i = 0; i < 4; ++i) acc ^= cur[i].x ^ cur[i].y ^ cur[i].z ^ cur[i].w;
        }
      }
    }
  }
  if (acc == 0x12345678u) out[0] = acc;
}

template <int MODE, int DIST>
void run(const char* name, uint8_t* bufs[], int nbuf, long long rows, long long pitch, int grid, int cta_split, uint32_t* out) {
  const int KB = (int)(pitch / 256);
  const long long total = rows / 8 * KB;
  const int smem = MODE == 1 ? 4 * 4 * 256 * 16 : 0;
  cudaFuncSetAttribute(reader<MODE, DIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < nbuf; ++i) reader<MODE, DIST><<<grid, 256, smem>>>(bufs[i], pitch, KB, total, out, cta_split);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int rep = 0; rep < 3; ++rep)
    for (int i = 0; i < nbuf; ++i) reader<MODE, DIST><<<grid, 256, smem>>>(bufs[i], pitch, KB, total, out, cta_split);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double us = ms * 1e3 / (3 * nbuf);
  printf("%-34s rows=%lld pitch=%lld grid=%d split=%d : %8.2f us  %7.0f GB/s  (%s)\n", name, rows, pitch, grid, cta_split, us,
         rows * pitch / us / 1e3, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const int nbuf = 8;
  struct Shape { long long rows, pitch; } shapes[] = {{7168, 4096}, {2048, 14336}, {2048, 4096}, {512, 4096}};
  uint32_t* out; cudaMalloc(&out, 4);
  for (auto sh : shapes) {
    uint8_t* bufs[nbuf];
    for (int i = 0; i < nbuf; ++i) { cudaMalloc(&bufs[i], sh.rows * sh.pitch); cudaMemset(bufs[i], i + 1, sh.rows * sh.pitch); }
    for (int grid : {148, 296}) {
      run<0, 1>("ldg r8x64B dist1", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
      run<0, 2>("ldg r8x64B dist2", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
      run<0, 4>("ldg r8x64B dist4", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
      run<0, 2>("ldg r8x64B dist2 ctasplit", bufs, nbuf, sh.rows, sh.pitch, grid, 1, out);
      run<1, 1>("cp.async ring4 r8x64B", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
      run<1, 1>("cp.async ring4 r8x64B ctasplit", bufs, nbuf, sh.rows, sh.pitch, grid, 1, out);
      run<2, 2>("ldg r2x256B dist2", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
      run<2, 4>("ldg r2x256B dist4", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
      run<3, 2>("ldg linear dist2", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
      run<3, 4>("ldg linear dist4", bufs, nbuf, sh.rows, sh.pitch, grid, 0, out);
    }
    for (int i = 0; i < nbuf; ++i) cudaFree(bufs[i]);
  }
  return 0;
}
