// Micro-benchmark behind the round-2 GEMM design: what bounds a 128 x 256 x 16 tcgen05.mma stream on one SM --
// the tensor pipe, or shared-memory bandwidth shared between the MMA's operand reads, the TMA writes of the activation tiles
// and the dequant warps' A-tile stores?  And: does the A operand work from TMEM (tcgen05.mma "ts" form, A written by
// tcgen05.st) with the layout lane = row, column c = {k = 2c, 2c + 1}?
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ummabench tools/ummabench.cu && ./ummabench
//
// Modes (one CTA per SM, 148 CTAs, 4-stage ring, per stage 4 MMAs of K = 16 each, fp16 in / fp32 accumulate in TMEM):
//   ss        A and B descriptors on static shared memory, nothing else running
//   ss+tma    + a bulk-copy engine stream of 32 KB per stage into the B ring (mbarrier-paced like the real kernel)
//   ss+tma+st + eight warps storing a 16 KB A stage per stage (st.shared.v4 + fence.proxy.async), mbarrier-paced
//   ts        A operand from TMEM (static), B from shared memory
//   ts+tma    + the B stream
//   ts+tma+st + eight warps writing the 32-column A stage with tcgen05.st per stage
// Test infrastructure: not linked into libhqq_b200.so.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#define CK(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e__ = (x);                                                                     \
    if (e__ != cudaSuccess) {                                                                  \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__);         \
      exit(1);                                                                                 \
    }                                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void mma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc512(uint32_t* dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(dst)) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc512(uint32_t t) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(t) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
      "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
        "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]),
        "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ inline uint32_t make_idesc_f16(int UN) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= (uint32_t)(UN >> 3) << 17;
  d |= (uint32_t)(128 >> 4) << 24;
  return d;
}

constexpr int kStages = 4;
constexpr int A_STAGE = 128 * 128;  // 128 rows x 64 fp16
constexpr int UN = 256;
constexpr int B_STAGE = UN * 128;
constexpr int SMEM_BYTES = kStages * (A_STAGE + B_STAGE) + 1024 + 256;
constexpr int kThreads = 64 + 256;

enum { F_TS = 1, F_TMA = 2, F_ST = 4, F_K2 = 8, F_N128 = 16 };  // F_K2: two K-interleaved accumulators; F_N128: two N = 128 accumulators

__global__ void __launch_bounds__(kThreads, 1) bench_kernel(int flags, int iters, const uint8_t* __restrict__ gsrc, long long* cycles_out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (A_STAGE + B_STAGE));
  uint64_t* full_a = bars;
  uint64_t* full_b = bars + kStages;
  uint64_t* empty = bars + 2 * kStages;
  uint64_t* done = bars + 3 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool ts = flags & F_TS, tma = flags & F_TMA, st = flags & F_ST;
  // benign operand contents: zeros (fp16 0) -- timing does not depend on the values
  for (int i = threadIdx.x; i < kStages * (A_STAGE + B_STAGE) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_a[s], 8); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(done, 1);
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc512(tmem_slot);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a0 = tmem_base + 256;  // A ring: 4 stages x 32 columns behind the 256 accumulator columns
  const uint8_t* src = gsrc + (size_t)blockIdx.x * (1u << 20);  // 1 MB window per CTA, L2 resident

  if (warp == 0) {
    if (lane == 0 && tma) {
      for (int kb = 0; kb < iters; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
        mbar_expect_tx(&full_b[s], B_STAGE);
        bulk_g2s(sB + s * B_STAGE, src + (size_t)(kb % 32) * B_STAGE, B_STAGE, &full_b[s]);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_f16(UN);
    long long t0 = 0;
    for (int kb = 0; kb < iters; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      if (st) mbar_wait(&full_a[s], ph);
      if (tma) mbar_wait(&full_b[s], ph);
      if (!st && !tma) mbar_wait(&empty[s], ph ^ 1);  // bound the number of MMAs in flight to one ring
      tc_fence_after();
      if (kb == 8) t0 = clock64();
      if (lane == 0) {
        const uint64_t adesc = make_desc_sw128(smem_u32(sA + s * A_STAGE));
        const uint64_t bdesc = make_desc_sw128(smem_u32(sB + s * B_STAGE));
        if (flags & F_K2) {
          // does the accumulate chain on ONE TMEM tile bound the stream?  even k -> accumulator 0, odd k -> accumulator 1 (256 columns each)
#pragma unroll
          for (int k = 0; k < 4; ++k)
            mma_ss(tmem_base + (uint32_t)(k & 1) * 256, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | (k >> 1)) != 0);
        } else if (flags & F_N128) {
          // the same work as eight N = 128 instructions on two independent accumulators (token halves of the B tile)
          const uint32_t idesc128 = make_idesc_f16(128);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            mma_ss(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc128, (kb | k) != 0);
            mma_ss(tmem_base + 128, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(1024 + k * 2), idesc128, (kb | k) != 0);
          }
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            if (ts) mma_ts(tmem_base, tmem_a0 + s * 32 + k * 8, bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
            else mma_ss(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
          }
        }
        tc_commit(&empty[s]);
        if (kb == iters - 1) tc_commit(done);
      }
      __syncwarp();
    }
    mbar_wait(done, 0);
    const long long t1 = clock64();
    if (lane == 0) cycles_out[blockIdx.x] = t1 - t0;
  } else if (st) {
    const int td = threadIdx.x - 64;
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    for (int kb = 0; kb < iters; ++kb) {
      const int s = kb % kStages;
      mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
      if (ts) {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = 0x3c003c00u + (uint32_t)(kb & 1);
        tmem_st16(tmem_a0 + ((uint32_t)(quarter * 32) << 16) + s * 32 + half * 16, v);
        tmem_wait_st();
        tc_fence_before();
      } else {
        // 16 KB per stage from 256 threads: 4 x 16 bytes each, conflict-free (consecutive threads, consecutive chunks)
        const uint32_t base = smem_u32(sA + s * A_STAGE) + td * 16;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("st.shared.v4.b32 [%0], {%1,%1,%1,%1};" ::"r"(base + i * 4096), "r"((uint32_t)(kb & 1)) : "memory");
        fence_async_smem();
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&full_a[s]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc512(tmem_base);
  }
}

// ---- functional check of the TS form: D_ss = A_smem x B^T, D_ts = A_tmem x B^T, N = 128, K = 64 ---------------------------------
__global__ void __launch_bounds__(128, 1) ts_check_kernel(const __half* __restrict__ A, const __half* __restrict__ B, float* __restrict__ Dss,
                                                          float* __restrict__ Dts) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;              // 128 x 64 fp16, SW128 K-major
  uint8_t* sB = smem + A_STAGE;    // 128 x 64 fp16
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 2 * A_STAGE);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = threadIdx.x;  // one row of A and of B per thread
  for (int k = 0; k < 64; ++k) {
    const uint32_t off = r * 128 + ((((uint32_t)k >> 3) ^ (r & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__half*>(sA + off) = A[r * 64 + k];
    *reinterpret_cast<__half*>(sB + off) = B[r * 64 + k];
  }
  if (warp == 0) {
    if (lane == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    __syncwarp();
    tmem_alloc512(tmem_slot);
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_a = tmem_base + 256;
  // A row r -> TMEM lane r, columns 256..287: column c holds {k = 2c (low half), k = 2c + 1 (high half)}
  {
    uint32_t v[16];
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int k = h * 32 + 2 * i;
        v[i] = (uint32_t)__half_as_ushort(A[r * 64 + k]) | ((uint32_t)__half_as_ushort(A[r * 64 + k + 1]) << 16);
      }
      tmem_st16(tmem_a + ((uint32_t)(warp * 32) << 16) + h * 16, v);
    }
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc_f16(128);
    const uint64_t adesc = make_desc_sw128(smem_u32(sA)), bdesc = make_desc_sw128(smem_u32(sB));
    for (int k = 0; k < 4; ++k) mma_ss(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, k != 0);
    for (int k = 0; k < 4; ++k) mma_ts(tmem_base + 128, tmem_a + k * 8, bdesc + (uint64_t)(k * 2), idesc, k != 0);
    tc_commit(bar);
  }
  mbar_wait(bar, 0);
  tc_fence_after();
  for (int c = 0; c < 256; c += 32) {
    uint32_t v[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + c, v);
    float* D = c < 128 ? Dss : Dts;
    for (int j = 0; j < 32; ++j) D[r * 128 + (c & 127) + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc512(tmem_base); }
}

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev));
  printf("device %s, %d SMs, max clock %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);

  // ---- functional check ----
  {
    std::vector<__half> hA(128 * 64), hB(128 * 64);
    srand(1);
    for (auto& v : hA) v = __float2half((float)(rand() % 15 - 7));
    for (auto& v : hB) v = __float2half((float)(rand() % 9 - 4));
    __half *dA, *dB;
    float *dss, *dts;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2));
    CK(cudaMalloc(&dss, 128 * 128 * 4)); CK(cudaMalloc(&dts, 128 * 128 * 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaFuncSetAttribute(ts_check_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * A_STAGE + 2048));
    ts_check_kernel<<<1, 128, 2 * A_STAGE + 2048>>>(dA, dB, dss, dts);
    CK(cudaDeviceSynchronize());
    std::vector<float> hss(128 * 128), hts(128 * 128);
    CK(cudaMemcpy(hss.data(), dss, hss.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hts.data(), dts, hts.size() * 4, cudaMemcpyDeviceToHost));
    int bad_ss = 0, bad_ts = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < 128; ++n) {
        float ref = 0.f;
        for (int k = 0; k < 64; ++k) ref += __half2float(hA[m * 64 + k]) * __half2float(hB[n * 64 + k]);
        bad_ss += hss[m * 128 + n] != ref;
        bad_ts += hts[m * 128 + n] != ref;
      }
    printf("functional: SS mismatches %d / 16384, TS (A from TMEM) mismatches %d / 16384\n", bad_ss, bad_ts);
  }

  // ---- throughput ----
  uint8_t* gsrc;
  CK(cudaMalloc(&gsrc, (size_t)148 << 20));
  CK(cudaMemset(gsrc, 0, (size_t)148 << 20));
  long long* dcyc;
  CK(cudaMalloc(&dcyc, 148 * sizeof(long long)));
  CK(cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int iters = 4096;  // stages per CTA: 4 MMAs of 128 x 256 x 16 each
  struct { const char* name; int flags; } modes[] = {{"ss", 0}, {"ss+tma", F_TMA}, {"ss+tma+st", F_TMA | F_ST}, {"ss+st", F_ST},
                                                      {"ts", F_TS}, {"ts+tma", F_TS | F_TMA}, {"ts+tma+st", F_TS | F_TMA | F_ST}, {"ts+st", F_TS | F_ST},
                                                      {"ss k2", F_K2}, {"ss k2+st", F_K2 | F_ST}, {"ss n128x2", F_N128}, {"ss n128x2+st", F_N128 | F_ST}};
  for (auto& m : modes) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEvent_t e0, e1;
      CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
      CK(cudaEventRecord(e0));
      bench_kernel<<<148, kThreads, SMEM_BYTES>>>(m.flags, iters, gsrc, dcyc);
      CK(cudaEventRecord(e1));
      CK(cudaDeviceSynchronize());
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      std::vector<long long> cyc(148);
      CK(cudaMemcpy(cyc.data(), dcyc, 148 * sizeof(long long), cudaMemcpyDeviceToHost));
      long long mx = 0, mn = 1LL << 60;
      for (auto c : cyc) { mx = c > mx ? c : mx; mn = c < mn ? c : mn; }
      const double flops = 148.0 * iters * 4 * 2.0 * 128 * 256 * 16;
      if (rep == 1)
        printf("%-10s: %8.3f ms  %7.1f TFLOP/s   cycles per 64-k stage: min %.1f max %.1f (tensor floor 512)\n", m.name, ms, flops / ms / 1e9,
               (double)mn / (iters - 8), (double)mx / (iters - 8));
    }
  }
  return 0;
}
