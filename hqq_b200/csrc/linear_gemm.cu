// HQQLinear.forward for large M (prefill / batched decode): fused unpack -> group-dequant -> tcgen05 GEMM.
//
//   y[M,N] = x[M,K] @ dequantize(W_q)^T (+bias)          reference: hqq/core/quantize.py:184-199, 880-898
//
// One CTA computes a [128 weight rows] x [UN tokens] output tile, accumulated in TMEM (fp32) by tcgen05.mma:
//   A operand (M=128 of the UMMA) = the weight tile.  The packed bytes keep the reference's slab layout (bitpack.py),
//       so 128/F packed rows x F slabs give 128 output rows.  Eight "dequant" warps stream the packed bytes from HBM
//       (read exactly once per token tile), expand them in registers with the reference's two roundings
//       W_r = fl(fl(q - z) * s) (bit-identical to Quantizer.dequantize) and store the fp16/bf16 tile into shared memory
//       in the K-major SWIZZLE_128B layout the tensor core reads.  The dequantised matrix never exists in HBM.
//   B operand (N=UN of the UMMA) = the activation tile [UN tokens x 64 k], fetched by TMA (cp.async.bulk.tensor, 128B
//       swizzle, out-of-range tokens zero-filled by the hardware).
//   4-stage mbarrier ring: TMA warp / dequant warps fill, one elected thread issues the MMAs, tcgen05.commit frees the
//   stage.  Epilogue: the dequant warps read the accumulator with tcgen05.ld (lane = weight row, column = token), add
//   the bias and store y (32 consecutive n per token -> coalesced).
// sm_100a only: tcgen05 / TMEM / TMA, no mma.sync fallback.
#include <cuda.h>  // CUtensorMap types only; the encode entry point is resolved through the runtime (no -lcuda)

#include "common.cuh"

namespace hqq {

namespace gemm {

constexpr int kStages = 4;
constexpr int kBlockK = 64;          // k elements per stage = one 128-byte swizzle row
constexpr int kTileRows = 128;       // weight rows per CTA = UMMA M
constexpr int kDequantThreads = 256;
constexpr int kThreads = 64 + kDequantThreads;  // warp 0: TMA + TMEM alloc, warp 1: MMA issue, warps 2..9: dequant + epilogue

struct Args {
  const uint8_t* Wq;
  const void* scale;
  const void* zero;
  const void* bias;
  void* y;
  int M, N, K;
  int step;  // packed rows = N / F
  int Gk;    // groups per row = K / GS
};

// ---- PTX wrappers -------------------------------------------------------------------------------------------------
#ifdef HQQ_EMU
// CPU emulation (tests/emu): the same entry points, backed by a functional model of mbarrier / TMA / tcgen05 / TMEM
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return ::emu::smem_offset(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) { ::emu::mbar_init(bar, count); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) { ::emu::mbar_arrive(bar); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { ::emu::mbar_expect_tx(bar, bytes); }
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) { ::emu::mbar_wait(bar, parity); }
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) { ::emu::tma_load_2d(smem_dst, map, bar, c0, c1); }
__device__ __forceinline__ void fence_async_smem() {}
__device__ __forceinline__ void fence_barrier_init() {}
__device__ __forceinline__ void tc_fence_before() {}
__device__ __forceinline__ void tc_fence_after() {}
__device__ __forceinline__ void tc_commit(uint64_t* bar) { ::emu::tc_commit(bar); }  // arrives once the MMAs issued before it have executed
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  ::emu::umma_f16(tmem_d, adesc, bdesc, idesc, accumulate);
}
template <int NCOLS> __device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem) { ::emu::tmem_alloc(dst_in_smem, NCOLS); }
template <int NCOLS> __device__ __forceinline__ void tmem_dealloc(uint32_t) {}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) { ::emu::tmem_ld32(taddr, v); }
#define HQQ_STS_V4(addr, a, b, c, d) ::emu::sts(addr, a, b, c, d)
#define HQQ_STS_V2(addr, a, b) ::emu::sts(addr, a, b)
#define HQQ_PREFETCH_TENSORMAP(p) ((void)(p))
#define HQQ_NAMED_BAR_SYNC(id, n) ::emu::named_barrier(id, n)
#else
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
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(smem_dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_in_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_in_smem)), "n"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS) : "memory");
}
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

#define HQQ_STS_V4(addr, a, b, c, d) asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory")
#define HQQ_STS_V2(addr, a, b) asm volatile("st.shared.v2.b32 [%0], {%1,%2};" ::"r"(addr), "r"(a), "r"(b) : "memory")
#define HQQ_PREFETCH_TENSORMAP(p) asm volatile("prefetch.tensormap [%0];" ::"l"(p) : "memory")
#define HQQ_NAMED_BAR_SYNC(id, n) asm volatile("bar.sync %0, %1;" ::"n"(id), "n"(n) : "memory")
#endif  // HQQ_EMU

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address >> 4 in bits [0,14),
// leading byte offset (unused for swizzled K-major, 1) in [16,30), stride byte offset = 1024 B between 8-row groups in
// [32,46), descriptor version 1 (Blackwell) in [46,48), layout type 2 = SWIZZLE_128B in [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): D = F32, A/B = F16 or BF16, both K-major, M = 128, N = UN.
template <typename T>
__device__ __forceinline__ uint32_t make_idesc(int UN) {
  const uint32_t fmt = std::is_same<T, __nv_bfloat16>::value ? 1u : 0u;
  uint32_t d = 0;
  d |= 1u << 4;                      // c_format = F32
  d |= fmt << 7;                     // a_format
  d |= fmt << 10;                    // b_format
  d |= (uint32_t)(UN >> 3) << 17;    // n_dim
  d |= (uint32_t)(128 >> 4) << 24;   // m_dim
  return d;
}

// ---- level -> T with the reference's roundings -------------------------------------------------------------------
// Two k-adjacent levels (bytes b0, b1 already masked to the field) -> T2 {fl(fl(q0 - z) * s), fl(fl(q1 - z) * s)}.
__device__ __forceinline__ uint32_t prmt_b32(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef HQQ_EMU
  return ::emu::prmt(a, b, sel);
#else
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
#endif
}

// Four k-adjacent levels (one per byte of `t`, already masked to the field) -> two packed T2
//   {fl(fl(q0 - z) * s), fl(fl(q1 - z) * s)}, {.. q2, q3 ..}
template <typename T> struct Pair;
template <> struct Pair<__half> {
  using T2 = __half2;
  __device__ __forceinline__ static void deq4(uint32_t t, __half2 z2, __half2 s2, uint32_t& lo, uint32_t& hi) {
    // byte | 0x6400 == 1024 + q exactly (one PRMT per pair); subtracting 1024 is exact, so (q - z) and (.. * s) round
    // exactly like the reference's two steps
    const __half2 k1024 = __half2half2(__ushort_as_half((unsigned short)0x6400));
    uint32_t a = prmt_b32(t, 0x64646464u, 0x4140u), b = prmt_b32(t, 0x64646464u, 0x4342u);
    __half2 ha = __hmul2(__hsub2(__hsub2(*reinterpret_cast<__half2*>(&a), k1024), z2), s2);
    __half2 hb = __hmul2(__hsub2(__hsub2(*reinterpret_cast<__half2*>(&b), k1024), z2), s2);
    lo = *reinterpret_cast<uint32_t*>(&ha);
    hi = *reinterpret_cast<uint32_t*>(&hb);
  }
  __device__ __forceinline__ static __half2 bcast(__half v) { return __half2half2(v); }
};
template <> struct Pair<__nv_bfloat16> {
  using T2 = __nv_bfloat162;
  __device__ __forceinline__ static void deq4(uint32_t t, __nv_bfloat162 z2, __nv_bfloat162 s2, uint32_t& lo, uint32_t& hi) {
    // levels < 256 are exact in bf16 (8 significant bits); convert through the exact float 2^23 + q trick
    const float f0 = __uint_as_float(0x4B000000u | (t & 0xFFu)) - 8388608.0f, f1 = __uint_as_float(0x4B000000u | ((t >> 8) & 0xFFu)) - 8388608.0f;
    const float f2 = __uint_as_float(0x4B000000u | ((t >> 16) & 0xFFu)) - 8388608.0f, f3 = __uint_as_float(0x4B000000u | (t >> 24)) - 8388608.0f;
    __nv_bfloat162 ha = __hmul2(__hsub2(__floats2bfloat162_rn(f0, f1), z2), s2);
    __nv_bfloat162 hb = __hmul2(__hsub2(__floats2bfloat162_rn(f2, f3), z2), s2);
    lo = *reinterpret_cast<uint32_t*>(&ha);
    hi = *reinterpret_cast<uint32_t*>(&hb);
  }
  __device__ __forceinline__ static __nv_bfloat162 bcast(__nv_bfloat16 v) { return __bfloat162bfloat162(v); }
};

template <typename T> __device__ __forceinline__ T cvt_out(float v);
template <> __device__ __forceinline__ __half cvt_out<__half>(float v) { return __float2half_rn(v); }
template <> __device__ __forceinline__ __nv_bfloat16 cvt_out<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

template <int UN>
struct Smem {
  static constexpr int A_STAGE = kTileRows * 128;  // 128 rows x 128 B
  static constexpr int B_STAGE = UN * 128;
  static constexpr int BYTES = kStages * (A_STAGE + B_STAGE) + 1024 /*align*/ + 256 /*barriers*/;
};

// DQ = dequant threads: 256 (default: 8 warps) or 512 (HQQ_B200_GEMM_VARIANT=dq16, experimental: 16 warps, each expanding half as
// many bytes per stage -- ncu shows the eight dequant warps latency-bound (37 % issue-active) with the tensor pipe 61 % active, so
// twice the warps per SM is the cheapest way to hide their HBM / shared-memory / barrier latencies; same MMAs, same results)
template <typename T, int NBITS, int GS, int UN, int DQ = kDequantThreads>
__global__ void __launch_bounds__(64 + DQ, 1) linear_gemm_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
  constexpr int F = 8 / NBITS;             // slabs per byte
  constexpr int PR = kTileRows / F;        // packed rows per tile
  constexpr int BPT = 64 * PR / DQ;        // packed bytes per dequant thread and k-block (32 / F with 256 threads)
  static_assert(BPT >= 4 && (DQ == 256 || DQ == 512), "a dequant thread expands at least four packed bytes per k-block");
  constexpr int TPR = 64 / BPT;            // dequant threads per packed row
  constexpr uint32_t MASK = (1u << NBITS) - 1u;
  using S = Smem<UN>;
  using P2 = Pair<T>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * S::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (S::A_STAGE + S::B_STAGE));
  uint64_t* full_a = bars;                 // [kStages] dequant warps -> MMA (one arrival per warp)
  uint64_t* full_b = bars + kStages;       // [kStages] TMA -> MMA (1 arrival + tx bytes)
  uint64_t* empty = bars + 2 * kStages;    // [kStages] MMA (tcgen05.commit) -> producers
  uint64_t* accum_full = bars + 3 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int prow0 = tile_n * PR;           // first packed row of the tile
  const int m0 = tile_m * UN;
  const int num_kb = a.K / kBlockK;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_a[s], DQ / 32); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(accum_full, 1);
      fence_barrier_init();
      HQQ_PREFETCH_TENSORMAP(&xmap);
    }
    __syncwarp();
    tmem_alloc<UN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
        mbar_expect_tx(&full_b[s], S::B_STAGE);
        tma_load_2d(sB + s * S::B_STAGE, &xmap, &full_b[s], kb * kBlockK, m0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one elected thread) =================
    const uint32_t idesc = make_idesc<T>(UN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&full_a[s], ph);
      mbar_wait(&full_b[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_desc_sw128(smem_u32(sA + s * S::A_STAGE));
        const uint64_t bdesc = make_desc_sw128(smem_u32(sB + s * S::B_STAGE));
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)  // UMMA_K = 16: advance 32 bytes inside the 128-byte swizzle row
          tc_mma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        tc_commit(&empty[s]);                          // frees the stage when these MMAs have read it
        if (kb == num_kb - 1) tc_commit(accum_full);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ================= dequant warps: packed bytes -> swizzled fp16/bf16 A tile =================
    static_assert(kStages == 4, "the dequant loop is unrolled over the 4 ring stages");
    const int td = threadIdx.x - 64;
    const int pr = td / TPR, c = td % TPR;
    const bool row_ok = (prow0 + pr) < a.step;
    const uint8_t* wptr = a.Wq + (long long)(row_ok ? prow0 + pr : 0) * a.K + c * BPT;
    // per-slab meta rows and shared-memory offsets are loop invariant
    constexpr int GPQ = 256 / GS;  // quantisation groups per 4 k-blocks (4 or 2): one vector load per slab and array
    const T* sptr[F];
    const T* zptr[F];
    uint32_t soff[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const long long mrow = (long long)(row_ok ? f * a.step + prow0 + pr : 0) * a.Gk;
      sptr[f] = reinterpret_cast<const T*>(a.scale) + mrow;
      zptr[f] = reinterpret_cast<const T*>(a.zero) + mrow;
      const int row = f * PR + pr;
      // K-major SWIZZLE_128B: 16-byte chunk index XOR (row % 8) inside each 8-row x 128-byte atom
      if constexpr (BPT >= 8) soff[f] = (uint32_t)(row * 128) | ((uint32_t)(row & 7) << 16);  // chunk applied below
      else soff[f] = (uint32_t)(row * 128 + (((c >> 1) ^ (row & 7)) << 4) + (c & 1) * 8);
    }
    const uint32_t sA_u32 = smem_u32(sA);

    // Packed bytes and scale/zero for the NEXT four k-blocks sit in registers while the current four are expanded: their
    // HBM/L2 latency stays off the critical path of the 64-k stages.
    uint32_t wbuf[4][BPT / 4];
    Vec<T, GPQ> sv[F], zv[F];
    auto load_w = [&](const uint8_t* p, uint32_t (&w)[BPT / 4]) {
      if constexpr (BPT == 32) { const uint4 v0 = ldg_stream_v4(p), v1 = ldg_stream_v4(p + 16); w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w; }
      else if constexpr (BPT == 16) { const uint4 v = ldg_stream_v4(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
      else if constexpr (BPT == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); w[0] = v.x; w[1] = v.y; }
      else { w[0] = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    };
    auto load_quad = [&]() {  // the four k-blocks starting at wptr, and their groups
#pragma unroll
      for (int d = 0; d < 4; ++d) load_w(wptr + d * kBlockK, wbuf[d]);
#pragma unroll
      for (int f = 0; f < F; ++f) { sv[f] = *reinterpret_cast<const Vec<T, GPQ>*>(sptr[f]); zv[f] = *reinterpret_cast<const Vec<T, GPQ>*>(zptr[f]); }
    };
    load_quad();
    const int num_quads = num_kb >> 2;  // K % 256 == 0 (checked by the router)
    for (int q = 0; q < num_quads; ++q) {
      uint32_t wq[4][BPT / 4];
      typename P2::T2 s2[4][F], z2[4][F];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
#pragma unroll
        for (int i = 0; i < BPT / 4; ++i) wq[d][i] = wbuf[d][i];
#pragma unroll
        for (int f = 0; f < F; ++f) { s2[d][f] = P2::bcast(sv[f].v[(d * kBlockK) / GS]); z2[d][f] = P2::bcast(zv[f].v[(d * kBlockK) / GS]); }
      }
      if (q + 1 < num_quads) {
        wptr += 4 * kBlockK;
#pragma unroll
        for (int f = 0; f < F; ++f) { sptr[f] += GPQ; zptr[f] += GPQ; }
        load_quad();
      }
      const uint32_t parity = (uint32_t)(q & 1) ^ 1u;
#pragma unroll
      for (int d = 0; d < 4; ++d) {  // stage index == d because the ring has exactly four stages
        mbar_wait(&empty[d], parity);
        const uint32_t stage = sA_u32 + d * S::A_STAGE;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const int sh = 8 - NBITS * (f + 1);
          uint32_t out[BPT / 2];  // BPT levels -> BPT/2 packed pairs
#pragma unroll
          for (int i = 0; i < BPT / 4; ++i) {
            const uint32_t t = (wq[d][i] >> sh) & (MASK * 0x01010101u);
            P2::deq4(t, z2[d][f], s2[d][f], out[2 * i], out[2 * i + 1]);
          }
          if constexpr (BPT >= 8) {
            const uint32_t rowbase = stage + (soff[f] & 0xFFFFu), rx = soff[f] >> 16;
#pragma unroll
            for (int ch = 0; ch < BPT / 8; ++ch) {
              const uint32_t addr = rowbase + (((uint32_t)(c * (BPT / 8) + ch) ^ rx) << 4);
              HQQ_STS_V4(addr, out[4 * ch], out[4 * ch + 1], out[4 * ch + 2], out[4 * ch + 3]);
            }
          } else {  // BPT == 4: half a chunk
            HQQ_STS_V2(stage + soff[f], out[0], out[1]);
          }
        }
        fence_async_smem();  // make the generic-proxy stores visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[d]);  // one arrival per warp: every lane has fenced its stores before the syncwarp
      }
    }

    // ================= epilogue: TMEM -> registers -> y =================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter .. +31
    constexpr int NPART = DQ / 128;               // warps sharing a quarter (2 or 4): they split the token columns
    static_assert(UN / NPART >= 32, "every epilogue warp reads at least one 32-column slab");
    const int half = (warp - 2) >> 2;
    const int t = quarter * 32 + lane;            // tile row = weight row inside the tile
    const int tf = t / PR, tp = t % PR;
    const bool n_ok = (prow0 + tp) < a.step;
    const int n = tf * a.step + prow0 + tp;
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const bool has_bias = bias != nullptr;
    T bn = cvt_out<T>(0.0f);
    if (has_bias && n_ok) bn = bias[n];
#pragma unroll 1
    for (int col = half * (UN / NPART); col < (half + 1) * (UN / NPART); col += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int m = m0 + col + j;
        if (n_ok && m < a.M) {
          T o = cvt_out<T>(__uint_as_float(v[j]));
          if (has_bias) o = __hadd(o, bn);  // out += bias: second rounding, as in the reference
          y[(long long)m * a.N + n] = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<UN>(tmem_base);
  }
}

// =====================================================================================================================
// Variant "un512" (HQQ_B200_GEMM_VARIANT=un512, experimental -- written after round 1's GPU budget was spent, not yet run):
// ncu on linear_gemm_kernel shows the dequant warps, not the tensor pipe, on the critical path (tensor pipe 61 % active).  Their
// work per weight is already near its floor for the reference's two-rounding dequant (2.5 ALU ops per weight), so the lever is
// amortisation: here every dequantised A stage feeds TWO 128 x 256 accumulators (all 512 TMEM columns), i.e. 512 tokens per
// weight tile instead of 256 -- half the dequant work, packed-byte traffic and A-stage stores per flop.  The A ring keeps four
// 16 KB stages, the B ring has two 64 KB stages (two TMA boxes of 256 tokens each) with their own empty barriers.
struct Smem512 {
  static constexpr int UN = 512, UNH = 256, kStagesB = 2;
  static constexpr int A_STAGE = kTileRows * 128;
  static constexpr int B_HALF = UNH * 128;
  static constexpr int B_STAGE = 2 * B_HALF;
  static constexpr int BYTES = kStages * A_STAGE + kStagesB * B_STAGE + 1024 /*align*/ + 256 /*barriers*/;
};

// DQ = 512 (HQQ_B200_GEMM_VARIANT=un512dq): sixteen dequant warps, as in linear_gemm_kernel<..., 512>
template <typename T, int NBITS, int GS, int DQ = kDequantThreads>
__global__ void __launch_bounds__(64 + DQ, 1) linear_gemm_un512_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
  constexpr int F = 8 / NBITS;             // slabs per byte
  constexpr int PR = kTileRows / F;        // packed rows per tile
  constexpr int BPT = 64 * PR / DQ;        // packed bytes per dequant thread and k-block (32 / F with 256 threads)
  static_assert(BPT >= 4 && (DQ == 256 || DQ == 512), "a dequant thread expands at least four packed bytes per k-block");
  constexpr int TPR = 64 / BPT;            // dequant threads per packed row
  constexpr uint32_t MASK = (1u << NBITS) - 1u;
  using S = Smem512;
  constexpr int UN = S::UN, UNH = S::UNH;
  using P2 = Pair<T>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * S::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * S::A_STAGE + S::kStagesB * S::B_STAGE);
  uint64_t* full_a = bars;                      // [kStages]  dequant warps -> MMA (one arrival per warp)
  uint64_t* empty = full_a + kStages;           // [kStages]  MMA (tcgen05.commit) -> dequant warps
  uint64_t* full_b = empty + kStages;           // [kStagesB] TMA -> MMA (1 arrival + tx bytes of both halves)
  uint64_t* empty_b = full_b + S::kStagesB;     // [kStagesB] MMA (tcgen05.commit) -> TMA
  uint64_t* accum_full = empty_b + S::kStagesB;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int prow0 = tile_n * PR;           // first packed row of the tile
  const int m0 = tile_m * UN;
  const int num_kb = a.K / kBlockK;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_a[s], DQ / 32); mbar_init(&empty[s], 1); }
      for (int s = 0; s < S::kStagesB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
      mbar_init(accum_full, 1);
      fence_barrier_init();
      HQQ_PREFETCH_TENSORMAP(&xmap);
    }
    __syncwarp();
    tmem_alloc<UN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % S::kStagesB;
        mbar_wait(&empty_b[s], ((kb / S::kStagesB) & 1) ^ 1);
        mbar_expect_tx(&full_b[s], S::B_STAGE);  // both boxes; a box past the last token is zero-filled and still counts in full
        tma_load_2d(sB + s * S::B_STAGE, &xmap, &full_b[s], kb * kBlockK, m0);
        tma_load_2d(sB + s * S::B_STAGE + S::B_HALF, &xmap, &full_b[s], kb * kBlockK, m0 + UNH);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one elected thread) =================
    const uint32_t idesc = make_idesc<T>(UNH);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kStages, sb = kb % S::kStagesB;
      mbar_wait(&full_a[s], (kb / kStages) & 1);
      mbar_wait(&full_b[sb], (kb / S::kStagesB) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_desc_sw128(smem_u32(sA + s * S::A_STAGE));
        const uint64_t bdesc0 = make_desc_sw128(smem_u32(sB + sb * S::B_STAGE));
        const uint64_t bdesc1 = make_desc_sw128(smem_u32(sB + sb * S::B_STAGE + S::B_HALF));
        // the same dequantised A stage feeds two accumulators (TMEM columns 0..255 and 256..511): 512 tokens per weight tile
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)  // UMMA_K = 16: advance 32 bytes inside the 128-byte swizzle row
          tc_mma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc0 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          tc_mma_f16(tmem_base + (uint32_t)UNH, adesc + (uint64_t)(k * 2), bdesc1 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        tc_commit(&empty[s]);                          // frees the A stage when these MMAs have read it
        tc_commit(&empty_b[sb]);                       // and the B stage
        if (kb == num_kb - 1) tc_commit(accum_full);   // accumulators complete
      }
      __syncwarp();
    }
  } else {
    // ================= dequant warps: packed bytes -> swizzled fp16/bf16 A tile =================
    static_assert(kStages == 4, "the dequant loop is unrolled over the 4 ring stages");
    const int td = threadIdx.x - 64;
    const int pr = td / TPR, c = td % TPR;
    const bool row_ok = (prow0 + pr) < a.step;
    const uint8_t* wptr = a.Wq + (long long)(row_ok ? prow0 + pr : 0) * a.K + c * BPT;
    // per-slab meta rows and shared-memory offsets are loop invariant
    constexpr int GPQ = 256 / GS;  // quantisation groups per 4 k-blocks (4 or 2): one vector load per slab and array
    const T* sptr[F];
    const T* zptr[F];
    uint32_t soff[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const long long mrow = (long long)(row_ok ? f * a.step + prow0 + pr : 0) * a.Gk;
      sptr[f] = reinterpret_cast<const T*>(a.scale) + mrow;
      zptr[f] = reinterpret_cast<const T*>(a.zero) + mrow;
      const int row = f * PR + pr;
      // K-major SWIZZLE_128B: 16-byte chunk index XOR (row % 8) inside each 8-row x 128-byte atom
      if constexpr (BPT >= 8) soff[f] = (uint32_t)(row * 128) | ((uint32_t)(row & 7) << 16);  // chunk applied below
      else soff[f] = (uint32_t)(row * 128 + (((c >> 1) ^ (row & 7)) << 4) + (c & 1) * 8);
    }
    const uint32_t sA_u32 = smem_u32(sA);

    // Packed bytes and scale/zero for the NEXT four k-blocks sit in registers while the current four are expanded: their
    // HBM/L2 latency stays off the critical path of the 64-k stages.
    uint32_t wbuf[4][BPT / 4];
    Vec<T, GPQ> sv[F], zv[F];
    auto load_w = [&](const uint8_t* p, uint32_t (&w)[BPT / 4]) {
      if constexpr (BPT == 32) { const uint4 v0 = ldg_stream_v4(p), v1 = ldg_stream_v4(p + 16); w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w; }
      else if constexpr (BPT == 16) { const uint4 v = ldg_stream_v4(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
      else if constexpr (BPT == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); w[0] = v.x; w[1] = v.y; }
      else { w[0] = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    };
    auto load_quad = [&]() {  // the four k-blocks starting at wptr, and their groups
#pragma unroll
      for (int d = 0; d < 4; ++d) load_w(wptr + d * kBlockK, wbuf[d]);
#pragma unroll
      for (int f = 0; f < F; ++f) { sv[f] = *reinterpret_cast<const Vec<T, GPQ>*>(sptr[f]); zv[f] = *reinterpret_cast<const Vec<T, GPQ>*>(zptr[f]); }
    };
    load_quad();
    const int num_quads = num_kb >> 2;  // K % 256 == 0 (checked by the router)
    for (int q = 0; q < num_quads; ++q) {
      uint32_t wq[4][BPT / 4];
      typename P2::T2 s2[4][F], z2[4][F];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
#pragma unroll
        for (int i = 0; i < BPT / 4; ++i) wq[d][i] = wbuf[d][i];
#pragma unroll
        for (int f = 0; f < F; ++f) { s2[d][f] = P2::bcast(sv[f].v[(d * kBlockK) / GS]); z2[d][f] = P2::bcast(zv[f].v[(d * kBlockK) / GS]); }
      }
      if (q + 1 < num_quads) {
        wptr += 4 * kBlockK;
#pragma unroll
        for (int f = 0; f < F; ++f) { sptr[f] += GPQ; zptr[f] += GPQ; }
        load_quad();
      }
      const uint32_t parity = (uint32_t)(q & 1) ^ 1u;
#pragma unroll
      for (int d = 0; d < 4; ++d) {  // stage index == d because the ring has exactly four stages
        mbar_wait(&empty[d], parity);
        const uint32_t stage = sA_u32 + d * S::A_STAGE;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const int sh = 8 - NBITS * (f + 1);
          uint32_t out[BPT / 2];  // BPT levels -> BPT/2 packed pairs
#pragma unroll
          for (int i = 0; i < BPT / 4; ++i) {
            const uint32_t t = (wq[d][i] >> sh) & (MASK * 0x01010101u);
            P2::deq4(t, z2[d][f], s2[d][f], out[2 * i], out[2 * i + 1]);
          }
          if constexpr (BPT >= 8) {
            const uint32_t rowbase = stage + (soff[f] & 0xFFFFu), rx = soff[f] >> 16;
#pragma unroll
            for (int ch = 0; ch < BPT / 8; ++ch) {
              const uint32_t addr = rowbase + (((uint32_t)(c * (BPT / 8) + ch) ^ rx) << 4);
              HQQ_STS_V4(addr, out[4 * ch], out[4 * ch + 1], out[4 * ch + 2], out[4 * ch + 3]);
            }
          } else {  // BPT == 4: half a chunk
            HQQ_STS_V2(stage + soff[f], out[0], out[1]);
          }
        }
        fence_async_smem();  // make the generic-proxy stores visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[d]);  // one arrival per warp: every lane has fenced its stores before the syncwarp
      }
    }

    // ================= epilogue: TMEM -> registers -> y =================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter .. +31
    constexpr int NPART = DQ / 128;               // warps sharing a quarter (2 or 4): they split the token columns
    const int half = (warp - 2) >> 2;
    const int t = quarter * 32 + lane;            // tile row = weight row inside the tile
    const int tf = t / PR, tp = t % PR;
    const bool n_ok = (prow0 + tp) < a.step;
    const int n = tf * a.step + prow0 + tp;
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const bool has_bias = bias != nullptr;
    T bn = cvt_out<T>(0.0f);
    if (has_bias && n_ok) bn = bias[n];
#pragma unroll 1
    for (int col = half * (UN / NPART); col < (half + 1) * (UN / NPART); col += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int m = m0 + col + j;
        if (n_ok && m < a.M) {
          T o = cvt_out<T>(__uint_as_float(v[j]));
          if (has_bias) o = __hadd(o, bn);  // out += bias: second rounding, as in the reference
          y[(long long)m * a.N + n] = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<UN>(tmem_base);
  }
}

// =====================================================================================================================
// Variant "split-K" (HQQ_B200_GEMM_SPLITK=1, experimental -- written after round 1's GPU budget was spent, not yet run):
// for 32 < M <= ~256 the grid of linear_gemm_kernel is (N/128) x 1 tiles -- 32 CTAs for a 4096-row matrix on 148 SMs, each
// walking all of K.  Here gridDim.z CTAs share an output tile, each accumulating a contiguous k-slice in TMEM; the slices meet
// as fp32 partials in a caller-provided workspace and the last CTA to arrive sums them in slice order (deterministic) and
// applies bias/rounding.  Same tiles, descriptors and dequant as the kernel above.
struct ArgsSK : Args {
  float* ws;           // [gridDim.z][M][N] fp32 partials
  unsigned* counters;  // [tiles_m][tiles_n], zero on entry, zero again on exit
};

template <typename T, int NBITS, int GS, int UN>
__global__ void __launch_bounds__(kThreads, 1) linear_gemm_splitk_kernel(const __grid_constant__ CUtensorMap xmap, const ArgsSK a) {
  constexpr int F = 8 / NBITS;             // slabs per byte
  constexpr int PR = kTileRows / F;        // packed rows per tile
  constexpr int BPT = 64 * PR / kDequantThreads;  // packed bytes per dequant thread and k-block (32 / F)
  constexpr int TPR = 64 / BPT;            // dequant threads per packed row
  constexpr uint32_t MASK = (1u << NBITS) - 1u;
  using S = Smem<UN>;
  using P2 = Pair<T>;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);  // SWIZZLE_128B atoms
  uint8_t* sA = smem;
  uint8_t* sB = smem + kStages * S::A_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * (S::A_STAGE + S::B_STAGE));
  uint64_t* full_a = bars;                 // [kStages] dequant warps -> MMA (one arrival per warp)
  uint64_t* full_b = bars + kStages;       // [kStages] TMA -> MMA (1 arrival + tx bytes)
  uint64_t* empty = bars + 2 * kStages;    // [kStages] MMA (tcgen05.commit) -> producers
  uint64_t* accum_full = bars + 3 * kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 1);
  volatile uint32_t* last_flag = tmem_slot + 1;  // 1 when this CTA is the last of its output tile to finish its k-slice

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int prow0 = tile_n * PR;           // first packed row of the tile
  const int m0 = tile_m * UN;
  // k-slice of this CTA, in quads of four 64-k blocks (the dequant loop's granularity): balanced to within one quad
  const int nq_total = a.K / (4 * kBlockK), KS = (int)gridDim.z, z = (int)blockIdx.z;
  const int q_first = (int)((long long)nq_total * z / KS), q_last = (int)((long long)nq_total * (z + 1) / KS);
  const int kb0 = 4 * q_first;
  const int num_kb = 4 * (q_last - q_first);

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_a[s], kDequantThreads / 32); mbar_init(&full_b[s], 1); mbar_init(&empty[s], 1); }
      mbar_init(accum_full, 1);
      fence_barrier_init();
      HQQ_PREFETCH_TENSORMAP(&xmap);
    }
    __syncwarp();
    tmem_alloc<UN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        mbar_wait(&empty[s], ((kb / kStages) & 1) ^ 1);
        mbar_expect_tx(&full_b[s], S::B_STAGE);
        tma_load_2d(sB + s * S::B_STAGE, &xmap, &full_b[s], (kb0 + kb) * kBlockK, m0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one elected thread) =================
    const uint32_t idesc = make_idesc<T>(UN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % kStages;
      const uint32_t ph = (kb / kStages) & 1;
      mbar_wait(&full_a[s], ph);
      mbar_wait(&full_b[s], ph);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_desc_sw128(smem_u32(sA + s * S::A_STAGE));
        const uint64_t bdesc = make_desc_sw128(smem_u32(sB + s * S::B_STAGE));
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)  // UMMA_K = 16: advance 32 bytes inside the 128-byte swizzle row
          tc_mma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        tc_commit(&empty[s]);                          // frees the stage when these MMAs have read it
        if (kb == num_kb - 1) tc_commit(accum_full);   // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ================= dequant warps: packed bytes -> swizzled fp16/bf16 A tile =================
    static_assert(kStages == 4, "the dequant loop is unrolled over the 4 ring stages");
    const int td = threadIdx.x - 64;
    const int pr = td / TPR, c = td % TPR;
    const bool row_ok = (prow0 + pr) < a.step;
    const uint8_t* wptr = a.Wq + (long long)(row_ok ? prow0 + pr : 0) * a.K + (long long)kb0 * kBlockK + c * BPT;
    // per-slab meta rows and shared-memory offsets are loop invariant
    constexpr int GPQ = 256 / GS;  // quantisation groups per 4 k-blocks (4 or 2): one vector load per slab and array
    const T* sptr[F];
    const T* zptr[F];
    uint32_t soff[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const long long mrow = (long long)(row_ok ? f * a.step + prow0 + pr : 0) * a.Gk;
      sptr[f] = reinterpret_cast<const T*>(a.scale) + mrow + q_first * GPQ;
      zptr[f] = reinterpret_cast<const T*>(a.zero) + mrow + q_first * GPQ;
      const int row = f * PR + pr;
      // K-major SWIZZLE_128B: 16-byte chunk index XOR (row % 8) inside each 8-row x 128-byte atom
      if constexpr (BPT >= 8) soff[f] = (uint32_t)(row * 128) | ((uint32_t)(row & 7) << 16);  // chunk applied below
      else soff[f] = (uint32_t)(row * 128 + (((c >> 1) ^ (row & 7)) << 4) + (c & 1) * 8);
    }
    const uint32_t sA_u32 = smem_u32(sA);

    // Packed bytes and scale/zero for the NEXT four k-blocks sit in registers while the current four are expanded: their
    // HBM/L2 latency stays off the critical path of the 64-k stages.
    uint32_t wbuf[4][BPT / 4];
    Vec<T, GPQ> sv[F], zv[F];
    auto load_w = [&](const uint8_t* p, uint32_t (&w)[BPT / 4]) {
      if constexpr (BPT == 32) { const uint4 v0 = ldg_stream_v4(p), v1 = ldg_stream_v4(p + 16); w[0] = v0.x; w[1] = v0.y; w[2] = v0.z; w[3] = v0.w; w[4] = v1.x; w[5] = v1.y; w[6] = v1.z; w[7] = v1.w; }
      else if constexpr (BPT == 16) { const uint4 v = ldg_stream_v4(p); w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w; }
      else if constexpr (BPT == 8) { const uint2 v = __ldg(reinterpret_cast<const uint2*>(p)); w[0] = v.x; w[1] = v.y; }
      else { w[0] = __ldg(reinterpret_cast<const uint32_t*>(p)); }
    };
    auto load_quad = [&]() {  // the four k-blocks starting at wptr, and their groups
#pragma unroll
      for (int d = 0; d < 4; ++d) load_w(wptr + d * kBlockK, wbuf[d]);
#pragma unroll
      for (int f = 0; f < F; ++f) { sv[f] = *reinterpret_cast<const Vec<T, GPQ>*>(sptr[f]); zv[f] = *reinterpret_cast<const Vec<T, GPQ>*>(zptr[f]); }
    };
    load_quad();
    const int num_quads = num_kb >> 2;  // K % 256 == 0 (checked by the router)
    for (int q = 0; q < num_quads; ++q) {
      uint32_t wq[4][BPT / 4];
      typename P2::T2 s2[4][F], z2[4][F];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
#pragma unroll
        for (int i = 0; i < BPT / 4; ++i) wq[d][i] = wbuf[d][i];
#pragma unroll
        for (int f = 0; f < F; ++f) { s2[d][f] = P2::bcast(sv[f].v[(d * kBlockK) / GS]); z2[d][f] = P2::bcast(zv[f].v[(d * kBlockK) / GS]); }
      }
      if (q + 1 < num_quads) {
        wptr += 4 * kBlockK;
#pragma unroll
        for (int f = 0; f < F; ++f) { sptr[f] += GPQ; zptr[f] += GPQ; }
        load_quad();
      }
      const uint32_t parity = (uint32_t)(q & 1) ^ 1u;
#pragma unroll
      for (int d = 0; d < 4; ++d) {  // stage index == d because the ring has exactly four stages
        mbar_wait(&empty[d], parity);
        const uint32_t stage = sA_u32 + d * S::A_STAGE;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const int sh = 8 - NBITS * (f + 1);
          uint32_t out[BPT / 2];  // BPT levels -> BPT/2 packed pairs
#pragma unroll
          for (int i = 0; i < BPT / 4; ++i) {
            const uint32_t t = (wq[d][i] >> sh) & (MASK * 0x01010101u);
            P2::deq4(t, z2[d][f], s2[d][f], out[2 * i], out[2 * i + 1]);
          }
          if constexpr (BPT >= 8) {
            const uint32_t rowbase = stage + (soff[f] & 0xFFFFu), rx = soff[f] >> 16;
#pragma unroll
            for (int ch = 0; ch < BPT / 8; ++ch) {
              const uint32_t addr = rowbase + (((uint32_t)(c * (BPT / 8) + ch) ^ rx) << 4);
              HQQ_STS_V4(addr, out[4 * ch], out[4 * ch + 1], out[4 * ch + 2], out[4 * ch + 3]);
            }
          } else {  // BPT == 4: half a chunk
            HQQ_STS_V2(stage + soff[f], out[0], out[1]);
          }
        }
        fence_async_smem();  // make the generic-proxy stores visible to the tensor core (async proxy)
        __syncwarp();
        if (lane == 0) mbar_arrive(&full_a[d]);  // one arrival per warp: every lane has fenced its stores before the syncwarp
      }
    }

    // ================= epilogue: TMEM -> registers -> y =================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter .. +31
    const int half = (warp - 2) >> 2;             // two warps share a quarter: split the token columns
    const int t = quarter * 32 + lane;            // tile row = weight row inside the tile
    const int tf = t / PR, tp = t % PR;
    const bool n_ok = (prow0 + tp) < a.step;
    const int n = tf * a.step + prow0 + tp;
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const bool has_bias = bias != nullptr;
    T bn = cvt_out<T>(0.0f);
    if (has_bias && n_ok) bn = bias[n];
    float* wsz = a.ws + (size_t)z * (size_t)a.M * (size_t)a.N;  // this k-slice's fp32 partial of y
#pragma unroll 1
    for (int col = half * (UN / 2); col < (half + 1) * (UN / 2); col += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int m = m0 + col + j;
        if (n_ok && m < a.M) __stcg(&wsz[(size_t)m * a.N + n], __uint_as_float(v[j]));
      }
    }
    // last-arriver reduction (the threadFenceReduction pattern): every slice publishes its partial, bumps the tile's counter, and
    // the CTA that observes KS-1 sums the KS partials in slice order -- a fixed order, so the result does not depend on timing
    __threadfence();
    HQQ_NAMED_BAR_SYNC(1, kDequantThreads);
    unsigned* ctr = a.counters + ((size_t)tile_m * gridDim.x + tile_n);
    if (td == 0) *last_flag = (atomicAdd(ctr, 1u) == (unsigned)(KS - 1)) ? 1u : 0u;
    HQQ_NAMED_BAR_SYNC(1, kDequantThreads);
    if (*last_flag) {
      __threadfence();
#pragma unroll 1
      for (int col = half * (UN / 2); col < (half + 1) * (UN / 2); col += 32) {
#pragma unroll 4
        for (int j = 0; j < 32; ++j) {
          const int m = m0 + col + j;
          if (n_ok && m < a.M) {
            float acc = 0.0f;
            for (int zz = 0; zz < KS; ++zz) acc += __ldcg(&a.ws[((size_t)zz * a.M + m) * (size_t)a.N + n]);
            T o = cvt_out<T>(acc);
            if (has_bias) o = __hadd(o, bn);  // out += bias: second rounding, as in the reference
            y[(long long)m * a.N + n] = o;
          }
        }
      }
      if (td == 0) *ctr = 0u;  // leave the counter clean for a replay of the same launch (CUDA graphs)
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<UN>(tmem_base);
  }
}

// =====================================================================================================================
// Variant "ld" (HQQ_B200_GEMM_VARIANT=ld, experimental -- written after round 1's GPU budget was spent, not yet run):
// ncu on the kernel above shows the dequant warps, not the tensor pipe, on the critical path (tensor pipe 61 % active), and
// a third of their stall samples sit on `fence.proxy.async` and on the first use of the register-prefetched bytes: the
// proxy fence waits for the thread's OWN outstanding global loads, so the one-quad-ahead prefetch is serialised behind DRAM
// latency at every stage.  Here the dequant warps never touch global memory: a loader warp streams the packed tile and its
// scale/zero into shared-memory rings with cp.async (completion signalled through an mbarrier by
// cp.async.mbarrier.arrive.noinc, up to 8 k-blocks ahead), the dequant warps read them with LDS.  Everything downstream
// (swizzled A stage, tcgen05.mma, TMEM epilogue) is unchanged; B gets 3 stages to make room for the rings.
constexpr int kStagesB = 3;
constexpr int kMetaSlots = 4;
constexpr int kLdThreads = 96 + kDequantThreads;  // warp 0: TMA(B) + TMEM alloc, warp 1: MMA, warp 2: loader, warps 3..10: dequant + epilogue

template <int UN, int NBITS>
struct SmemLd {
  static constexpr int PR = kTileRows / (8 / NBITS);
  static constexpr int A_STAGE = kTileRows * 128;
  static constexpr int B_STAGE = UN * 128;
  static constexpr int W_STAGE = PR * kBlockK;                                        // packed bytes of one k-block
  static constexpr int NW = (32 * 1024 / W_STAGE) < 8 ? (32 * 1024 / W_STAGE) : 8;    // 8 k-blocks ahead (8-bit: 4)
  static constexpr int M_SLOT = kTileRows * 2 * 8;                                    // {scale, zero} x 128 rows x <= 4 groups x 2 B
  static constexpr int BYTES = kStages * A_STAGE + kStagesB * B_STAGE + NW * W_STAGE + kMetaSlots * M_SLOT + 1024 /*align*/ + 512 /*barriers*/;
  static_assert(4 * kMetaSlots >= NW + 4, "a meta slot must outlive the W stages of its four k-blocks");
};

template <int BYTES>
__device__ __forceinline__ void cp_async_b(uint32_t smem_addr, const void* g) {
#ifdef HQQ_EMU
  ::emu::cp_async(::emu::smem_ptr(smem_addr), g, BYTES);  // lands when the mbarrier it is tied to says so
#else
  if constexpr (BYTES == 16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g) : "memory");
  else asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr), "l"(g), "n"(BYTES) : "memory");
#endif
}
// the mbarrier receives one arrival from this thread once all of its earlier cp.async have landed (the count is part of init)
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
#ifdef HQQ_EMU
  ::emu::cp_async_mbar_arrive(bar);
#else
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
#endif
}

template <typename T, int NBITS, int GS, int UN>
__global__ void __launch_bounds__(kLdThreads, 1) linear_gemm_ld_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
  constexpr int F = 8 / NBITS;
  constexpr int PR = kTileRows / F;
  constexpr int BPT = 64 * PR / kDequantThreads;
  constexpr int TPR = 64 / BPT;
  constexpr int GPQ = 256 / GS;  // groups per four k-blocks
  constexpr uint32_t MASK = (1u << NBITS) - 1u;
  using S = SmemLd<UN, NBITS>;
  using P2 = Pair<T>;
  constexpr int NW = S::NW;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + kStages * S::A_STAGE;
  uint8_t* sW = sB + kStagesB * S::B_STAGE;
  uint8_t* sM = sW + NW * S::W_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sM + kMetaSlots * S::M_SLOT);
  uint64_t* full_a = bars;                       // [kStages]  dequant warps -> MMA (one arrival per warp)
  uint64_t* empty_a = full_a + kStages;          // [kStages]  MMA (tcgen05.commit) -> dequant warps
  uint64_t* full_b = empty_a + kStages;          // [kStagesB] TMA -> MMA
  uint64_t* empty_b = full_b + kStagesB;         // [kStagesB] MMA -> TMA
  uint64_t* full_w = empty_b + kStagesB;         // [NW] loader lanes (32 async arrivals) -> dequant warps
  uint64_t* empty_w = full_w + NW;               // [NW] dequant warps (one arrival per warp) -> loader
  uint64_t* accum_full = empty_w + NW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int prow0 = tile_n * PR;
  const int m0 = tile_m * UN;
  const int num_kb = a.K / kBlockK;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_a[s], kDequantThreads / 32); mbar_init(&empty_a[s], 1); }
      for (int s = 0; s < kStagesB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
      for (int s = 0; s < NW; ++s) { mbar_init(&full_w[s], 32); mbar_init(&empty_w[s], kDequantThreads / 32); }
      mbar_init(accum_full, 1);
      fence_barrier_init();
      HQQ_PREFETCH_TENSORMAP(&xmap);
    }
    __syncwarp();
    tmem_alloc<UN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStagesB;
        mbar_wait(&empty_b[s], ((kb / kStagesB) & 1) ^ 1);
        mbar_expect_tx(&full_b[s], S::B_STAGE);
        tma_load_2d(sB + s * S::B_STAGE, &xmap, &full_b[s], kb * kBlockK, m0);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one elected thread) =================
    const uint32_t idesc = make_idesc<T>(UN);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int sa = kb % kStages, sb = kb % kStagesB;
      mbar_wait(&full_a[sa], (kb / kStages) & 1);
      mbar_wait(&full_b[sb], (kb / kStagesB) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_desc_sw128(smem_u32(sA + sa * S::A_STAGE));
        const uint64_t bdesc = make_desc_sw128(smem_u32(sB + sb * S::B_STAGE));
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          tc_mma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        tc_commit(&empty_a[sa]);
        tc_commit(&empty_b[sb]);
        if (kb == num_kb - 1) tc_commit(accum_full);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ================= loader: packed tile + scale/zero -> shared-memory rings (cp.async, no registers) =================
    const uint32_t sW_u32 = smem_u32(sW), sM_u32 = smem_u32(sM);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int sw = kb % NW;
      mbar_wait(&empty_w[sw], ((kb / NW) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < PR * 4 / 32; ++i) {  // PR rows x four 16-byte chunks, lanes along the row: coalesced 64-byte rows
        const int id = i * 32 + lane, row = id >> 2, ch = id & 3;
        const int prow = prow0 + row;
        const uint8_t* src = a.Wq + (long long)(prow < a.step ? prow : 0) * a.K + kb * kBlockK + ch * 16;
        cp_async_b<16>(sW_u32 + sw * S::W_STAGE + row * kBlockK + ch * 16, src);
      }
      if ((kb & 3) == 0) {  // the groups of k-blocks kb .. kb+3: GPQ values per row and array
        const int slot = (kb >> 2) % kMetaSlots;
#pragma unroll
        for (int i = 0; i < kTileRows * 2 / 32; ++i) {
          const int id = i * 32 + lane, arr = id >> 7, t = id & (kTileRows - 1);
          const int f = t / PR, prow = prow0 + t % PR;
          const long long mrow = (prow < a.step) ? (long long)f * a.step + prow : 0;
          const T* src = reinterpret_cast<const T*>(arr ? a.zero : a.scale) + mrow * a.Gk + (kb >> 2) * GPQ;
          cp_async_b<GPQ * 2>(sM_u32 + slot * S::M_SLOT + (arr * kTileRows + t) * (GPQ * 2), src);
        }
      }
      cp_async_mbar_arrive(&full_w[sw]);
    }
  } else {
    // ================= dequant warps: shared-memory packed bytes -> swizzled fp16/bf16 A tile =================
    static_assert(kStages == 4, "the dequant loop is unrolled over the 4 A stages");
    const int td = threadIdx.x - 96;
    const int pr = td / TPR, c = td % TPR;
    uint32_t soff[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const int row = f * PR + pr;
      if constexpr (BPT >= 8) soff[f] = (uint32_t)(row * 128) | ((uint32_t)(row & 7) << 16);
      else soff[f] = (uint32_t)(row * 128 + (((c >> 1) ^ (row & 7)) << 4) + (c & 1) * 8);
    }
    const uint32_t sA_u32 = smem_u32(sA);
    const int num_quads = num_kb >> 2;  // K % 256 == 0 (checked by the router)
    for (int q = 0; q < num_quads; ++q) {
      typename P2::T2 s2[4][F], z2[4][F];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int kb = 4 * q + d;
        const int sw = kb % NW;
        mbar_wait(&full_w[sw], (kb / NW) & 1);
        if (d == 0) {  // this quad's scale/zero arrived with its first k-block
          const uint8_t* slot = sM + (q % kMetaSlots) * S::M_SLOT;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const Vec<T, GPQ> sv = *reinterpret_cast<const Vec<T, GPQ>*>(slot + (0 * kTileRows + f * PR + pr) * (GPQ * 2));
            const Vec<T, GPQ> zv = *reinterpret_cast<const Vec<T, GPQ>*>(slot + (1 * kTileRows + f * PR + pr) * (GPQ * 2));
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) { s2[dd][f] = P2::bcast(sv.v[(dd * kBlockK) / GS]); z2[dd][f] = P2::bcast(zv.v[(dd * kBlockK) / GS]); }
          }
        }
        uint32_t wq[BPT / 4];
        {
          const uint8_t* p = sW + sw * S::W_STAGE + pr * kBlockK + c * BPT;
          if constexpr (BPT == 32) { const uint4 v0 = *reinterpret_cast<const uint4*>(p), v1 = *reinterpret_cast<const uint4*>(p + 16); wq[0] = v0.x; wq[1] = v0.y; wq[2] = v0.z; wq[3] = v0.w; wq[4] = v1.x; wq[5] = v1.y; wq[6] = v1.z; wq[7] = v1.w; }
          else if constexpr (BPT == 16) { const uint4 v = *reinterpret_cast<const uint4*>(p); wq[0] = v.x; wq[1] = v.y; wq[2] = v.z; wq[3] = v.w; }
          else if constexpr (BPT == 8) { const uint2 v = *reinterpret_cast<const uint2*>(p); wq[0] = v.x; wq[1] = v.y; }
          else { wq[0] = *reinterpret_cast<const uint32_t*>(p); }
        }
        mbar_wait(&empty_a[d], (uint32_t)(q & 1) ^ 1u);  // A stage index == d (four stages, four k-blocks per quad)
        const uint32_t stage = sA_u32 + d * S::A_STAGE;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const int sh = 8 - NBITS * (f + 1);
          uint32_t out[BPT / 2];
#pragma unroll
          for (int i = 0; i < BPT / 4; ++i) {
            const uint32_t t = (wq[i] >> sh) & (MASK * 0x01010101u);
            P2::deq4(t, z2[d][f], s2[d][f], out[2 * i], out[2 * i + 1]);
          }
          if constexpr (BPT >= 8) {
            const uint32_t rowbase = stage + (soff[f] & 0xFFFFu), rx = soff[f] >> 16;
#pragma unroll
            for (int ch = 0; ch < BPT / 8; ++ch) {
              const uint32_t addr = rowbase + (((uint32_t)(c * (BPT / 8) + ch) ^ rx) << 4);
              HQQ_STS_V4(addr, out[4 * ch], out[4 * ch + 1], out[4 * ch + 2], out[4 * ch + 3]);
            }
          } else {
            HQQ_STS_V2(stage + soff[f], out[0], out[1]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&full_a[d]); mbar_arrive(&empty_w[sw]); }  // every lane's packed bytes (and meta) are in registers
      }
    }

    // ================= epilogue: TMEM -> registers -> y =================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter .. +31
    const int half = (warp - 3) >> 2;             // two warps share a quarter: split the token columns
    const int t = quarter * 32 + lane;
    const int tf = t / PR, tp = t % PR;
    const bool n_ok = (prow0 + tp) < a.step;
    const int n = tf * a.step + prow0 + tp;
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const bool has_bias = bias != nullptr;
    T bn = cvt_out<T>(0.0f);
    if (has_bias && n_ok) bn = bias[n];
#pragma unroll 1
    for (int col = half * (UN / 2); col < (half + 1) * (UN / 2); col += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int m = m0 + col + j;
        if (n_ok && m < a.M) {
          T o = cvt_out<T>(__uint_as_float(v[j]));
          if (has_bias) o = __hadd(o, bn);
          y[(long long)m * a.N + n] = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<UN>(tmem_base);
  }
}

// =====================================================================================================================
// Variant "ld512" (HQQ_B200_GEMM_VARIANT=ld512, experimental): the loader-warp kernel above with the two accumulators of
// "un512" -- dequant warps that never touch global memory AND half the dequant work per flop.  Shared memory: 4 x 16 KB A
// stages, 2 x 64 KB B stages, a 16 KB packed-byte ring (8-bit: 2 k-blocks, 4-bit: 4, 2/1-bit: 8), 4 meta slots.
template <int NBITS>
struct SmemLd512 {
  static constexpr int UN = 512, UNH = 256, kStagesB = 2;
  static constexpr int PR = kTileRows / (8 / NBITS);
  static constexpr int A_STAGE = kTileRows * 128;
  static constexpr int B_HALF = UNH * 128;
  static constexpr int B_STAGE = 2 * B_HALF;
  static constexpr int W_STAGE = PR * kBlockK;
  static constexpr int NW = (16 * 1024 / W_STAGE) < 8 ? (16 * 1024 / W_STAGE) : 8;
  static constexpr int M_SLOT = kTileRows * 2 * 8;
  static constexpr int BYTES = kStages * A_STAGE + kStagesB * B_STAGE + NW * W_STAGE + kMetaSlots * M_SLOT + 1024 /*align*/ + 512 /*barriers*/;
  static_assert(4 * kMetaSlots >= NW + 4, "a meta slot must outlive the W stages of its four k-blocks");
  static_assert(BYTES <= 227 * 1024, "shared-memory budget");
};

template <typename T, int NBITS, int GS>
__global__ void __launch_bounds__(kLdThreads, 1) linear_gemm_ld512_kernel(const __grid_constant__ CUtensorMap xmap, const Args a) {
  constexpr int F = 8 / NBITS;
  constexpr int PR = kTileRows / F;
  constexpr int BPT = 64 * PR / kDequantThreads;
  constexpr int TPR = 64 / BPT;
  constexpr int GPQ = 256 / GS;  // groups per four k-blocks
  constexpr uint32_t MASK = (1u << NBITS) - 1u;
  using S = SmemLd512<NBITS>;
  constexpr int UN = S::UN, UNH = S::UNH, kStagesB = S::kStagesB;
  using P2 = Pair<T>;
  constexpr int NW = S::NW;

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* sA = smem;
  uint8_t* sB = sA + kStages * S::A_STAGE;
  uint8_t* sW = sB + kStagesB * S::B_STAGE;
  uint8_t* sM = sW + NW * S::W_STAGE;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sM + kMetaSlots * S::M_SLOT);
  uint64_t* full_a = bars;                       // [kStages]  dequant warps -> MMA (one arrival per warp)
  uint64_t* empty_a = full_a + kStages;          // [kStages]  MMA (tcgen05.commit) -> dequant warps
  uint64_t* full_b = empty_a + kStages;          // [kStagesB] TMA -> MMA
  uint64_t* empty_b = full_b + kStagesB;         // [kStagesB] MMA -> TMA
  uint64_t* full_w = empty_b + kStagesB;         // [NW] loader lanes (32 async arrivals) -> dequant warps
  uint64_t* empty_w = full_w + NW;               // [NW] dequant warps (one arrival per warp) -> loader
  uint64_t* accum_full = empty_w + NW;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile_n = blockIdx.x, tile_m = blockIdx.y;
  const int prow0 = tile_n * PR;
  const int m0 = tile_m * UN;
  const int num_kb = a.K / kBlockK;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < kStages; ++s) { mbar_init(&full_a[s], kDequantThreads / 32); mbar_init(&empty_a[s], 1); }
      for (int s = 0; s < kStagesB; ++s) { mbar_init(&full_b[s], 1); mbar_init(&empty_b[s], 1); }
      for (int s = 0; s < NW; ++s) { mbar_init(&full_w[s], 32); mbar_init(&empty_w[s], kDequantThreads / 32); }
      mbar_init(accum_full, 1);
      fence_barrier_init();
      HQQ_PREFETCH_TENSORMAP(&xmap);
    }
    __syncwarp();
    tmem_alloc<UN>(tmem_slot);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================= TMA producer: activation tiles =================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStagesB;
        mbar_wait(&empty_b[s], ((kb / kStagesB) & 1) ^ 1);
        mbar_expect_tx(&full_b[s], S::B_STAGE);  // both 256-token boxes
        tma_load_2d(sB + s * S::B_STAGE, &xmap, &full_b[s], kb * kBlockK, m0);
        tma_load_2d(sB + s * S::B_STAGE + S::B_HALF, &xmap, &full_b[s], kb * kBlockK, m0 + UNH);
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer (one elected thread) =================
    const uint32_t idesc = make_idesc<T>(UNH);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int sa = kb % kStages, sb = kb % kStagesB;
      mbar_wait(&full_a[sa], (kb / kStages) & 1);
      mbar_wait(&full_b[sb], (kb / kStagesB) & 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t adesc = make_desc_sw128(smem_u32(sA + sa * S::A_STAGE));
        const uint64_t bdesc0 = make_desc_sw128(smem_u32(sB + sb * S::B_STAGE));
        const uint64_t bdesc1 = make_desc_sw128(smem_u32(sB + sb * S::B_STAGE + S::B_HALF));
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)
          tc_mma_f16(tmem_base, adesc + (uint64_t)(k * 2), bdesc0 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
#pragma unroll
        for (int k = 0; k < kBlockK / 16; ++k)  // second accumulator: TMEM columns 256..511
          tc_mma_f16(tmem_base + (uint32_t)UNH, adesc + (uint64_t)(k * 2), bdesc1 + (uint64_t)(k * 2), idesc, (kb | k) != 0);
        tc_commit(&empty_a[sa]);
        tc_commit(&empty_b[sb]);
        if (kb == num_kb - 1) tc_commit(accum_full);
      }
      __syncwarp();
    }
  } else if (warp == 2) {
    // ================= loader: packed tile + scale/zero -> shared-memory rings (cp.async, no registers) =================
    const uint32_t sW_u32 = smem_u32(sW), sM_u32 = smem_u32(sM);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int sw = kb % NW;
      mbar_wait(&empty_w[sw], ((kb / NW) & 1) ^ 1);
#pragma unroll
      for (int i = 0; i < PR * 4 / 32; ++i) {  // PR rows x four 16-byte chunks, lanes along the row: coalesced 64-byte rows
        const int id = i * 32 + lane, row = id >> 2, ch = id & 3;
        const int prow = prow0 + row;
        const uint8_t* src = a.Wq + (long long)(prow < a.step ? prow : 0) * a.K + kb * kBlockK + ch * 16;
        cp_async_b<16>(sW_u32 + sw * S::W_STAGE + row * kBlockK + ch * 16, src);
      }
      if ((kb & 3) == 0) {  // the groups of k-blocks kb .. kb+3: GPQ values per row and array
        const int slot = (kb >> 2) % kMetaSlots;
#pragma unroll
        for (int i = 0; i < kTileRows * 2 / 32; ++i) {
          const int id = i * 32 + lane, arr = id >> 7, t = id & (kTileRows - 1);
          const int f = t / PR, prow = prow0 + t % PR;
          const long long mrow = (prow < a.step) ? (long long)f * a.step + prow : 0;
          const T* src = reinterpret_cast<const T*>(arr ? a.zero : a.scale) + mrow * a.Gk + (kb >> 2) * GPQ;
          cp_async_b<GPQ * 2>(sM_u32 + slot * S::M_SLOT + (arr * kTileRows + t) * (GPQ * 2), src);
        }
      }
      cp_async_mbar_arrive(&full_w[sw]);
    }
  } else {
    // ================= dequant warps: shared-memory packed bytes -> swizzled fp16/bf16 A tile =================
    static_assert(kStages == 4, "the dequant loop is unrolled over the 4 A stages");
    const int td = threadIdx.x - 96;
    const int pr = td / TPR, c = td % TPR;
    uint32_t soff[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
      const int row = f * PR + pr;
      if constexpr (BPT >= 8) soff[f] = (uint32_t)(row * 128) | ((uint32_t)(row & 7) << 16);
      else soff[f] = (uint32_t)(row * 128 + (((c >> 1) ^ (row & 7)) << 4) + (c & 1) * 8);
    }
    const uint32_t sA_u32 = smem_u32(sA);
    const int num_quads = num_kb >> 2;  // K % 256 == 0 (checked by the router)
    for (int q = 0; q < num_quads; ++q) {
      typename P2::T2 s2[4][F], z2[4][F];
#pragma unroll
      for (int d = 0; d < 4; ++d) {
        const int kb = 4 * q + d;
        const int sw = kb % NW;
        mbar_wait(&full_w[sw], (kb / NW) & 1);
        if (d == 0) {  // this quad's scale/zero arrived with its first k-block
          const uint8_t* slot = sM + (q % kMetaSlots) * S::M_SLOT;
#pragma unroll
          for (int f = 0; f < F; ++f) {
            const Vec<T, GPQ> sv = *reinterpret_cast<const Vec<T, GPQ>*>(slot + (0 * kTileRows + f * PR + pr) * (GPQ * 2));
            const Vec<T, GPQ> zv = *reinterpret_cast<const Vec<T, GPQ>*>(slot + (1 * kTileRows + f * PR + pr) * (GPQ * 2));
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) { s2[dd][f] = P2::bcast(sv.v[(dd * kBlockK) / GS]); z2[dd][f] = P2::bcast(zv.v[(dd * kBlockK) / GS]); }
          }
        }
        uint32_t wq[BPT / 4];
        {
          const uint8_t* p = sW + sw * S::W_STAGE + pr * kBlockK + c * BPT;
          if constexpr (BPT == 32) { const uint4 v0 = *reinterpret_cast<const uint4*>(p), v1 = *reinterpret_cast<const uint4*>(p + 16); wq[0] = v0.x; wq[1] = v0.y; wq[2] = v0.z; wq[3] = v0.w; wq[4] = v1.x; wq[5] = v1.y; wq[6] = v1.z; wq[7] = v1.w; }
          else if constexpr (BPT == 16) { const uint4 v = *reinterpret_cast<const uint4*>(p); wq[0] = v.x; wq[1] = v.y; wq[2] = v.z; wq[3] = v.w; }
          else if constexpr (BPT == 8) { const uint2 v = *reinterpret_cast<const uint2*>(p); wq[0] = v.x; wq[1] = v.y; }
          else { wq[0] = *reinterpret_cast<const uint32_t*>(p); }
        }
        mbar_wait(&empty_a[d], (uint32_t)(q & 1) ^ 1u);  // A stage index == d (four stages, four k-blocks per quad)
        const uint32_t stage = sA_u32 + d * S::A_STAGE;
#pragma unroll
        for (int f = 0; f < F; ++f) {
          const int sh = 8 - NBITS * (f + 1);
          uint32_t out[BPT / 2];
#pragma unroll
          for (int i = 0; i < BPT / 4; ++i) {
            const uint32_t t = (wq[i] >> sh) & (MASK * 0x01010101u);
            P2::deq4(t, z2[d][f], s2[d][f], out[2 * i], out[2 * i + 1]);
          }
          if constexpr (BPT >= 8) {
            const uint32_t rowbase = stage + (soff[f] & 0xFFFFu), rx = soff[f] >> 16;
#pragma unroll
            for (int ch = 0; ch < BPT / 8; ++ch) {
              const uint32_t addr = rowbase + (((uint32_t)(c * (BPT / 8) + ch) ^ rx) << 4);
              HQQ_STS_V4(addr, out[4 * ch], out[4 * ch + 1], out[4 * ch + 2], out[4 * ch + 3]);
            }
          } else {
            HQQ_STS_V2(stage + soff[f], out[0], out[1]);
          }
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&full_a[d]); mbar_arrive(&empty_w[sw]); }  // every lane's packed bytes (and meta) are in registers
      }
    }

    // ================= epilogue: TMEM -> registers -> y =================
    mbar_wait(accum_full, 0);
    tc_fence_after();
    const int quarter = warp & 3;                 // TMEM lanes this warp may touch: 32*quarter .. +31
    const int half = (warp - 3) >> 2;             // two warps share a quarter: split the token columns
    const int t = quarter * 32 + lane;
    const int tf = t / PR, tp = t % PR;
    const bool n_ok = (prow0 + tp) < a.step;
    const int n = tf * a.step + prow0 + tp;
    T* y = reinterpret_cast<T*>(a.y);
    const T* bias = reinterpret_cast<const T*>(a.bias);
    const bool has_bias = bias != nullptr;
    T bn = cvt_out<T>(0.0f);
    if (has_bias && n_ok) bn = bias[n];
#pragma unroll 1
    for (int col = half * (UN / 2); col < (half + 1) * (UN / 2); col += 32) {
      uint32_t v[32];
      tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)col, v);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const int m = m0 + col + j;
        if (n_ok && m < a.M) {
          T o = cvt_out<T>(__uint_as_float(v[j]));
          if (has_bias) o = __hadd(o, bn);
          y[(long long)m * a.N + n] = o;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc<UN>(tmem_base);
  }
}

// ---- host side ------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
#ifdef HQQ_EMU
  return &::emu::encode_tiled;
#endif
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// HQQ_B200_GEMM_VARIANT: "ld" = loader-warp kernel (linear_gemm_ld_kernel), "un512" = two accumulators per weight tile
// (linear_gemm_un512_kernel, M > 256 only), "ld512" = both (linear_gemm_ld512_kernel, M > 256 only), "dq16" = the default kernel with
// sixteen dequant warps (M > 128, not 1-bit), "un512dq" = un512 with sixteen dequant warps (M > 256; dq16 for 128 < M <= 256); all
// experimental
static int gemm_variant() {
  HQQ_ENV_KNOB(variant, ([] {
    const char* e = getenv("HQQ_B200_GEMM_VARIANT");
    return (e && !strcmp(e, "ld")) ? 1 : (e && !strcmp(e, "un512")) ? 2 : (e && !strcmp(e, "ld512")) ? 3 : (e && !strcmp(e, "dq16")) ? 4 : (e && !strcmp(e, "un512dq")) ? 5 : 0;
  })());
  return variant;
}

static int encode_xmap(CUtensorMap* xmap, const void* x, const Args& a, CUtensorMapDataType dt, size_t esize, int box_tokens) {
  EncodeTiledFn enc = get_encode();
  HQQ_REQUIRE(enc != nullptr, HQQ_E_CUDA, "hqq_b200_linear_fwd: cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.M};
  const cuuint64_t strides[1] = {(cuuint64_t)a.K * esize};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)box_tokens};
  const cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(xmap, dt, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HQQ_REQUIRE(r == CUDA_SUCCESS, HQQ_E_CUDA, "hqq_b200_linear_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
  return HQQ_OK;
}

template <typename T, int NBITS, int GS, int DQ = kDequantThreads>
static int launch_un512(const void* x, const Args& a, cudaStream_t st) {
  CUtensorMap xmap;
  const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int rc = encode_xmap(&xmap, x, a, dt, sizeof(T), Smem512::UNH);
  if (rc) return rc;
  constexpr int PR = kTileRows / (8 / NBITS);
  const dim3 grid((unsigned)cdiv(a.step, PR), (unsigned)cdiv(a.M, Smem512::UN));
  auto k = linear_gemm_un512_kernel<T, NBITS, GS, DQ>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem512::BYTES);
    HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: cannot reserve %d bytes of shared memory: %s", Smem512::BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  k<<<grid, 64 + DQ, Smem512::BYTES, st>>>(xmap, a);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/tcgen05-un512");
  return HQQ_OK;
}

// ---- split-K selection (opt-in) ---------------------------------------------------------------------------------------
static bool splitk_enabled() {
  HQQ_ENV_KNOB(on, ([] { const char* e = getenv("HQQ_B200_GEMM_SPLITK"); return (e && e[0] == '1') ? 1 : 0; })());
  return on == 1;
}
static int un_for(int64_t M) { return M <= 64 ? 64 : (M <= 128 ? 128 : 256); }
// k-slices per output tile: fill the 148 SMs when the tile grid alone cannot (at most 8 slices, at least one 256-k quad each)
static int splitk_factor(int64_t M, int64_t N, int64_t K, int nbits) {
  if (!splitk_enabled() || M > 1024) return 1;
  const int64_t tiles = cdiv(N, kTileRows) * cdiv(M, un_for(M));  // N % F == 0, so cdiv(step, PR) == cdiv(N, 128)
  (void)nbits;
  if (tiles * 2 > kNumSMs) return 1;
  int64_t S = kNumSMs / tiles;
  if (S > 8) S = 8;
  if (S > K / 256) S = K / 256;
  return S < 1 ? 1 : (int)S;
}
static size_t splitk_counter_bytes(int64_t M, int64_t N) { return (size_t)((cdiv(N, kTileRows) * cdiv(M, 64) * 4 + 255) & ~(int64_t)255); }

template <typename T, int NBITS, int GS>
static int launch_ld512(const void* x, const Args& a, cudaStream_t st) {
  CUtensorMap xmap;
  const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  using S = SmemLd512<NBITS>;
  int rc = encode_xmap(&xmap, x, a, dt, sizeof(T), S::UNH);
  if (rc) return rc;
  const dim3 grid((unsigned)cdiv(a.step, S::PR), (unsigned)cdiv(a.M, S::UN));
  auto k = linear_gemm_ld512_kernel<T, NBITS, GS>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, S::BYTES);
    HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: cannot reserve %d bytes of shared memory: %s", S::BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  k<<<grid, kLdThreads, S::BYTES, st>>>(xmap, a);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/tcgen05-ld512");
  return HQQ_OK;
}

template <typename T, int NBITS, int GS, int UN>
static int launch_splitk(const void* x, const Args& a0, int S, void* ws, size_t ws_bytes, cudaStream_t st) {
  const size_t cbytes = splitk_counter_bytes(a0.M, a0.N), need = cbytes + (size_t)S * a0.M * a0.N * sizeof(float);
  HQQ_REQUIRE(ws != nullptr && ws_bytes >= need && aligned(ws, 256), HQQ_E_WORKSPACE,
              "hqq_b200_linear_fwd: split-K needs a 256-byte aligned workspace of %zu bytes (got %zu)", need, ws_bytes);
  CUtensorMap xmap;
  const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  int rc = encode_xmap(&xmap, x, a0, dt, sizeof(T), UN);
  if (rc) return rc;
  ArgsSK a;
  static_cast<Args&>(a) = a0;
  a.counters = reinterpret_cast<unsigned*>(ws);
  a.ws = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + cbytes);
  cudaError_t e = cudaMemsetAsync(a.counters, 0, cbytes, st);  // the workspace is the caller's scratch: never assume it is clean
  HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: counter memset failed: %s", cudaGetErrorString(e));
  constexpr int PR = kTileRows / (8 / NBITS);
  const dim3 grid((unsigned)cdiv(a.step, PR), (unsigned)cdiv(a.M, UN), (unsigned)S);
  auto k = linear_gemm_splitk_kernel<T, NBITS, GS, UN>;
  static bool attr_set = false;
  if (!attr_set) {
    e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<UN>::BYTES);
    HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: cannot reserve %d bytes of shared memory: %s", Smem<UN>::BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  k<<<grid, kThreads, Smem<UN>::BYTES, st>>>(xmap, a);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/tcgen05-splitk");
  return HQQ_OK;
}

template <typename T, int NBITS, int GS, int UN, int DQ = kDequantThreads>
static int launch(const void* x, const Args& a, cudaStream_t st) {
  EncodeTiledFn enc = get_encode();
  HQQ_REQUIRE(enc != nullptr, HQQ_E_CUDA, "hqq_b200_linear_fwd: cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap xmap;
  const cuuint64_t dims[2] = {(cuuint64_t)a.K, (cuuint64_t)a.M};
  const cuuint64_t strides[1] = {(cuuint64_t)a.K * sizeof(T)};
  const cuuint32_t box[2] = {(cuuint32_t)kBlockK, (cuuint32_t)UN};
  const cuuint32_t estr[2] = {1, 1};
  const CUtensorMapDataType dt = std::is_same<T, __half>::value ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = enc(&xmap, dt, 2, const_cast<void*>(x), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  HQQ_REQUIRE(r == CUDA_SUCCESS, HQQ_E_CUDA, "hqq_b200_linear_fwd: cuTensorMapEncodeTiled failed (%d)", (int)r);
  constexpr int PR = kTileRows / (8 / NBITS);
  const dim3 grid((unsigned)cdiv(a.step, PR), (unsigned)cdiv(a.M, UN));
  const int variant = gemm_variant();
  if (variant == 1 && DQ == kDequantThreads) {
    auto kl = linear_gemm_ld_kernel<T, NBITS, GS, UN>;
    using SL = SmemLd<UN, NBITS>;
    static bool attr_set_ld = false;
    if (!attr_set_ld) {
      cudaError_t e = cudaFuncSetAttribute(kl, cudaFuncAttributeMaxDynamicSharedMemorySize, SL::BYTES);
      HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: cannot reserve %d bytes of shared memory: %s", SL::BYTES, cudaGetErrorString(e));
      attr_set_ld = true;
    }
    kl<<<grid, kLdThreads, SL::BYTES, st>>>(xmap, a);
    HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/tcgen05-ld");
    return HQQ_OK;
  }
  auto k = linear_gemm_kernel<T, NBITS, GS, UN, DQ>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Smem<UN>::BYTES);
    HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: cannot reserve %d bytes of shared memory: %s", Smem<UN>::BYTES, cudaGetErrorString(e));
    attr_set = true;
  }
  k<<<grid, 64 + DQ, Smem<UN>::BYTES, st>>>(xmap, a);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/tcgen05");
  return HQQ_OK;
}

template <typename T, int NBITS, int GS>
static int by_un(const void* x, const Args& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int S = splitk_factor(a.M, a.N, a.K, NBITS);
  if (S > 1) {
    if (a.M <= 64) return launch_splitk<T, NBITS, GS, 64>(x, a, S, ws, ws_bytes, st);
    if (a.M <= 128) return launch_splitk<T, NBITS, GS, 128>(x, a, S, ws, ws_bytes, st);
    return launch_splitk<T, NBITS, GS, 256>(x, a, S, ws, ws_bytes, st);
  }
  // HQQ_B200_GEMM_UN=128 (tuning knob): cap the token tile, e.g. to trade dequant work for wave efficiency
  HQQ_ENV_KNOB(un_cap, ([] { const char* e = getenv("HQQ_B200_GEMM_UN"); return e ? atoi(e) : 256; })());
  if (a.M > 256 && gemm_variant() == 2) return launch_un512<T, NBITS, GS>(x, a, st);
  if (a.M > 256 && gemm_variant() == 3) return launch_ld512<T, NBITS, GS>(x, a, st);
  if constexpr (NBITS != 1) {  // 1-bit: 16 packed rows per tile leave only two bytes per thread and k-block
    if (a.M > 256 && gemm_variant() == 5) return launch_un512<T, NBITS, GS, 512>(x, a, st);
    if (a.M > 128 && (gemm_variant() == 4 || gemm_variant() == 5) && un_cap > 128) return launch<T, NBITS, GS, 256, 512>(x, a, st);
  }
  if (a.M <= 64 || un_cap <= 64) return launch<T, NBITS, GS, 64>(x, a, st);
  if (a.M <= 128 || un_cap <= 128) return launch<T, NBITS, GS, 128>(x, a, st);
  return launch<T, NBITS, GS, 256>(x, a, st);
}

template <typename T, int NBITS>
static int by_gs(const void* x, const Args& a, int gs, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (gs == 64) return by_un<T, NBITS, 64>(x, a, ws, ws_bytes, st);
  return by_un<T, NBITS, 128>(x, a, ws, ws_bytes, st);
}

template <typename T>
static int by_bits(const void* x, const Args& a, int gs, int nbits, void* ws, size_t ws_bytes, cudaStream_t st) {
  switch (nbits) {
    case 8: return by_gs<T, 8>(x, a, gs, ws, ws_bytes, st);
    case 4: return by_gs<T, 4>(x, a, gs, ws, ws_bytes, st);
    case 2: return by_gs<T, 2>(x, a, gs, ws, ws_bytes, st);
    case 1: return by_gs<T, 1>(x, a, gs, ws, ws_bytes, st);
  }
  return HQQ_E_UNSUPPORTED;
}

}  // namespace gemm

bool gemm_route_ok(int64_t M, int64_t N, int64_t K, int gs, int nbits, int axis, int dtype) {
  if (axis != 1) return false;
  if (dtype != HQQ_F16 && dtype != HQQ_BF16) return false;
  if (!(nbits == 8 || nbits == 4 || nbits == 2 || nbits == 1)) return false;
  if (!(gs == 64 || gs == 128)) return false;     // one 64-k stage never straddles a group
  if (M < 1 || K % 256 != 0 || K % gs != 0) return false;  // the dequant loop handles four 64-k stages per iteration
  if (N % (8 / nbits) != 0) return false;
  if (K % 8 != 0 || N > (1 << 28) || K > (1 << 28) || M > (1 << 28)) return false;
  return true;
}

size_t gemm_workspace_bytes(int64_t M, int64_t N, int64_t K, int, int nbits, int) {
  const int S = gemm::splitk_factor(M, N, K, nbits);
  return S > 1 ? gemm::splitk_counter_bytes(M, N) + (size_t)S * M * N * sizeof(float) : 0;
}

int linear_gemm(const void* x, const void* Wq, const void* scale, const void* zero, const void* bias, void* y, int64_t M, int64_t N,
                int64_t K, int gs, int nbits, int dtype, void* ws, size_t ws_bytes, cudaStream_t st) {
  HQQ_REQUIRE(aligned(x, 16) && aligned(Wq, 16), HQQ_E_INVALID, "hqq_b200_linear_fwd: x and W_q must be 16-byte aligned");
  gemm::Args a;
  a.Wq = (const uint8_t*)Wq; a.scale = scale; a.zero = zero; a.bias = bias; a.y = y;
  a.M = (int)M; a.N = (int)N; a.K = (int)K;
  a.step = (int)(N / (8 / nbits));
  a.Gk = (int)(K / gs);
  if (dtype == HQQ_F16) return gemm::by_bits<__half>(x, a, gs, nbits, ws, ws_bytes, st);
  return gemm::by_bits<__nv_bfloat16>(x, a, gs, nbits, ws, ws_bytes, st);
}

}  // namespace hqq
