// HQQLinear.forward for small M (decode): fused unpack -> group-dequant -> MMA, weight-streaming (HBM-bound).
//
// y[M,N] = x[M,K] @ dequantize(W_q)^T (+bias)        reference: hqq/core/quantize.py:880-898
//
// The packed tensor keeps the reference's slab layout (bitpack.py): for axis=1 a byte at packed row p,
// column k holds the levels of output rows p + f*(N/F), f = 0..F-1 (F = 8/nbits).  A warp owns a 16-row
// MMA tile made of P = 16/F packed rows x F slabs and streams them along K with 16-byte loads straight
// into registers (each weight byte is read exactly once, L1::no_allocate).  The levels are NOT dequantised
// per element: with per-group scale s and zero z
//        sum_k x_k (q_k - z) s  =  s * (sum_k q_k x_k)  -  s z * (sum_k x_k)
// so the tensor core contracts the raw levels (bit-tricked into fp16/bf16 lanes, 6-9 ALU ops per 8 weights)
// against x, a second MMA with an all-ones A tile yields sum_k x_k in the same fragment layout, and the
// affine correction is applied once per group per accumulator (fp32).  mma.sync m16n8k16 with register A
// fragments is used on purpose: at M <= 32 the kernel is bound by HBM and instruction issue, and a
// register-operand MMA avoids the shared-memory round trip a tcgen05 operand would need.
//
// Scheduling is stream-K: the (16-row tile) x (256-k unit) space of up to four weight matrices that share the
// activation (q/k/v, gate/up) is one linear sequence split evenly over all resident warps.  Every warp streams its
// contiguous slice through a private cp.async ring in shared memory (3 units in flight per warp, no register
// staging, no block-level barriers).  A tile that is split across warps is finished by the warp holding its first
// k-chunk, which adds the other warps' partials in ascending warp order: deterministic, no atomics on the data.
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "common.cuh"

namespace hqq {

constexpr int kMaxProb = 4;   // weight matrices sharing one activation in a single launch (q/k/v, gate/up)

struct SKProb {
  const uint8_t* Wq;
  const void* scale;
  const void* zero;
  const void* bias;
  void* y;
  int N;
  int step;   // packed rows = N / F
  int tile0;  // first global 16-row tile of this matrix
};

struct SKArgs {
  SKProb p[kMaxProb];
  int nprob;
  const void* x;
  int M, K;
  int Gk;           // groups per output row = K / GS
  int KB;           // 256-k units per row tile = K / 256
  int total_tiles;
  long long total_units;
  float* ws_partial;  // [warps][MT][128] split-K partial tiles in fragment layout
  int* ws_flags;      // [warps] 0 = empty, 1 = partial ready (left zeroed on exit)
};

template <typename T> struct MT16;
template <> struct MT16<__half> {
  static constexpr uint32_t ONE2 = 0x3C003C00u;
  __device__ __forceinline__ static void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  // same with C = 0 (first MMA of a group): no accumulator clearing instructions needed
  __device__ __forceinline__ static void mma0(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.0f));
  }
  __device__ __forceinline__ static float ld(const void* p, long long i) { return __half2float(reinterpret_cast<const __half*>(p)[i]); }
  __device__ __forceinline__ static void st(void* p, long long i, float v, const void* bias, int n) {
    __half o = __float2half_rn(v);
    if (bias) o = __hadd(o, reinterpret_cast<const __half*>(bias)[n]);  // out += bias, second rounding as in the reference
    reinterpret_cast<__half*>(p)[i] = o;
  }
};
template <> struct MT16<__nv_bfloat16> {
  static constexpr uint32_t ONE2 = 0x3F803F80u;
  __device__ __forceinline__ static void mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  __device__ __forceinline__ static void mma0(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%10,%10,%10,%10};"
                 : "=f"(d[0]), "=f"(d[1]), "=f"(d[2]), "=f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "f"(0.0f));
  }
  __device__ __forceinline__ static float ld(const void* p, long long i) { return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p)[i]); }
  __device__ __forceinline__ static void st(void* p, long long i, float v, const void* bias, int n) {
    __nv_bfloat16 o = __float2bfloat16_rn(v);
    if (bias) o = __hadd(o, reinterpret_cast<const __nv_bfloat16*>(bias)[n]);
    reinterpret_cast<__nv_bfloat16*>(p)[i] = o;
  }
};

__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t s) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
  return r;
}
template <int LUT>
__device__ __forceinline__ uint32_t lop3(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r;
  asm("lop3.b32 %0, %1, %2, %3, %4;" : "=r"(r) : "r"(a), "r"(b), "r"(c), "n"(LUT));
  return r;
}
// (a & b) | c
__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t b, uint32_t c) { return lop3<0xEA>(a, b, c); }

// How the integer levels are planted into 16-bit float lanes (lane value = OFF + q * V):
//   MAGIC_OFFSET    fp16: bits | 0x6400 -> 1024 + q*2^sh          bf16: (bits >> sh) | 0x4300 -> 128 + q
//   MAGIC_SUBNORMAL fp16 only: bits taken as a subnormal          -> q * 2^(sh-24)   (no offset, exact)
enum { MAGIC_OFFSET = 0, MAGIC_SUBNORMAL = 1 };

template <typename T, int NBITS, int MAGIC> struct Lanes;

// fp16, sub-byte fields: mask in place, no shift (1 PRMT + 1 SHF + 4 LOP3 per 8 weights)
template <int NBITS, int MAGIC>
struct Lanes<__half, NBITS, MAGIC> {
  static constexpr uint32_t OR = (MAGIC == MAGIC_OFFSET) ? 0x64006400u : 0u;
  uint32_t mask_a, mask_b;
  float invV_a, invV_b, offV_a, offV_b;
  __device__ __forceinline__ void init(int sh_a, int sh_b) {
    const uint32_t m = (1u << NBITS) - 1u;
    mask_a = (m << sh_a) * 0x00010001u;
    mask_b = (m << sh_b) * 0x00010001u;
    if (MAGIC == MAGIC_OFFSET) {
      invV_a = exp2f(-(float)sh_a); invV_b = exp2f(-(float)sh_b);
      offV_a = 1024.0f * invV_a;    offV_b = 1024.0f * invV_b;
    } else {
      invV_a = exp2f(24.0f - (float)sh_a); invV_b = exp2f(24.0f - (float)sh_b);
      offV_a = 0.0f; offV_b = 0.0f;
    }
  }
  // w: 4 consecutive k-bytes of one packed row.  a0/a2: field A for k{0,1} / k{2,3}; a1/a3: field B.
  __device__ __forceinline__ void extract(uint32_t w, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) const {
    const uint32_t wp = prmt(w, 0u, 0x3120u);  // bytes [k0,k2,k1,k3]: 16-bit lanes {k0|k2<<8, k1|k3<<8}
    const uint32_t wh = wp >> 8;
    a0 = and_or(wp, mask_a, OR);
    a1 = and_or(wp, mask_b, OR);
    a2 = and_or(wh, mask_a, OR);
    a3 = and_or(wh, mask_b, OR);
  }
};

// bf16, sub-byte fields: only 7 mantissa bits -> shift the field down to bit 0 first
template <int NBITS, int MAGIC>
struct Lanes<__nv_bfloat16, NBITS, MAGIC> {
  int sh_a, sh_b;
  float invV_a, invV_b, offV_a, offV_b;
  __device__ __forceinline__ void init(int sa, int sb) {
    sh_a = sa; sh_b = sb;
    invV_a = invV_b = 1.0f;
    offV_a = offV_b = 128.0f;
  }
  __device__ __forceinline__ void extract(uint32_t w, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) const {
    constexpr uint32_t M = ((1u << NBITS) - 1u) * 0x00010001u;
    const uint32_t wp = prmt(w, 0u, 0x3120u);
    a0 = and_or(wp >> sh_a, M, 0x43004300u);
    a1 = and_or(wp >> sh_b, M, 0x43004300u);
    a2 = and_or(wp >> (sh_a + 8), M, 0x43004300u);
    a3 = and_or(wp >> (sh_b + 8), M, 0x43004300u);
  }
};

// fp16, 8-bit: whole bytes, two packed rows per thread (rows r and r+8 of the tile)
template <int MAGIC>
struct Lanes<__half, 8, MAGIC> {
  static constexpr uint32_t HB = (MAGIC == MAGIC_OFFSET) ? 0x64646464u : 0u;
  float invV_a, invV_b, offV_a, offV_b;
  __device__ __forceinline__ void init(int, int) {
    if (MAGIC == MAGIC_OFFSET) { invV_a = invV_b = 1.0f; offV_a = offV_b = 1024.0f; }
    else { invV_a = invV_b = 16777216.0f; offV_a = offV_b = 0.0f; }
  }
  __device__ __forceinline__ void extract2(uint32_t wa, uint32_t wb, uint32_t& a0, uint32_t& a1, uint32_t& a2, uint32_t& a3) const {
    a0 = prmt(wa, HB, 0x4140u);  // lanes {k0, k1} of row r
    a2 = prmt(wa, HB, 0x4342u);  // lanes {k2, k3}
    a1 = prmt(wb, HB, 0x4140u);  // row r+8
    a3 = prmt(wb, HB, 0x4342u);
  }
};

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(g) : "memory");
}
template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* smem, const void* g) {
  const uint32_t s = (uint32_t)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(s), "l"(g), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <typename T, int NBITS, int GS, int MT, int MAGIC>
struct SKCfg {
  static constexpr int F = 8 / NBITS;            // fields (slabs) per byte
  static constexpr int P = 16 / F;               // packed rows per 16-row MMA tile
  static constexpr int MPG = GS / 16;            // MMAs per quantisation group
  static constexpr int GPB = 256 / GS;           // quantisation groups per 256-k unit
  static constexpr int MB = GPB * 2;             // bytes of scale (or zero) per unit and row
  static constexpr int NWV = (F == 1) ? 8 : 4;   // 16-byte weight vectors per thread and unit
  static constexpr int ST = (F == 1) ? 2 : 4;    // ring stages
  static constexpr int W_BYTES = ST * NWV * 256 * 16;
  static constexpr int M_BYTES = ST * 4 * 256 * MB;
  static constexpr int SMEM = W_BYTES + M_BYTES;
};

template <typename T, int NBITS, int GS, int MT, int MAGIC>
__global__ void __launch_bounds__(256, 2) linear_streamk_kernel(const __grid_constant__ SKArgs a) {
  using C = SKCfg<T, NBITS, GS, MT, MAGIC>;
  constexpr int F = C::F, P = C::P, MPG = C::MPG, GPB = C::GPB, MB = C::MB, NWV = C::NWV, ST = C::ST;
  using MM = MT16<T>;
  extern __shared__ __align__(16) uint8_t smem[];
  uint4* wring = reinterpret_cast<uint4*>(smem);   // [ST][NWV][256] one 16-byte slot per thread
  uint8_t* mring = smem + C::W_BYTES;              // [ST][4][256][MB]

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int r = lane >> 2, c = lane & 3;
  const long long TW = (long long)gridDim.x * 8, gw = (long long)blockIdx.x * 8 + warp;
  const long long u0 = gw * a.total_units / TW, u1 = (gw + 1) * a.total_units / TW;
  if (u0 >= u1) return;  // no block-level synchronisation anywhere below: warps are independent

  const int p = (F == 1) ? r : (r % P);
  const int fa = (F == 1) ? 0 : (r / P), fb = (F == 1) ? 0 : (F / 2 + r / P);
  Lanes<T, NBITS, MAGIC> lanes;
  lanes.init(8 - NBITS * (fa + 1), 8 - NBITS * (fb + 1));

  // ---- tile lookup: which matrix a global tile belongs to, and this thread's two rows in it ----------------
  struct Tile {
    const uint8_t* Wq; const T* scale; const T* zero; const T* bias; T* y;
    int N, prow_a, prow_b, n_a, n_b; bool ok_a, ok_b;
  };
  auto locate = [&](int gt, Tile& t) {
    int pi = 0;
#pragma unroll
    for (int i = 1; i < kMaxProb; ++i)
      if (i < a.nprob && gt >= a.p[i].tile0) pi = i;
    const uint8_t* Wq = a.p[0].Wq; const void* sc = a.p[0].scale; const void* ze = a.p[0].zero; const void* bi = a.p[0].bias;
    void* y = a.p[0].y; int N = a.p[0].N, step = a.p[0].step, tile0 = a.p[0].tile0;
#pragma unroll
    for (int i = 1; i < kMaxProb; ++i)
      if (pi == i) { Wq = a.p[i].Wq; sc = a.p[i].scale; ze = a.p[i].zero; bi = a.p[i].bias; y = a.p[i].y; N = a.p[i].N; step = a.p[i].step; tile0 = a.p[i].tile0; }
    t.Wq = Wq; t.scale = reinterpret_cast<const T*>(sc); t.zero = reinterpret_cast<const T*>(ze);
    t.bias = reinterpret_cast<const T*>(bi); t.y = reinterpret_cast<T*>(y); t.N = N;
    t.prow_a = (gt - tile0) * P + p;
    t.prow_b = (F == 1) ? t.prow_a + 8 : t.prow_a;
    t.ok_a = t.prow_a < step; t.ok_b = t.prow_b < step;
    t.n_a = fa * step + t.prow_a; t.n_b = fb * step + t.prow_b;
  };

  // ---- issue cursor: where the next cp.async unit comes from -------------------------------------------------
  int i_gt = (int)(u0 / a.KB), i_kb = (int)(u0 % a.KB);
  const uint8_t *iw_a, *iw_b;
  const T *is_a, *iz_a, *is_b, *iz_b;
  auto issue_setup = [&]() {
    Tile t; locate(i_gt, t);
    // rows past the ragged edge re-read row 0 (always mapped); their results are never stored
    const long long ra = t.ok_a ? t.prow_a : 0, rb = t.ok_b ? t.prow_b : 0;
    const long long na = t.ok_a ? t.n_a : 0, nb = t.ok_b ? t.n_b : 0;
    iw_a = t.Wq + ra * a.K + (long long)i_kb * 256 + 16 * c;
    iw_b = t.Wq + rb * a.K + (long long)i_kb * 256 + 16 * c;
    is_a = t.scale + na * a.Gk + i_kb * GPB; iz_a = t.zero + na * a.Gk + i_kb * GPB;
    is_b = t.scale + nb * a.Gk + i_kb * GPB; iz_b = t.zero + nb * a.Gk + i_kb * GPB;
  };
  issue_setup();
  long long issued = u0;
  auto issue = [&](int stage) {
    if (issued < u1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) cp_async16(&wring[(stage * NWV + i) * 256 + tid], iw_a + i * 64);
      if (F == 1) {
#pragma unroll
        for (int i = 0; i < 4; ++i) cp_async16(&wring[(stage * NWV + 4 + i) * 256 + tid], iw_b + i * 64);
      }
      cp_async_small<MB>(mring + ((stage * 4 + 0) * 256 + tid) * MB, is_a);
      cp_async_small<MB>(mring + ((stage * 4 + 1) * 256 + tid) * MB, iz_a);
      cp_async_small<MB>(mring + ((stage * 4 + 2) * 256 + tid) * MB, is_b);
      cp_async_small<MB>(mring + ((stage * 4 + 3) * 256 + tid) * MB, iz_b);
      ++issued;
      if (++i_kb == a.KB) {
        i_kb = 0; ++i_gt;
        if (issued < u1) issue_setup();
      } else {
        iw_a += 256; iw_b += 256; is_a += GPB; iz_a += GPB; is_b += GPB; iz_b += GPB;
      }
    }
    cp_async_commit();  // always commit (possibly empty) so the group count per iteration is uniform
  };
#pragma unroll
  for (int s = 0; s < ST - 1; ++s) issue(s);

  // ---- consume cursor ----------------------------------------------------------------------------------------
  int c_gt = (int)(u0 / a.KB), c_kb = (int)(u0 % a.KB);
  bool new_tile = true;
  int k_first = 0;
  Tile ct;
  const T* xp[MT];
  float tot[MT][4];
  int stage = 0;

  for (long long u = u0; u < u1; ++u) {
    {
      int is = stage + (ST - 1);
      if (is >= ST) is -= ST;
      issue(is);
    }
    if (new_tile) {
      locate(c_gt, ct);
      k_first = c_kb;
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        // token columns >= M alias the last real token: MMA columns are independent and never stored
        const int m = min(mt * 8 + r, a.M - 1);
        xp[mt] = reinterpret_cast<const T*>(a.x) + (long long)m * a.K + 16 * c;
#pragma unroll
        for (int i = 0; i < 4; ++i) tot[mt][i] = 0.0f;
      }
      new_tile = false;
    }
    cp_async_wait<ST - 1>();  // the group of unit u (and everything older) has landed in this thread's slots

    float sA[GPB], zA[GPB], sB[GPB], zB[GPB];
    {
      const Vec<T, GPB> v0 = *reinterpret_cast<const Vec<T, GPB>*>(mring + ((stage * 4 + 0) * 256 + tid) * MB);
      const Vec<T, GPB> v1 = *reinterpret_cast<const Vec<T, GPB>*>(mring + ((stage * 4 + 1) * 256 + tid) * MB);
      const Vec<T, GPB> v2 = *reinterpret_cast<const Vec<T, GPB>*>(mring + ((stage * 4 + 2) * 256 + tid) * MB);
      const Vec<T, GPB> v3 = *reinterpret_cast<const Vec<T, GPB>*>(mring + ((stage * 4 + 3) * 256 + tid) * MB);
#pragma unroll
      for (int i = 0; i < GPB; ++i) { sA[i] = to_f32<T>(v0.v[i]); zA[i] = to_f32<T>(v1.v[i]); sB[i] = to_f32<T>(v2.v[i]); zB[i] = to_f32<T>(v3.v[i]); }
    }
    float Sg[MT][4], Xg[MT][4];
#pragma unroll
    for (int us = 0; us < 4; ++us) {
      // activations for this k64 step: 16 consecutive k per thread, one column (token) per 4-lane group
      uint4 xa[MT], xb[MT];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        const uint4* q = reinterpret_cast<const uint4*>(xp[mt] + (long long)c_kb * 256 + us * 64);
        xa[mt] = __ldg(q);
        xb[mt] = __ldg(q + 1);
      }
      const uint4 va = wring[(stage * NWV + us) * 256 + tid];
      uint4 vb = va;
      if (F == 1) vb = wring[(stage * NWV + 4 + us) * 256 + tid];
      const uint32_t wa[4] = {va.x, va.y, va.z, va.w};
      const uint32_t wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t a0, a1, a2, a3;
        if constexpr (F == 1) lanes.extract2(wa[j], wb[j], a0, a1, a2, a3);
        else lanes.extract(wa[j], a0, a1, a2, a3);
        const bool first = ((us * 4 + j) % MPG) == 0;  // first MMA of a group starts from C = 0
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t b0 = (j == 0) ? xa[mt].x : (j == 1) ? xa[mt].z : (j == 2) ? xb[mt].x : xb[mt].z;
          const uint32_t b1 = (j == 0) ? xa[mt].y : (j == 1) ? xa[mt].w : (j == 2) ? xb[mt].y : xb[mt].w;
          if (first) {
            MM::mma0(Sg[mt], a0, a1, a2, a3, b0, b1);
            MM::mma0(Xg[mt], MM::ONE2, MM::ONE2, MM::ONE2, MM::ONE2, b0, b1);
          } else {
            MM::mma(Sg[mt], a0, a1, a2, a3, b0, b1);
            MM::mma(Xg[mt], MM::ONE2, MM::ONE2, MM::ONE2, MM::ONE2, b0, b1);
          }
        }
        if (((us * 4 + j + 1) % MPG) == 0) {
          // a quantisation group is complete: tot += s*(Q - z*X), with lane value = OFF + q*V folded in
          const int gi = (us * 4 + j) / MPG;
          const float ka = sA[gi] * lanes.invV_a, la = -sA[gi] * (lanes.offV_a + zA[gi]);
          const float kb = sB[gi] * lanes.invV_b, lb = -sB[gi] * (lanes.offV_b + zB[gi]);
#pragma unroll
          for (int mt = 0; mt < MT; ++mt) {
            tot[mt][0] = fmaf(ka, Sg[mt][0], fmaf(la, Xg[mt][0], tot[mt][0]));
            tot[mt][1] = fmaf(ka, Sg[mt][1], fmaf(la, Xg[mt][1], tot[mt][1]));
            tot[mt][2] = fmaf(kb, Sg[mt][2], fmaf(lb, Xg[mt][0], tot[mt][2]));
            tot[mt][3] = fmaf(kb, Sg[mt][3], fmaf(lb, Xg[mt][1], tot[mt][3]));
          }
        }
      }
    }

    // ---- end of unit: close the tile if this was its last unit in our slice -----------------------------------
    const bool tile_end = (c_kb == a.KB - 1);
    if (tile_end || u == u1 - 1) {
      const bool covers_start = (k_first == 0);
      if (!covers_start) {
        // contributor: the tile began in an earlier warp's slice -> publish our partial (fragment layout)
        float* dst = a.ws_partial + (gw * MT) * 128 + lane * 4;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
          *reinterpret_cast<float4*>(dst + mt * 128) = make_float4(tot[mt][0], tot[mt][1], tot[mt][2], tot[mt][3]);
        __threadfence();
        __syncwarp();
        if (lane == 0) st_release(a.ws_flags + gw, 1);
      } else {
        if (!tile_end) {
          // finisher: we hold the first k-chunk; later warps hold the rest.  Add their partials in warp order.
          const long long tile_end_unit = ((long long)c_gt + 1) * a.KB;
          for (long long w2 = gw + 1; w2 < TW; ++w2) {
            const long long s2 = w2 * a.total_units / TW;
            if (s2 >= tile_end_unit) break;
            if (s2 >= (w2 + 1) * a.total_units / TW) continue;  // that warp has no units
            while (ld_acquire(a.ws_flags + w2) == 0) {}
            const float* src = a.ws_partial + (w2 * MT) * 128 + lane * 4;
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
              const float4 v = __ldcg(reinterpret_cast<const float4*>(src + mt * 128));
              tot[mt][0] += v.x; tot[mt][1] += v.y; tot[mt][2] += v.z; tot[mt][3] += v.w;
            }
            __syncwarp();
            if (lane == 0) a.ws_flags[w2] = 0;  // leave the workspace clean for the next launch / graph replay
          }
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const int m0 = mt * 8 + 2 * c;
          if (m0 < a.M) {
            if (ct.ok_a) MM::st(ct.y, (long long)m0 * ct.N + ct.n_a, tot[mt][0], ct.bias, ct.n_a);
            if (ct.ok_b) MM::st(ct.y, (long long)m0 * ct.N + ct.n_b, tot[mt][2], ct.bias, ct.n_b);
          }
          if (m0 + 1 < a.M) {
            if (ct.ok_a) MM::st(ct.y, (long long)(m0 + 1) * ct.N + ct.n_a, tot[mt][1], ct.bias, ct.n_a);
            if (ct.ok_b) MM::st(ct.y, (long long)(m0 + 1) * ct.N + ct.n_b, tot[mt][3], ct.bias, ct.n_b);
          }
        }
      }
    }
    if (tile_end) { c_kb = 0; ++c_gt; new_tile = true; } else { ++c_kb; }
    if (++stage == ST) stage = 0;
  }
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------------------------------
static int magic_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("HQQ_B200_GEMV_MAGIC");
    mode = (e && !strcmp(e, "subnormal")) ? MAGIC_SUBNORMAL : MAGIC_OFFSET;
  }
  return mode;
}

static int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = kNumSMs;
  }
  return n;
}

// The persistent grid: every CTA must be co-resident (split tiles are finished by spinning on peer warps).
template <typename T, int NBITS, int GS, int MT, int MAGIC>
static int grid_for_kernel(int* grid_out) {
  using C = SKCfg<T, NBITS, GS, MT, MAGIC>;
  static int grid = 0;
  if (!grid) {
    auto k = linear_streamk_kernel<T, NBITS, GS, MT, MAGIC>;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
    HQQ_REQUIRE(e == cudaSuccess, HQQ_E_CUDA, "hqq_b200_linear_fwd: cannot reserve %d bytes of shared memory: %s", C::SMEM, cudaGetErrorString(e));
    int occ = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, 256, C::SMEM);
    HQQ_REQUIRE(e == cudaSuccess && occ > 0, HQQ_E_CUDA, "hqq_b200_linear_fwd: occupancy query failed: %s", cudaGetErrorString(e));
    if (occ > 2) occ = 2;
    grid = sm_count() * occ;
  }
  *grid_out = grid;
  return HQQ_OK;
}

constexpr int kMaxGrid = 2 * 160;  // upper bound used to size the workspace before the kernel variant is known

template <typename T, int NBITS, int GS, int MT, int MAGIC>
static int launch_sk(SKArgs& a, void* ws, size_t ws_bytes, cudaStream_t st) {
  using C = SKCfg<T, NBITS, GS, MT, MAGIC>;
  int grid = 0;
  int rc = grid_for_kernel<T, NBITS, GS, MT, MAGIC>(&grid);
  if (rc) return rc;
  const size_t need = (size_t)grid * 8 * MT * 128 * sizeof(float) + (size_t)kMaxGrid * 8 * sizeof(int);
  HQQ_REQUIRE(ws && ws_bytes >= need, HQQ_E_WORKSPACE, "hqq_b200_linear_fwd: workspace %zu < required %zu bytes", ws_bytes, need);
  a.ws_flags = reinterpret_cast<int*>(ws);  // flags first: they must stay zero between launches
  a.ws_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + (size_t)kMaxGrid * 8 * sizeof(int));
  linear_streamk_kernel<T, NBITS, GS, MT, MAGIC><<<grid, 256, C::SMEM, st>>>(a);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/streamk");
  return HQQ_OK;
}

template <typename T, int NBITS, int GS, int MAGIC>
static int sk_mt(SKArgs& a, void* ws, size_t wsb, cudaStream_t st) {
  if (a.M <= 8) return launch_sk<T, NBITS, GS, 1, MAGIC>(a, ws, wsb, st);
  if (a.M <= 16) return launch_sk<T, NBITS, GS, 2, MAGIC>(a, ws, wsb, st);
  return launch_sk<T, NBITS, GS, 4, MAGIC>(a, ws, wsb, st);
}

template <typename T, int NBITS, int MAGIC>
static int sk_gs(SKArgs& a, int gs, void* ws, size_t wsb, cudaStream_t st) {
  switch (gs) {
    case 64: return sk_mt<T, NBITS, 64, MAGIC>(a, ws, wsb, st);
    case 128: return sk_mt<T, NBITS, 128, MAGIC>(a, ws, wsb, st);
  }
  return HQQ_E_UNSUPPORTED;
}

template <typename T>
static int sk_bits(SKArgs& a, int gs, int nbits, void* ws, size_t wsb, cudaStream_t st) {
  const bool sub = std::is_same<T, __half>::value && magic_mode() == MAGIC_SUBNORMAL;
  switch (nbits) {
    case 8:
      if constexpr (std::is_same<T, __half>::value) return sub ? sk_gs<T, 8, MAGIC_SUBNORMAL>(a, gs, ws, wsb, st) : sk_gs<T, 8, MAGIC_OFFSET>(a, gs, ws, wsb, st);
      else return HQQ_E_UNSUPPORTED;
    case 4:
      if constexpr (std::is_same<T, __half>::value) return sub ? sk_gs<T, 4, MAGIC_SUBNORMAL>(a, gs, ws, wsb, st) : sk_gs<T, 4, MAGIC_OFFSET>(a, gs, ws, wsb, st);
      else return sk_gs<T, 4, MAGIC_OFFSET>(a, gs, ws, wsb, st);
    case 2:
      if constexpr (std::is_same<T, __half>::value) return sub ? sk_gs<T, 2, MAGIC_SUBNORMAL>(a, gs, ws, wsb, st) : sk_gs<T, 2, MAGIC_OFFSET>(a, gs, ws, wsb, st);
      else return sk_gs<T, 2, MAGIC_OFFSET>(a, gs, ws, wsb, st);
    case 1:
      if constexpr (std::is_same<T, __half>::value) return sub ? sk_gs<T, 1, MAGIC_SUBNORMAL>(a, gs, ws, wsb, st) : sk_gs<T, 1, MAGIC_OFFSET>(a, gs, ws, wsb, st);
      else return sk_gs<T, 1, MAGIC_OFFSET>(a, gs, ws, wsb, st);
  }
  return HQQ_E_UNSUPPORTED;
}

bool small_route_ok(int64_t M, int64_t N, int64_t K, int gs, int nbits, int axis, int dtype) {
  if (axis != 1) return false;
  if (dtype != HQQ_F16 && dtype != HQQ_BF16) return false;
  if (!(nbits == 8 || nbits == 4 || nbits == 2 || nbits == 1)) return false;
  if (nbits == 8 && dtype == HQQ_BF16) return false;  // 8-bit levels do not fit a bf16 mantissa trick
  if (!(gs == 64 || gs == 128)) return false;         // a 64-k MMA step must not straddle groups; meta is staged 4/8 bytes at a time
  if (M < 1 || M > 32) return false;
  if (K % 256 != 0 || K % gs != 0) return false;      // 256-k units; groups never straddle a row
  if (N % (8 / nbits) != 0) return false;
  if (N > (1 << 28) || K > (1 << 28)) return false;
  return true;
}

size_t small_workspace_bytes(int64_t M) {
  const int MT = M <= 8 ? 1 : (M <= 16 ? 2 : 4);
  return (size_t)kMaxGrid * 8 * sizeof(int) + (size_t)kMaxGrid * 8 * MT * 128 * sizeof(float);
}

int linear_small_multi(const void* x, int nprob, const void* const* Wq, const void* const* scale, const void* const* zero,
                       const void* const* bias, void* const* y, const int64_t* N, int64_t M, int64_t K, int gs, int nbits, int dtype,
                       void* ws, size_t ws_bytes, cudaStream_t st) {
  HQQ_REQUIRE(nprob >= 1 && nprob <= kMaxProb, HQQ_E_INVALID, "hqq_b200_linear_fwd_multi: 1..%d matrices per launch (got %d)", kMaxProb, nprob);
  HQQ_REQUIRE(aligned(x, 16), HQQ_E_INVALID, "hqq_b200_linear_fwd: x must be 16-byte aligned");
  HQQ_REQUIRE(aligned(ws, 16), HQQ_E_INVALID, "hqq_b200_linear_fwd: workspace must be 16-byte aligned");
  const int F = 8 / nbits, P = 16 / F;
  SKArgs a;
  a.nprob = nprob; a.x = x; a.M = (int)M; a.K = (int)K; a.Gk = (int)(K / gs); a.KB = (int)(K / 256);
  int tiles = 0;
  for (int i = 0; i < kMaxProb; ++i) {
    const int j = i < nprob ? i : 0;
    HQQ_REQUIRE(Wq[j] && scale[j] && zero[j] && y[j], HQQ_E_INVALID, "hqq_b200_linear_fwd: null pointer");
    HQQ_REQUIRE(aligned(Wq[j], 16) && aligned(scale[j], 8) && aligned(zero[j], 8), HQQ_E_INVALID,
                "hqq_b200_linear_fwd: W_q must be 16-byte and scale/zero 8-byte aligned");
    a.p[i].Wq = (const uint8_t*)Wq[j]; a.p[i].scale = scale[j]; a.p[i].zero = zero[j]; a.p[i].bias = bias ? bias[j] : nullptr;
    a.p[i].y = y[j]; a.p[i].N = (int)N[j]; a.p[i].step = (int)(N[j] / F); a.p[i].tile0 = tiles;
    if (i < nprob) tiles += (int)cdiv(a.p[i].step, P);
  }
  a.total_tiles = tiles;
  a.total_units = (long long)tiles * a.KB;
  if (dtype == HQQ_F16) return sk_bits<__half>(a, gs, nbits, ws, ws_bytes, st);
  return sk_bits<__nv_bfloat16>(a, gs, nbits, ws, ws_bytes, st);
}

}  // namespace hqq
