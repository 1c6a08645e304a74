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
// K is split across the 8 warps of a CTA (contiguous chunks) and reduced through shared memory in a fixed
// order: no atomics, deterministic output.
#include <stdlib.h>
#include <string.h>
#include <type_traits>

#include "common.cuh"

namespace hqq {

struct GemvArgs {
  const void* x;
  const uint8_t* Wq;
  const void* scale;
  const void* zero;
  const void* bias;
  void* y;
  int M, N, K;
  int step;  // packed rows = N / F
  int Gk;    // groups per output row = K / GS
  int S;     // k64-steps per warp
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

// 16-bit -> fp32 widening of up to 8 consecutive scale/zero values (one vector load per batch)
template <typename T, int NV> struct MetaVec {
  float v[NV];
  __device__ __forceinline__ void load(const T* p) {
    Vec<T, NV> r = *reinterpret_cast<const Vec<T, NV>*>(p);
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = to_f32<T>(r.v[i]);
  }
  __device__ __forceinline__ void zero_fill() {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = 0.0f;
  }
};

template <typename T, int NBITS, int GS, int MT, int MAGIC>
__global__ void __launch_bounds__(256, (MT <= 2 ? 2 : 1)) linear_small_kernel(GemvArgs a) {
  constexpr int F = 8 / NBITS;               // fields (slabs) per byte
  constexpr int P = 16 / F;                  // packed rows per 16-row MMA tile
  constexpr int MPG = GS / 16;               // MMAs per quantisation group
  constexpr int U = 4;                       // k64-steps per register batch (= 256 k)
  constexpr int GPB = 256 / GS;              // quantisation groups per batch (GS <= 256)
  constexpr int NW = 8;                      // warps per CTA
  using MM = MT16<T>;
  __shared__ float red[NW][MT][16][8];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane >> 2, c = lane & 3;
  const int p = (F == 1) ? r : (r % P);
  const int fa = (F == 1) ? 0 : (r / P), fb = (F == 1) ? 0 : (F / 2 + r / P);
  const int prow_a = blockIdx.x * P + p;
  const int prow_b = (F == 1) ? prow_a + 8 : prow_a;
  const bool ok_a = prow_a < a.step, ok_b = prow_b < a.step;
  const int n_a = fa * a.step + prow_a;      // output rows of the thread's two fragment rows
  const int n_b = fb * a.step + prow_b;

  Lanes<T, NBITS, MAGIC> lanes;
  lanes.init(8 - NBITS * (fa + 1), 8 - NBITS * (fb + 1));

  const int nsteps = a.K >> 6;               // multiple of U (K % 256 == 0)
  const int s0 = warp * a.S;                 // a.S is a multiple of U
  const int s1 = min(nsteps, s0 + a.S);

  float tot[MT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int i = 0; i < 4; ++i) tot[mt][i] = 0.0f;

  if (s0 < s1) {
    // running pointers (advanced once per batch)
    const uint8_t* wp_a = a.Wq + (long long)(ok_a ? prow_a : 0) * a.K + ((long long)s0 << 6) + 16 * c;
    const uint8_t* wp_b = a.Wq + (long long)(ok_b ? prow_b : 0) * a.K + ((long long)s0 << 6) + 16 * c;
    const T* sp_a = reinterpret_cast<const T*>(a.scale) + (long long)(ok_a ? n_a : 0) * a.Gk + (s0 * 64) / GS;
    const T* zp_a = reinterpret_cast<const T*>(a.zero) + (long long)(ok_a ? n_a : 0) * a.Gk + (s0 * 64) / GS;
    const T* sp_b = reinterpret_cast<const T*>(a.scale) + (long long)(ok_b ? n_b : 0) * a.Gk + (s0 * 64) / GS;
    const T* zp_b = reinterpret_cast<const T*>(a.zero) + (long long)(ok_b ? n_b : 0) * a.Gk + (s0 * 64) / GS;
    const T* xp[MT];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
      const int m = min(mt * 8 + r, a.M - 1);
      xp[mt] = reinterpret_cast<const T*>(a.x) + (long long)m * a.K + ((long long)s0 << 6) + 16 * c;
    }

    float Sg[MT][4], Xg[MT][4];
    uint4 nxt_a[U], nxt_b[U];
    auto load_batch = [&]() {
#pragma unroll
      for (int u = 0; u < U; ++u) {
        // rows past the ragged edge re-read row 0 (always mapped); their results are never stored
        nxt_a[u] = ldg_stream_v4(wp_a + u * 64);
        if (F == 1) nxt_b[u] = ldg_stream_v4(wp_b + u * 64);
      }
      wp_a += U * 64;
      wp_b += U * 64;
    };
    load_batch();

    for (int sb = s0; sb < s1; sb += U) {
      uint4 cur_a[U], cur_b[U];
#pragma unroll
      for (int u = 0; u < U; ++u) { cur_a[u] = nxt_a[u]; if (F == 1) cur_b[u] = nxt_b[u]; }
      if (sb + U < s1) load_batch();

      // per-group meta for the whole batch: one vector load per (scale|zero) x (row a|row b)
      MetaVec<T, GPB> sA, zA, sB, zB;
      sA.load(sp_a); zA.load(zp_a);
      sB.load(sp_b); zB.load(zp_b);
      sp_a += GPB; zp_a += GPB; sp_b += GPB; zp_b += GPB;

#pragma unroll
      for (int u = 0; u < U; ++u) {
        // activations for this k64 step: 16 consecutive k per thread, one column (token) per 4-lane group
        uint4 xa[MT], xb[MT];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          // token columns >= M alias the last real token: MMA columns are independent and never stored
          const uint4* q = reinterpret_cast<const uint4*>(xp[mt] + u * 64);
          xa[mt] = __ldg(q);
          xb[mt] = __ldg(q + 1);
        }
        const uint32_t wa[4] = {cur_a[u].x, cur_a[u].y, cur_a[u].z, cur_a[u].w};
        const uint32_t wb[4] = {cur_b[u].x, cur_b[u].y, cur_b[u].z, cur_b[u].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint32_t a0, a1, a2, a3;
          if constexpr (F == 1) lanes.extract2(wa[j], wb[j], a0, a1, a2, a3);
          else lanes.extract(wa[j], a0, a1, a2, a3);
          const bool first = ((u * 4 + j) % MPG) == 0;  // first MMA of a group starts from C = 0
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
          if (((u * 4 + j + 1) % MPG) == 0) {
            // a quantisation group is complete: tot += s*(Q - z*X), with lane value = OFF + q*V folded in
            const int gi = (u * 4 + j) / MPG;  // compile-time index into the batch's meta vectors
            const float ka = sA.v[gi] * lanes.invV_a, la = -sA.v[gi] * (lanes.offV_a + zA.v[gi]);
            const float kb = sB.v[gi] * lanes.invV_b, lb = -sB.v[gi] * (lanes.offV_b + zB.v[gi]);
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
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) xp[mt] += U * 64;
    }
  }

  // fixed-order cross-warp (split-K) reduction
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    red[warp][mt][r][2 * c] = tot[mt][0];
    red[warp][mt][r][2 * c + 1] = tot[mt][1];
    red[warp][mt][r + 8][2 * c] = tot[mt][2];
    red[warp][mt][r + 8][2 * c + 1] = tot[mt][3];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < MT * 128; idx += 256) {
    const int mt = idx >> 7, t = (idx >> 3) & 15, col = idx & 7;
    const int m = mt * 8 + col;
    const int tp = (F == 1) ? t : (t % P), tf = (F == 1) ? 0 : (t / P);
    const int prow = blockIdx.x * P + tp;
    if (m < a.M && prow < a.step) {
      float s = 0.0f;
#pragma unroll
      for (int w = 0; w < NW; ++w) s += red[w][mt][t][col];
      const int n = tf * a.step + prow;
      MM::st(a.y, (long long)m * a.N + n, s, a.bias, n);
    }
  }
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

template <typename T, int NBITS, int GS, int MT, int MAGIC>
static int launch_small(const GemvArgs& a, cudaStream_t st) {
  constexpr int F = 8 / NBITS, P = 16 / F;
  const unsigned grid = (unsigned)cdiv(a.step, P);
  linear_small_kernel<T, NBITS, GS, MT, MAGIC><<<grid, 256, 0, st>>>(a);
  HQQ_LAUNCH_CHECK("hqq_b200_linear_fwd/small");
  return HQQ_OK;
}

template <typename T, int NBITS, int GS, int MAGIC>
static int small_mt(const GemvArgs& a, cudaStream_t st) {
  if (a.M <= 8) return launch_small<T, NBITS, GS, 1, MAGIC>(a, st);
  if (a.M <= 16) return launch_small<T, NBITS, GS, 2, MAGIC>(a, st);
  return launch_small<T, NBITS, GS, 4, MAGIC>(a, st);
}

template <typename T, int NBITS, int MAGIC>
static int small_gs(const GemvArgs& a, int gs, cudaStream_t st) {
  switch (gs) {
    case 32: return small_mt<T, NBITS, 32, MAGIC>(a, st);
    case 64: return small_mt<T, NBITS, 64, MAGIC>(a, st);
    case 128: return small_mt<T, NBITS, 128, MAGIC>(a, st);
    case 256: return small_mt<T, NBITS, 256, MAGIC>(a, st);
  }
  return HQQ_E_UNSUPPORTED;
}

template <typename T>
static int small_bits(const GemvArgs& a, int gs, int nbits, cudaStream_t st) {
  const bool sub = std::is_same<T, __half>::value && magic_mode() == MAGIC_SUBNORMAL;
  switch (nbits) {
    case 8:
      if constexpr (std::is_same<T, __half>::value) return sub ? small_gs<T, 8, MAGIC_SUBNORMAL>(a, gs, st) : small_gs<T, 8, MAGIC_OFFSET>(a, gs, st);
      else return HQQ_E_UNSUPPORTED;
    case 4:
      if constexpr (std::is_same<T, __half>::value) return sub ? small_gs<T, 4, MAGIC_SUBNORMAL>(a, gs, st) : small_gs<T, 4, MAGIC_OFFSET>(a, gs, st);
      else return small_gs<T, 4, MAGIC_OFFSET>(a, gs, st);
    case 2:
      if constexpr (std::is_same<T, __half>::value) return sub ? small_gs<T, 2, MAGIC_SUBNORMAL>(a, gs, st) : small_gs<T, 2, MAGIC_OFFSET>(a, gs, st);
      else return small_gs<T, 2, MAGIC_OFFSET>(a, gs, st);
    case 1:
      if constexpr (std::is_same<T, __half>::value) return sub ? small_gs<T, 1, MAGIC_SUBNORMAL>(a, gs, st) : small_gs<T, 1, MAGIC_OFFSET>(a, gs, st);
      else return small_gs<T, 1, MAGIC_OFFSET>(a, gs, st);
  }
  return HQQ_E_UNSUPPORTED;
}

bool small_route_ok(int64_t M, int64_t N, int64_t K, int gs, int nbits, int axis, int dtype) {
  if (axis != 1) return false;
  if (dtype != HQQ_F16 && dtype != HQQ_BF16) return false;
  if (!(nbits == 8 || nbits == 4 || nbits == 2 || nbits == 1)) return false;
  if (nbits == 8 && dtype == HQQ_BF16) return false;  // 8-bit levels do not fit a bf16 mantissa trick
  if (!(gs == 32 || gs == 64 || gs == 128 || gs == 256)) return false;
  if (M < 1 || M > 32) return false;
  if (K % 256 != 0 || K % gs != 0) return false;  // 256-k register batches; groups never straddle a row
  if (N % (8 / nbits) != 0) return false;
  if (N > (1 << 30) || K > (1 << 30)) return false;
  return true;
}

int linear_small(const void* x, const void* Wq, const void* scale, const void* zero, const void* bias, void* y, int64_t M,
                 int64_t N, int64_t K, int gs, int nbits, int dtype, cudaStream_t st) {
  HQQ_REQUIRE(aligned(x, 16) && aligned(Wq, 16), HQQ_E_INVALID, "hqq_b200_linear_fwd: x and W_q must be 16-byte aligned");
  GemvArgs a;
  a.x = x; a.Wq = (const uint8_t*)Wq; a.scale = scale; a.zero = zero; a.bias = bias; a.y = y;
  a.M = (int)M; a.N = (int)N; a.K = (int)K;
  a.step = (int)(N / (8 / nbits));
  a.Gk = (int)(K / gs);
  const int nsteps = (int)(K / 64);
  a.S = (int)(cdiv(cdiv(nsteps, 8), 4) * 4);
  if (dtype == HQQ_F16) return small_bits<__half>(a, gs, nbits, st);
  return small_bits<__nv_bfloat16>(a, gs, nbits, st);
}

}  // namespace hqq
