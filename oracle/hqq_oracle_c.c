/* CPU oracle for the HQQ quantize-and-infer hot path, plain C + OpenMP  --  TEST INFRASTRUCTURE ONLY.
 *
 * A second, independent restatement of the reference algorithm (mobiusml/hqq @ e0b1d00) next to oracle/hqq_oracle.py:
 * the same arithmetic, written as loops over the reference's data flow, multi-threaded so that bench.py's `cpu_baseline` /
 * `--impl reference` legs time the reference's CPU path with every host core (torch's CPU operators are OpenMP-parallel too).
 * Only tests/, __graft_entry__ and those two bench legs may build or load it; the product package never does.
 *
 * Parity pinning: tests/test_oracle_c.py checks every entry point against the .npz fixtures under tests/golden, which
 * tests/golden/make_golden.py produced by importing the real reference in the build container, and against the numpy oracle.
 *
 * Reference map (paths relative to /root/reference):
 *   hqq_oc_pack / hqq_oc_unpack      hqq/core/bitpack.py:10-144          (slab interleave along dim 0, MSB field first)
 *   shrink_lp                        hqq/core/optimize.py:96-108
 *   hqq_oc_quantize                  hqq/core/quantize.py:102-176 + hqq/core/optimize.py:201-255 (float32 = the CPU dtype)
 *   hqq_oc_dequantize                hqq/core/quantize.py:184-199, Quantizer.to_inplace :202-217
 *   hqq_oc_linear_forward_f32        hqq/core/quantize.py:880-898 (HQQBackend.PYTORCH: dequantise the whole matrix, matmul)
 *
 * Build: oracle/build_c.py (gcc -O2 -fopenmp -ffp-contract=off -shared -fPIC).  No FMA contraction, no fast-math: the final
 * levels depend on the rounding of W*s + z.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OC_OK 0
#define OC_E_INVALID (-1)

int hqq_oc_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Team size of every following parallel region (bench.py pins it: one socket's physical cores, whatever OMP_NUM_THREADS a
 * launcher such as torchrun exported).  Returns the size now in effect. */
int hqq_oc_set_threads(int n) {
#ifdef _OPENMP
  if (n >= 1) omp_set_num_threads(n);
  return omp_get_max_threads();
#else
  (void)n;
  return 1;
#endif
}

static int fields_of(int nbits) { return nbits == 3 ? 10 : 8 / nbits; }
static int valid_bits(int nbits) { return nbits == 8 || nbits == 4 || nbits == 3 || nbits == 2 || nbits == 1; }

/* rows of the packed tensor for `rows` unpacked rows (bitpack.py:26,45,71-76,118) */
int64_t hqq_oc_packed_rows(int nbits, int64_t rows) {
  const int f = fields_of(nbits);
  if (nbits == 3) return (rows + 9) / 10; /* zero-padded to a multiple of 10 */
  return rows / f;                        /* int(len / f): the reference requires divisibility */
}

/* ---- BitPack.pack_* : levels [rows, cols] (one byte each) -> packed [packed_rows, cols] (u8, or int32 for 3-bit) ------------- */
int hqq_oc_pack(int nbits, const uint8_t* levels, int64_t rows, int64_t cols, void* out) {
  if (!valid_bits(nbits) || !levels || !out || rows < 0 || cols < 0) return OC_E_INVALID;
  const int f = fields_of(nbits);
  const int64_t step = hqq_oc_packed_rows(nbits, rows);
  if (nbits == 3) {
    int32_t* o = (int32_t*)out;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < step; ++i)
      for (int64_t c = 0; c < cols; ++c) {
        uint32_t w = 0;
        for (int j = 0; j < 10; ++j) {
          const int64_t r = i + (int64_t)j * step;
          const uint32_t q = r < rows ? levels[r * cols + c] : 0u; /* padding rows are zero (bitpack.py:72-76) */
          w |= q << (27 - 3 * j);
        }
        o[i * cols + c] = (int32_t)w;
      }
    return OC_OK;
  }
  if (rows % f) return OC_E_INVALID;
  uint8_t* o = (uint8_t*)out;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < step; ++i)
    for (int64_t c = 0; c < cols; ++c) {
      unsigned w = 0;
      for (int j = 0; j < f; ++j) w |= ((unsigned)levels[(i + (int64_t)j * step) * cols + c] << (8 - nbits * (j + 1))) & 0xFFu; /* uint8 shift wraps */
      o[i * cols + c] = (uint8_t)w;
    }
  return OC_OK;
}

/* ---- BitPack.unpack_* : packed [prows, cols] -> levels [f * prows, cols] (3-bit: the padded rows included) -------------------- */
int hqq_oc_unpack(int nbits, const void* packed, int64_t prows, int64_t cols, uint8_t* out) {
  if (!valid_bits(nbits) || !packed || !out || prows < 0 || cols < 0) return OC_E_INVALID;
  const int f = fields_of(nbits);
  if (nbits == 3) {
    const int32_t* p = (const int32_t*)packed;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < prows; ++i)
      for (int64_t c = 0; c < cols; ++c) {
        const uint32_t w = (uint32_t)p[i * cols + c];
        for (int j = 0; j < 10; ++j) out[(i + (int64_t)j * prows) * cols + c] = (uint8_t)((w >> (27 - 3 * j)) & 7u);
      }
    return OC_OK;
  }
  const uint8_t* p = (const uint8_t*)packed;
  const unsigned mask = (1u << nbits) - 1u;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < prows; ++i)
    for (int64_t c = 0; c < cols; ++c) {
      const unsigned w = p[i * cols + c];
      for (int j = 0; j < f; ++j) out[(i + (int64_t)j * prows) * cols + c] = (uint8_t)((w >> (8 - nbits * (j + 1))) & mask);
    }
  return OC_OK;
}

/* ---- rounding to the compute dtype (values stay float32-typed): 0 = float32, 1 = float16, 2 = bfloat16 ----------------------- */
static float round_bf16(float x) {
  uint32_t u;
  memcpy(&u, &x, 4);
  if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x007FFFFFu)) return x; /* NaN */
  u = (u + 0x7FFFu + ((u >> 16) & 1u)) & 0xFFFF0000u;                  /* round to nearest even */
  memcpy(&x, &u, 4);
  return x;
}

static float round_f16(float x) {
  /* float32 -> nearest-even binary16 -> float32, by hand (no reliance on _Float16 support) */
  uint32_t u;
  memcpy(&u, &x, 4);
  const uint32_t sign = u & 0x80000000u;
  uint32_t a = u & 0x7FFFFFFFu;
  if (a >= 0x7F800000u) return x;                      /* inf / NaN */
  if (a >= 0x477FF000u) {                              /* >= 65520 rounds to inf */
    a = 0x7F800000u;
  } else if (a < 0x38800000u) {                        /* below 2^-14: binary16 subnormal, spacing 2^-24 */
    float f;
    memcpy(&f, &a, 4);
    f = f * 16777216.0f;                               /* exact scaling by 2^24 */
    f = nearbyintf(f);                                 /* half-to-even in the default rounding mode */
    f = f * (1.0f / 16777216.0f);
    memcpy(&a, &f, 4);
  } else {                                             /* normal: keep 10 mantissa bits */
    a = (a + 0xFFFu + ((a >> 13) & 1u)) & 0xFFFFE000u;
  }
  a |= sign;
  memcpy(&x, &a, 4);
  return x;
}

static float round_to(float x, int dtype) { return dtype == 1 ? round_f16(x) : (dtype == 2 ? round_bf16(x) : x); }

/* ---- Quantizer.dequantize: unpack -> cast -> (W - zero) rounded -> * scale rounded -> [N, K] ----------------------------------
 * axis = 1: levels [R = N*K/gs, gs], meta [R];  axis = 0: levels [gs, C = N*K/gs], meta [C].  `out` holds N*K float32 values that
 * are representable in the compute dtype.  3-bit: the padded rows are dropped (quantize.py:190-195).                            */
int hqq_oc_dequantize(const void* Wq, const float* scale, const float* zero, int64_t N, int64_t K, int gs, int nbits, int axis, int dtype,
                      float* out) {
  if (!valid_bits(nbits) || !Wq || !scale || !zero || !out || gs <= 0 || (N * K) % gs || (axis != 0 && axis != 1)) return OC_E_INVALID;
  const int64_t rows = axis == 1 ? N * K / gs : gs, cols = axis == 1 ? gs : N * K / gs;
  const int64_t prows = hqq_oc_packed_rows(nbits, rows);
  const int f = fields_of(nbits);
  uint8_t* lv = (uint8_t*)malloc((size_t)(prows * f * cols) + 1);
  if (!lv) return OC_E_INVALID;
  int rc = hqq_oc_unpack(nbits, Wq, prows, cols, lv);
  if (rc == OC_OK) {
#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < rows; ++r)
      for (int64_t c = 0; c < cols; ++c) {
        const int64_t g = axis == 1 ? r : c;
        const float z = round_to(zero[g], dtype), s = round_to(scale[g], dtype);
        const float d = round_to((float)lv[r * cols + c] - z, dtype);
        out[r * cols + c] = round_to(d * s, dtype); /* d*s is exact in float32 for 16-bit operands: one rounding */
      }
  }
  free(lv);
  return rc;
}

/* ---- HQQBackend.PYTORCH forward at float32, the reference's data flow: whole-matrix unpack, subtract, multiply (three passes
 * over an N*K float matrix, quantize.py:184-199), then y = x @ W_r^T (+ bias) (quantize.py:880-898).  `W_r` is caller scratch of
 * N*K floats -- the dequantised matrix the reference materialises on every call.  axis = 1 only (the path's configuration).     */
int hqq_oc_linear_forward_f32(const float* x, int64_t M, const void* Wq, const float* scale, const float* zero, const float* bias,
                              int64_t N, int64_t K, int gs, int nbits, float* W_r, float* y) {
  if (!valid_bits(nbits) || !x || !Wq || !scale || !zero || !W_r || !y || gs <= 0 || K % gs) return OC_E_INVALID;
  const int64_t rows = N * K / gs, cols = gs;
  const int64_t prows = hqq_oc_packed_rows(nbits, rows);
  const int f = fields_of(nbits);
  /* pass 1: unpack + cast to float32 (bitpack.py unpack_*(dtype=compute_dtype)) */
  if (nbits == 3) {
    const int32_t* p = (const int32_t*)Wq;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < prows; ++i)
      for (int64_t c = 0; c < cols; ++c) {
        const uint32_t w = (uint32_t)p[i * cols + c];
        for (int j = 0; j < 10; ++j) {
          const int64_t r = i + (int64_t)j * prows;
          if (r < rows) W_r[r * cols + c] = (float)((w >> (27 - 3 * j)) & 7u);
        }
      }
  } else {
    const uint8_t* p = (const uint8_t*)Wq;
    const unsigned mask = (1u << nbits) - 1u;
    for (int j = 0; j < f; ++j) { /* one masked-shift pass per field, as the reference issues them */
      float* dst = W_r + (int64_t)j * prows * cols;
      const int sh = 8 - nbits * (j + 1);
#pragma omp parallel for schedule(static)
      for (int64_t i = 0; i < prows * cols; ++i) dst[i] = (float)((p[i] >> sh) & mask);
    }
  }
  /* passes 2 and 3: (W_r - zero) then * scale, broadcast over the group (quantize.py:198) */
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float z = zero[r];
    for (int64_t c = 0; c < cols; ++c) W_r[r * cols + c] -= z;
  }
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < rows; ++r) {
    const float s = scale[r];
    for (int64_t c = 0; c < cols; ++c) W_r[r * cols + c] *= s;
  }
  /* matmul: y[m, n] = sum_k x[m, k] * W_r[n, k] */
#pragma omp parallel for schedule(static)
  for (int64_t n = 0; n < N; ++n) {
    const float* w = W_r + n * K;
    for (int64_t m = 0; m < M; ++m) {
      const float* xr = x + m * K;
      float acc = 0.0f;
#pragma omp simd reduction(+ : acc)
      for (int64_t k = 0; k < K; ++k) acc += xr[k] * w[k];
      y[m * N + n] = acc + (bias ? bias[n] : 0.0f);
    }
  }
  return OC_OK;
}

/* ---- shrink_lp_op (optimize.py:96-108): sign(x) * relu(|x| - (1/beta) * |x|^(p-1)); x = 0, p < 1 -> 0 ------------------------- */
static float shrink_lp(float x, float inv_beta, float lp_norm) {
  const float a = fabsf(x);
  float e;
  if (lp_norm == 1.0f) e = a - inv_beta;
  else e = a - inv_beta * powf(a, lp_norm - 1.0f); /* a = 0: powf = +inf -> e = -inf */
  if (!(e > 0.0f)) e = 0.0f;
  return x > 0.0f ? e : (x < 0.0f ? -e : 0.0f);
}

static float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---- Quantizer.quantize without the packing (quantize.py:102-176): group min/max -> inverse scale / zero -> proximal solver with
 * the whole-tensor early stop (optimize.py:209-255) -> final round/clamp.  Outputs: levels (one byte each, group layout [R, gs]
 * or [gs, C]), scale_out = 1/inverse-scale (quantize.py:154) and zero_out per group, the number of solver iterations executed and
 * the per-iteration mean |W - W_r| (float32, as the reference compares them).  Group sums are accumulated in float64 and rounded
 * once (torch: float32 with an unspecified order; see oracle/hqq_oracle.py).                                                     */
int hqq_oc_quantize(const float* W, int64_t N, int64_t K, int gs, int nbits, int axis, int round_zero, int optimize, float lp_norm,
                    float beta, int iters, uint8_t* levels, float* scale_out, float* zero_out, int* iters_done, float* errors) {
  if (!valid_bits(nbits) || !W || !levels || !scale_out || !zero_out || gs <= 0 || (N * K) % gs || (axis != 0 && axis != 1) || iters < 0)
    return OC_E_INVALID;
  const int64_t total = N * K, G = total / gs;
  /* element e of group g: axis 1 -> W[g*gs + e]; axis 0 -> W[e*G + g] */
  const int64_t gstride = axis == 1 ? gs : 1, estride = axis == 1 ? 1 : G;
  const float maxv = (float)((1 << nbits) - 1);
  float* s_inv = (float*)malloc(sizeof(float) * (size_t)G);
  float* zero = (float*)malloc(sizeof(float) * (size_t)G);
  float* znew = (float*)malloc(sizeof(float) * (size_t)G);
  if (!s_inv || !zero || !znew) { free(s_inv); free(zero); free(znew); return OC_E_INVALID; }
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < G; ++g) {
    const float* w = W + g * gstride;
    float mn = w[0], mx = w[0];
    for (int e = 1; e < gs; ++e) {
      const float v = w[e * estride];
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
    const float denom = mx - mn;
    float s = (1.0f / denom) * maxv; /* `max_v / denom` is reciprocal(denom) * max_v in torch: two roundings */
    if (fabsf(denom) <= 1e-4f) s = 1.0f;
    if (s > 2e4f) s = 2e4f;
    float z = -mn * s;
    if (round_zero) z = nearbyintf(z);
    s_inv[g] = s;
    zero[g] = z;
  }
  int done = 0;
  if (optimize) {
    const float inv_beta = (float)(1.0 / (double)beta);
    float best = INFINITY;
    for (int it = 0; it < iters; ++it) {
      double err_sum = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : err_sum)
      for (int64_t g = 0; g < G; ++g) {
        const float* w = W + g * gstride;
        const float s = s_inv[g], z = zero[g];
        double zs = 0.0, es = 0.0;
        for (int e = 0; e < gs; ++e) {
          const float wf = w[e * estride];
          const float q = clampf(nearbyintf(wf * s + z), 0.0f, maxv);
          const float wr = (q - z) / s;
          const float d = wf - wr;
          es += (double)fabsf(d);
          const float we = shrink_lp(d, inv_beta, lp_norm);
          zs += (double)(q - (wf - we) * s);
        }
        znew[g] = (float)(zs / (double)gs);
        err_sum += es;
      }
      float* t = zero; zero = znew; znew = t; /* the zero of the breaking iteration is kept (optimize.py:239-247) */
      const float err = (float)(err_sum / (double)total);
      if (errors) errors[it] = err;
      ++done;
      if (err < best) best = err;
      else break;
    }
  }
  if (iters_done) *iters_done = done;
#pragma omp parallel for schedule(static)
  for (int64_t g = 0; g < G; ++g) {
    const float* w = W + g * gstride;
    const float s = s_inv[g], z = zero[g];
    for (int e = 0; e < gs; ++e) levels[g * gstride + e * estride] = (uint8_t)clampf(nearbyintf(w[e * estride] * s + z), 0.0f, maxv);
    scale_out[g] = 1.0f / s;
    zero_out[g] = z;
  }
  free(s_inv); free(zero); free(znew);
  return OC_OK;
}
