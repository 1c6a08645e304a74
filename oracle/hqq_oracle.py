"""CPU oracle for the HQQ quantize-and-infer hot path  --  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch numpy restatement of the reference algorithm
(mobiusml/hqq @ e0b1d00).  It exists so that the CUDA path in ``hqq_b200`` can be
checked against something that does not share code with it.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import it; the product package never does (and fails loudly when
its CUDA library is missing instead of falling back to this).

Parity pinning: every function here is checked in ``tests/test_oracle_golden.py``
against fixtures in ``tests/golden/*.npz`` that were produced by importing the real
reference (``/root/reference``) in the build container with
``tests/golden/make_golden.py``.

Reference map (paths relative to /root/reference):
  pack_* / unpack_*            hqq/core/bitpack.py:10-144
  shrink_lp_op                 hqq/core/optimize.py:96-108
  proximal_step / solver       hqq/core/optimize.py:201-255
  quantize_init / quantize     hqq/core/quantize.py:76-180
  dequantize                   hqq/core/quantize.py:184-199
  linear_forward               hqq/core/quantize.py:880-898 (HQQBackend.PYTORCH)
"""
from __future__ import annotations

import numpy as np

SUPPORTED_BITS = (8, 4, 3, 2, 1)
BIT_TO_PACKING = {8: "8bit_u8", 4: "4bit_u8", 3: "3bit_32", 2: "2bit_u8", 1: "1bit_u8"}
FIELDS = {"8bit_u8": 1, "4bit_u8": 2, "2bit_u8": 4, "1bit_u8": 8, "3bit_32": 10}
NBITS_OF = {"8bit_u8": 8, "4bit_u8": 4, "2bit_u8": 2, "1bit_u8": 1, "3bit_32": 3}


# ----------------------------------------------------------------------------------
# low-precision helpers (numpy has float16 but no bfloat16)
# ----------------------------------------------------------------------------------
def round_to_bf16(x: np.ndarray) -> np.ndarray:
    """fp32 -> nearest-even bfloat16, returned as fp32 values that are bf16-representable."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    u = x.view(np.uint32).astype(np.uint64)
    nan = np.isnan(x)
    lsb = (u >> 16) & 1
    u = (u + 0x7FFF + lsb) & 0xFFFF0000
    out = u.astype(np.uint32).view(np.float32).copy()
    out[nan] = np.nan
    return out


def round_to(x: np.ndarray, dtype: str) -> np.ndarray:
    """Round an fp32/fp64 array to `dtype` in {"float32","float16","bfloat16"}; result is fp32-valued."""
    if dtype == "float32":
        return np.asarray(x, dtype=np.float32)
    if dtype == "float16":
        with np.errstate(over="ignore"):
            return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)
    if dtype == "bfloat16":
        return round_to_bf16(np.asarray(x, dtype=np.float32))
    raise ValueError(dtype)


# ----------------------------------------------------------------------------------
# BitPack (bitpack.py:10-144).  Layout: "slab interleave" along dim 0 -- field f of
# packed row i holds unpacked row i + f*step, most significant field first.
# ----------------------------------------------------------------------------------
def _to_u8(W_q: np.ndarray) -> np.ndarray:
    # torch `.to(uint8)` on float tensors truncates; on ints it wraps modulo 256
    a = np.asarray(W_q)
    if a.dtype.kind == "f":
        a = a.astype(np.int64)
    return (a.astype(np.int64) & 0xFF).astype(np.uint8)


def pack_8bit_u8(W_q):  # bitpack.py:14-15
    return _to_u8(W_q)


def unpack_8bit_u8(W_q):  # bitpack.py:18-19
    return np.asarray(W_q, dtype=np.uint8).copy()


def _pack_u8(W_q, nbits):
    w = _to_u8(W_q)
    f = 8 // nbits
    step = w.shape[0] // f  # int(len/2) etc. (bitpack.py:26,45,118)
    out = np.zeros((step,) + w.shape[1:], dtype=np.uint8)
    for j in range(f):
        shift = 8 - nbits * (j + 1)
        # uint8 shift wraps modulo 256, exactly like torch's `uint8 << k`
        out |= ((w[j * step:(j + 1) * step].astype(np.uint16) << shift) & 0xFF).astype(np.uint8)
    return out


def _unpack_u8(W_q, nbits):
    w = np.asarray(W_q, dtype=np.uint8)
    f = 8 // nbits
    step = w.shape[0]
    out = np.empty((f * step,) + w.shape[1:], dtype=np.uint8)
    mask = (1 << nbits) - 1
    for j in range(f):
        shift = 8 - nbits * (j + 1)
        out[j * step:(j + 1) * step] = (w >> shift) & mask
    return out


def pack_4bit_u8(W_q):  # bitpack.py:24-28
    return _pack_u8(W_q, 4)


def unpack_4bit_u8(W_q):  # bitpack.py:31-38
    return _unpack_u8(W_q, 4)


def pack_2bit_u8(W_q):  # bitpack.py:43-52
    return _pack_u8(W_q, 2)


def unpack_2bit_u8(W_q):  # bitpack.py:55-64
    return _unpack_u8(W_q, 2)


def pack_1bit_u8(W_q):  # bitpack.py:115-128
    return _pack_u8(W_q, 1)


def unpack_1bit_u8(W_q):  # bitpack.py:131-144
    return _unpack_u8(W_q, 1)


def pack_3bit_32(W_q_in):  # bitpack.py:69-91 : rows zero-padded to a multiple of 10, 10 fields / int32
    a = np.asarray(W_q_in)
    if a.dtype.kind == "f":
        a = a.astype(np.int64)
    a = a.astype(np.int64).astype(np.int32)
    rows = int(10 * np.ceil(a.shape[0] / 10.0))
    w = np.zeros((rows,) + a.shape[1:], dtype=np.int32)
    w[: a.shape[0]] = a
    step = rows // 10
    out = np.zeros((step,) + a.shape[1:], dtype=np.int32)
    for j in range(10):
        out |= (w[j * step:(j + 1) * step] << (27 - 3 * j)).astype(np.int32)
    return out


def unpack_3bit_32(W_q):  # bitpack.py:95-110 (returns padded rows; the caller slices)
    w = np.asarray(W_q, dtype=np.int32)
    step = w.shape[0]
    out = np.empty((10 * step,) + w.shape[1:], dtype=np.uint8)
    for j in range(10):
        out[j * step:(j + 1) * step] = ((w >> (27 - 3 * j)) & 7).astype(np.uint8)
    return out


PACK = {"8bit_u8": pack_8bit_u8, "4bit_u8": pack_4bit_u8, "3bit_32": pack_3bit_32,
        "2bit_u8": pack_2bit_u8, "1bit_u8": pack_1bit_u8}
UNPACK = {"8bit_u8": unpack_8bit_u8, "4bit_u8": unpack_4bit_u8, "3bit_32": unpack_3bit_32,
          "2bit_u8": unpack_2bit_u8, "1bit_u8": unpack_1bit_u8}


# ----------------------------------------------------------------------------------
# Proximal solver (optimize.py:96-108, 201-255), float32 = the reference's CPU dtype
# ----------------------------------------------------------------------------------
def shrink_lp_op(x: np.ndarray, beta: float, lp_norm: float) -> np.ndarray:
    """sign(x) * relu(|x| - (1/beta) * |x|^(p-1))   (optimize.py:96-108).
    At x == 0 with p < 1: 0^(p-1) = inf -> |x| - inf = -inf -> relu -> 0 ; times sign(0) = 0."""
    f32 = np.float32
    a = np.abs(x).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore", over="ignore"):
        if lp_norm == 1:
            out = np.maximum(a - f32(1.0 / beta), f32(0))
        else:
            out = np.maximum(a - f32(1.0 / beta) * np.power(a, f32(lp_norm - 1)), f32(0))
        out = out * np.sign(x).astype(f32)
    return out.astype(f32)


def proximal_step(W_f, scale, zero, min_max, beta, lp_norm, axis):
    """One half-quadratic iteration (optimize.py:201-206). `scale` is the INVERSE scale."""
    f32 = np.float32
    W_q = np.clip(np.round(W_f * scale + zero), f32(min_max[0]), f32(min_max[1])).astype(f32)
    W_r = ((W_q - zero) / scale).astype(f32)
    W_e = shrink_lp_op(W_f - W_r, beta, lp_norm)
    # group mean: float64 accumulation rounded once -- on the golden fixtures this reproduces torch.mean (float32, vectorised
    # partial sums) level for level, where a float32 pairwise sum flips 2 of 393 216 levels
    zero = np.mean((W_q - (W_f - W_e) * scale).astype(f32), axis=axis, keepdims=True, dtype=np.float64)
    return W_r, W_q, zero.astype(f32)


def optimize_weights_proximal(W_f, scale, zero, min_max, axis,
                              lp_norm=0.7, beta=1e1, iters=20, return_trace=False):
    """optimize.py:209-255 on the CPU/float32 path.

    Quirks kept on purpose: `beta` is never multiplied by kappa and `scale` is never
    updated; the early stop compares the WHOLE-TENSOR mean |W - W_r| with the best so
    far and breaks at the first non-decrease -- returning the zero computed *in* the
    breaking iteration.  The error mean is accumulated in float64 and rounded to
    float32 before the comparison (torch accumulates in float32 with its own order; the
    difference only matters when consecutive errors agree to ~1e-7 relative).
    """
    f32 = np.float32
    W_f = np.asarray(W_f, dtype=f32)
    scale = np.asarray(scale, dtype=f32)
    zero = np.asarray(zero, dtype=f32)
    best = f32(np.inf)
    errs = []
    n_done = 0
    for _ in range(iters):
        W_r, _Wq, zero = proximal_step(W_f, scale, zero, min_max, beta, lp_norm, axis)
        err = f32(np.mean(np.abs(W_f - W_r), dtype=np.float64))
        errs.append(float(err))
        n_done += 1
        if err < best:
            best = err
        else:
            break
    W_q = np.clip(np.round(W_f * scale + zero), f32(min_max[0]), f32(min_max[1])).astype(f32)
    if return_trace:
        return W_q, scale, zero, n_done, errs
    return W_q, scale, zero


# ----------------------------------------------------------------------------------
# Quantizer.quantize / dequantize (quantize.py:76-199)
# ----------------------------------------------------------------------------------
def quantize_init(tensor, nbits, group_size, axis, round_zero, channel_wise=True):
    """Group reshape, min/max, inverse-scale / zero init (quantize.py:102-134)."""
    f32 = np.float32
    W = np.asarray(tensor, dtype=f32)
    if group_size is not None and channel_wise:
        W = W.reshape(-1, group_size) if axis == 1 else W.reshape(group_size, -1)
    if not channel_wise:
        _min, _max = W.min(), W.max()
    else:
        _min = W.min(axis=axis, keepdims=True)
        _max = W.max(axis=axis, keepdims=True)
    max_v = round(2 ** nbits - 1)
    denom = (_max - _min).astype(f32)
    with np.errstate(divide="ignore", invalid="ignore"):
        # `max_v / denom` with a python scalar on the left is Tensor.__rtruediv__, which torch
        # evaluates as reciprocal(denom) * max_v -- two float32 roundings, reproduced here.
        scale = ((f32(1.0) / denom).astype(f32) * f32(max_v)).astype(f32)
    scale = np.where(np.abs(denom) <= f32(1e-4), f32(1.0), scale).astype(f32)
    scale = np.minimum(scale, f32(2e4))
    zero = (-_min * scale).astype(f32)
    if round_zero:
        zero = np.round(zero).astype(f32)
    return W, scale, zero, [0, max_v]


def quantize(tensor, nbits=4, channel_wise=True, group_size=64, optimize=True,
             round_zero=False, axis=0, bitpack=True, return_trace=False):
    """Quantizer.quantize (quantize.py:76-180). Returns (W_q, meta); meta['scale'] is the
    dequantisation scale (1/inverse-scale, quantize.py:154), float32, shape [R,1] or [1,C]."""
    assert nbits in SUPPORTED_BITS
    assert axis in (0, 1)
    shape = tuple(np.asarray(tensor).shape)
    W, scale, zero, min_max = quantize_init(tensor, nbits, group_size, axis, round_zero, channel_wise)
    trace = None
    if optimize and channel_wise:
        res = optimize_weights_proximal(W, scale, zero, min_max, axis, return_trace=True)
        W_q, scale, zero, n_done, errs = res
        trace = {"iters": n_done, "errors": errs}
    else:
        W_q = np.clip(np.round(W * scale + zero), min_max[0], min_max[1]).astype(np.float32)
    meta = {
        "nbits": nbits, "group_size": group_size, "shape": shape,
        "scale": (np.float32(1.0) / scale).astype(np.float32), "zero": zero.astype(np.float32),
        "axis": axis, "packing": BIT_TO_PACKING[nbits],
    }
    if bitpack:
        W_q = PACK[meta["packing"]](W_q)
    else:
        meta["packing"] = None
    if return_trace:
        return W_q, meta, trace
    return W_q, meta


def dequantize(W_q, meta, compute_dtype="float32"):
    """Quantizer.dequantize (quantize.py:184-199): unpack -> cast to compute_dtype ->
    (W - zero) rounded to compute_dtype -> * scale rounded to compute_dtype -> reshape.
    meta scale/zero are first cast to compute_dtype (Quantizer.to_inplace, quantize.py:202-217)."""
    if meta["packing"]:
        W_r = UNPACK[meta["packing"]](W_q).astype(np.float32)
        if meta["nbits"] == 3:
            n = int(np.prod(meta["shape"]))
            rows = meta["group_size"] if meta["axis"] == 0 else n // meta["group_size"]
            W_r = W_r[:rows]
    else:
        W_r = np.asarray(W_q, dtype=np.float32)
    z = round_to(meta["zero"], compute_dtype)
    s = round_to(meta["scale"], compute_dtype)
    d = round_to(W_r - z, compute_dtype)
    out = round_to(d.astype(np.float64) * s.astype(np.float64), compute_dtype)
    return out.reshape(meta["shape"])


def linear_forward(x, W_q, meta, bias=None, compute_dtype="float32"):
    """HQQBackend.PYTORCH forward (quantize.py:880-898): y = x @ dequantize().T (+ bias).
    Products are accumulated in float64 and rounded once to compute_dtype, i.e. this is the
    exact-arithmetic value of the reference's GEMM (whose accumulation order is unspecified)."""
    W_r = dequantize(W_q, meta, compute_dtype).astype(np.float64)
    xs = round_to(np.asarray(x, dtype=np.float32), compute_dtype).astype(np.float64)
    y = xs.reshape(-1, xs.shape[-1]) @ W_r.T
    y = round_to(y, compute_dtype)
    if bias is not None:
        y = round_to(y.astype(np.float64) + round_to(bias, compute_dtype), compute_dtype)
    return y.reshape(tuple(xs.shape[:-1]) + (W_r.shape[0],))


def linear_forward_f32_fast(x, W_q, meta, bias=None):
    """The reference's CPU data flow at float32 (quantize.py:184-199, 880-898) without the float64 bookkeeping of
    `linear_forward`: unpack the whole matrix, (W - zero) * scale, then one BLAS matmul.  Used only as the timed
    CPU baseline in bench.py (it does per token exactly the passes over the N x K matrix the reference does)."""
    W_r = UNPACK[meta["packing"]](W_q).astype(np.float32)
    W_r -= meta["zero"]
    W_r *= meta["scale"]
    y = np.asarray(x, dtype=np.float32) @ W_r.reshape(meta["shape"]).T
    if bias is not None:
        y = y + bias
    return y
