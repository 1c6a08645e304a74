"""ctypes face of oracle/hqq_oracle_c.c  --  TEST INFRASTRUCTURE ONLY (see the C file's header for the reference map).

Same call shapes as oracle/hqq_oracle.py (`quantize`, `dequantize`, `PACK`, `UNPACK`, `linear_forward_f32`), so the tests can run
both restatements side by side; `threads()` is the OpenMP team size the timed CPU baseline reports as `cores`."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int64, c_void_p

import numpy as np

from . import build_c

BIT_TO_PACKING = {8: "8bit_u8", 4: "4bit_u8", 3: "3bit_32", 2: "2bit_u8", 1: "1bit_u8"}
NBITS_OF = {v: k for k, v in BIT_TO_PACKING.items()}
DTYPES = {"float32": 0, "float16": 1, "bfloat16": 2}
_lib = None


def lib():
    global _lib
    if _lib is None:
        L = ctypes.CDLL(build_c.build())
        L.hqq_oc_threads.restype = c_int
        L.hqq_oc_set_threads.restype = c_int
        L.hqq_oc_set_threads.argtypes = [c_int]
        L.hqq_oc_packed_rows.restype = c_int64
        L.hqq_oc_packed_rows.argtypes = [c_int, c_int64]
        L.hqq_oc_pack.argtypes = [c_int, c_void_p, c_int64, c_int64, c_void_p]
        L.hqq_oc_unpack.argtypes = [c_int, c_void_p, c_int64, c_int64, c_void_p]
        L.hqq_oc_dequantize.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p]
        L.hqq_oc_linear_forward_f32.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                                c_void_p, c_void_p]
        L.hqq_oc_quantize.argtypes = [c_void_p, c_int64, c_int64, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]
        _lib = L
    return _lib


def threads() -> int:
    return int(lib().hqq_oc_threads())


def set_threads(n: int) -> int:
    """Pin the OpenMP team size (overrides OMP_NUM_THREADS); returns the size in effect."""
    return int(lib().hqq_oc_set_threads(int(n)))


def _p(a):
    return None if a is None else a.ctypes.data_as(c_void_p)


def _check(rc, what):
    if rc != 0:
        raise ValueError(f"hqq_oracle_c.{what}: invalid arguments (code {rc})")


def _levels_u8(W_q):
    a = np.asarray(W_q)
    if a.dtype.kind == "f":
        a = a.astype(np.int64)
    return np.ascontiguousarray((a.astype(np.int64) & 0xFF).astype(np.uint8))  # torch `.to(uint8)`: truncate / wrap


def pack(name, W_q):
    nbits = NBITS_OF[name]
    lv = _levels_u8(W_q)
    lv2 = lv.reshape(lv.shape[0], -1)
    rows, cols = lv2.shape
    prows = int(lib().hqq_oc_packed_rows(nbits, rows))
    out = np.empty((prows, cols), dtype=np.int32 if nbits == 3 else np.uint8)
    _check(lib().hqq_oc_pack(nbits, _p(lv2), rows, cols, _p(out)), "pack")
    return out.reshape((prows,) + lv.shape[1:])


def unpack(name, packed):
    nbits = NBITS_OF[name]
    p = np.ascontiguousarray(packed, dtype=np.int32 if nbits == 3 else np.uint8)
    p2 = p.reshape(p.shape[0], -1)
    f = 10 if nbits == 3 else 8 // nbits
    out = np.empty((f * p2.shape[0], p2.shape[1]), dtype=np.uint8)
    _check(lib().hqq_oc_unpack(nbits, _p(p2), p2.shape[0], p2.shape[1], _p(out)), "unpack")
    return out.reshape((f * p.shape[0],) + p.shape[1:])


PACK = {n: (lambda W, n=n: pack(n, W)) for n in NBITS_OF}
UNPACK = {n: (lambda W, n=n: unpack(n, W)) for n in NBITS_OF}


def quantize(tensor, nbits=4, group_size=64, optimize=True, round_zero=False, axis=0, lp_norm=0.7, beta=1e1, iters=20, return_trace=False):
    """Quantizer.quantize (quantize.py:76-180) -> (W_q packed, meta) like oracle.hqq_oracle.quantize."""
    W = np.ascontiguousarray(tensor, dtype=np.float32)
    N, K = W.shape
    G = N * K // group_size
    gshape = (G, group_size) if axis == 1 else (group_size, G)
    levels = np.empty(gshape, dtype=np.uint8)
    scale, zero = np.empty(G, dtype=np.float32), np.empty(G, dtype=np.float32)
    errors = np.zeros(max(iters, 1), dtype=np.float32)
    done = c_int(0)
    _check(lib().hqq_oc_quantize(_p(W), N, K, group_size, nbits, axis, int(round_zero), int(optimize), c_float(lp_norm), c_float(beta), iters,
                                 _p(levels), _p(scale), _p(zero), ctypes.byref(done), _p(errors)), "quantize")
    mshape = (G, 1) if axis == 1 else (1, G)
    meta = {"nbits": nbits, "group_size": group_size, "shape": (N, K), "scale": scale.reshape(mshape), "zero": zero.reshape(mshape),
            "axis": axis, "packing": BIT_TO_PACKING[nbits]}
    W_q = pack(meta["packing"], levels)
    if return_trace:
        return W_q, meta, {"iters": int(done.value), "errors": [float(e) for e in errors[: done.value]]}
    return W_q, meta


def dequantize(W_q, meta, compute_dtype="float32"):
    N, K = meta["shape"]
    nbits = meta["nbits"]
    Wq = np.ascontiguousarray(W_q, dtype=np.int32 if nbits == 3 else np.uint8)
    s = np.ascontiguousarray(meta["scale"], dtype=np.float32).reshape(-1)
    z = np.ascontiguousarray(meta["zero"], dtype=np.float32).reshape(-1)
    out = np.empty((N, K), dtype=np.float32)
    _check(lib().hqq_oc_dequantize(_p(Wq), _p(s), _p(z), N, K, meta["group_size"], nbits, meta["axis"], DTYPES[compute_dtype], _p(out)),
           "dequantize")
    return out


class Forward:
    """HQQBackend.PYTORCH at float32 for one layer (axis = 1): holds the N*K float scratch the reference re-materialises per call."""

    def __init__(self, W_q, meta):
        assert meta["axis"] == 1
        self.N, self.K = meta["shape"]
        self.nbits, self.gs = meta["nbits"], meta["group_size"]
        self.Wq = np.ascontiguousarray(W_q, dtype=np.int32 if self.nbits == 3 else np.uint8)
        self.s = np.ascontiguousarray(meta["scale"], dtype=np.float32).reshape(-1)
        self.z = np.ascontiguousarray(meta["zero"], dtype=np.float32).reshape(-1)
        self.W_r = np.empty(self.N * self.K, dtype=np.float32)

    def __call__(self, x, bias=None):
        x = np.ascontiguousarray(x, dtype=np.float32).reshape(-1, self.K)
        b = None if bias is None else np.ascontiguousarray(bias, dtype=np.float32)
        y = np.empty((x.shape[0], self.N), dtype=np.float32)
        _check(lib().hqq_oc_linear_forward_f32(_p(x), x.shape[0], _p(self.Wq), _p(self.s), _p(self.z), _p(b), self.N, self.K, self.gs,
                                               self.nbits, _p(self.W_r), _p(y)), "linear_forward_f32")
        return y


def linear_forward_f32(x, W_q, meta, bias=None):
    return Forward(W_q, meta)(x, bias)
