"""TEST INFRASTRUCTURE ONLY: ctypes mirror of hqq_b200_decode_desc for the emulator runners (the package's own mirror lives in
hqq_b200/_lib.py and is checked against the C struct by tests/test_abi.py)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from hqq_b200._lib import DecodeDesc  # noqa: E402  (a ctypes.Structure: importing it loads no library)


def make_descs(phases, nbits, gs, dtype_code):
    """phases: dicts with x, layers (run_small.make_layer results), outs, K and optionally x_op, x2, x_weight, h_out, eps."""
    n = len(phases)
    descs = (DecodeDesc * n)()
    keep = []
    P = lambda a: a.ctypes.data if a is not None else None  # noqa: E731
    for i, ph in enumerate(phases):
        k = len(ph["layers"])
        VP = ctypes.c_void_p * k
        arrs = [VP(*[P(L["Wq"]) for L in ph["layers"]]), VP(*[P(L["scale"]) for L in ph["layers"]]), VP(*[P(L["zero"]) for L in ph["layers"]]),
                VP(*[P(L["bias"]) for L in ph["layers"]]), VP(*[P(o) for o in ph["outs"]]), (ctypes.c_int64 * k)(*[L["N"] for L in ph["layers"]])]
        keep.append(arrs)
        d = descs[i]
        d.x, d.x_op, d.x2, d.x_weight, d.h_out, d.eps, d.count = P(ph["x"]), ph.get("x_op", 0), P(ph.get("x2")), P(ph.get("x_weight")), P(ph.get("h_out")), ph.get("eps", 0.0), k
        d.W_q, d.scale, d.zero, d.bias, d.y, d.N = (ctypes.cast(a, ctypes.c_void_p) for a in arrs)
        d.K, d.group_size, d.nbits, d.dtype, d.tp, d.rank = ph["K"], gs, nbits, dtype_code, 1, 0
    return descs, keep
