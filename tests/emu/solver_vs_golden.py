"""TEST INFRASTRUCTURE ONLY: the solver kernels (source on the emulator) against the reference fixtures, one line per configuration: mismatching levels, iteration counts, zero-point agreement.  python tests/emu/solver_vs_golden.py"""
import os, sys, ctypes, numpy as np
HERE = os.path.dirname(os.path.abspath(__file__)); ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, HERE); sys.path.insert(0, ROOT); os.chdir(ROOT)
import test_emu_cpu as T
import build_emu
from oracle import hqq_oracle as o
lib = ctypes.CDLL(build_emu.build())
lib.hqq_b200_quantize_workspace_bytes.restype = ctypes.c_size_t
lib.hqq_b200_quantize_workspace_bytes.argtypes = [ctypes.c_int64, ctypes.c_int64, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
lib.hqq_b200_last_error.restype = ctypes.c_char_p
g = np.load("tests/golden/quantize_small.npz")
W = g["W"]
tot = 0; n = 0
for variant in (0, 1):
    for nbits in (8, 4, 3, 2, 1):
        for axis in (0, 1):
            for gs in ((64,) if nbits != 4 else (64, 32, 128)):
                key = f"b{nbits}_a{axis}_g{gs}"
                Wq, s, z, info, err, _ = T.quantize(lib, W, T.F32, nbits, gs, variant, axis=axis)
                pk = o.BIT_TO_PACKING[nbits]
                rows = W.size // gs if axis == 1 else gs
                a = o.UNPACK[pk](Wq)[:rows].astype(int); b = o.UNPACK[pk](g[key + "/W_q"])[:rows].astype(int)
                zr = g[key + "/zero"].ravel()
                print(variant, key, "mismatch", int((a != b).sum()), "iters", int(info[0]), int(g[key + "/iters"]), "zero maxdiff", float(np.abs(z - zr).max()), "zero exact frac", float((z == zr).mean()))
                tot += int((a != b).sum())
print("TOTAL mismatching levels", tot)
