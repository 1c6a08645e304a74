"""Debug/validation sweep of the fused small-M forward vs the oracle; prints one line per configuration."""
import itertools
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import hqq_oracle as o
from hqq_b200 import ops

DT = {"float16": torch.float16, "bfloat16": torch.bfloat16}
rows = []
for nbits, dtype, gs, (Nm, K), M in itertools.product([8, 4, 2, 1], ["float16", "bfloat16"], [64, 128],
                                                        [(16, 256), (16, 512), (24, 512), (32, 2048)], [1, 5, 20]):
    if nbits == 8 and dtype == "bfloat16":
        continue
    f = 8 // nbits
    N = Nm * f
    rng = np.random.RandomState(1)
    R = N * K // gs
    levels = rng.randint(0, 2 ** nbits, size=(R, gs))
    packing = o.BIT_TO_PACKING[nbits]
    W_q = o.PACK[packing](levels)
    scale = (rng.rand(R, 1) * 0.01 + 2e-3).astype(np.float32)
    zero = (rng.rand(R, 1) * (2 ** nbits - 1)).astype(np.float32)
    x = rng.randn(M, K).astype(np.float32)
    meta = {"nbits": nbits, "group_size": gs, "shape": (N, K), "axis": 1, "packing": packing, "scale": scale, "zero": zero}
    ref = o.linear_forward(x, W_q, meta, None, dtype)
    dt = DT[dtype]
    y = ops.linear_fwd(torch.from_numpy(x).cuda().to(dt), torch.from_numpy(W_q).cuda(), torch.from_numpy(scale).cuda().to(dt),
                       torch.from_numpy(zero).cuda().to(dt), None, N, K, gs, nbits, 1)
    if y is None:
        print(f"nbits={nbits} {dtype} gs={gs} N={N} K={K} M={M}: no route")
        continue
    y = y.float().cpu().numpy()
    err = np.linalg.norm(y - ref) / np.linalg.norm(ref)
    # per-row error pattern helps localise indexing bugs
    rowerr = np.abs(y - ref).max(axis=0) / (np.abs(ref).max() + 1e-9)
    bad = np.nonzero(rowerr > 0.02)[0]
    flag = "OK " if err < (2e-3 if dtype == "float16" else 1e-2) else "BAD"
    print(f"{flag} nbits={nbits} {dtype} gs={gs} N={N} K={K} M={M}: rel={err:.2e} badrows={bad[:12].tolist()}{'...' if len(bad) > 12 else ''}")
