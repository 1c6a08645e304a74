"""Quantize-only timing (BASELINE config 4 shapes): hqq_b200_quantize per layer shape, achieved algorithmic GB/s."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import ops

shapes = {"8b": [(4096, 4096), (1024, 4096), (14336, 4096), (4096, 14336)], "70b": [(8192, 8192), (1024, 8192), (28672, 8192), (8192, 28672)]}
which = sys.argv[1] if len(sys.argv) > 1 else "8b"
bits = [int(b) for b in sys.argv[2].split(",")] if len(sys.argv) > 2 else [4]
torch.manual_seed(0)
for nbits in bits:
    for N, K in shapes[which]:
        W = (torch.randn(N, K, device="cuda") * 0.02).half()
        for _ in range(2):
            out = ops.quantize(W, nbits, 64, 1, nbits == 4, True, want_trace=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            out = ops.quantize(W, nbits, 64, 1, nbits == 4, True, want_trace=True)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        nbytes = N * K * (2 + (nbits / 8 if nbits != 3 else 0.4)) + 2 * (N * K // 64) * 4
        print(f"nbits={nbits} N={N} K={K}: {ms * 1e3:9.1f} us  iters={int(out[3]['info'][0])}  {nbytes / ms / 1e6:7.1f} GB/s algorithmic"
              f"  {N * K / ms / 1e6:8.1f} Gweights/s", flush=True)
