"""Launch the fused small-M forward a few times on Llama-3-8B linear shapes (for ncu captures / quick timing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import ops
from hqq_b200.core.quantize import BaseQuantizeConfig, HQQLinear

shapes = {"gate": (14336, 4096), "down": (4096, 14336), "q": (4096, 4096), "k": (1024, 4096)}
which = sys.argv[1].split(",") if len(sys.argv) > 1 else list(shapes)
M = int(sys.argv[2]) if len(sys.argv) > 2 else 1
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
torch.manual_seed(0)
cfg = BaseQuantizeConfig(nbits=4, group_size=64, axis=1)
for name in which:
    N, K = shapes[name]
    layers = [HQQLinear.from_weights((torch.randn(N, K, device="cuda") * 0.02).half(), None, cfg, compute_dtype=torch.float16, device="cuda")
              for _ in range(6)]
    x = torch.randn(M, K, device="cuda").half()
    out = torch.empty(M, N, device="cuda", dtype=torch.float16)
    for l in layers:
        ops.linear_fwd(x, l.W_q, l.meta["scale"], l.meta["zero"], None, N, K, 64, 4, 1, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for l in layers:
                ops.linear_fwd(x, l.W_q, l.meta["scale"], l.meta["zero"], None, N, K, 64, 4, 1, out=out)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(layers))
    nbytes = N * K * 0.5 + 2 * (N * K // 64) * 2
    print(f"{name} N={N} K={K} M={M}: {us:.2f} us  {nbytes / us / 1e3:.0f} GB/s", flush=True)

# fused launches: q+k+v and gate+up share the activation
for names in (("q", "k", "k"), ("gate", "gate")):
    if not set(names) <= set(which):
        continue
    K = shapes[names[0]][1]
    groups = [[HQQLinear.from_weights((torch.randn(shapes[n][0], K, device="cuda") * 0.02).half(), None, cfg, compute_dtype=torch.float16, device="cuda")
               for n in names] for _ in range(6)]
    x = torch.randn(M, K, device="cuda").half()
    outs = [torch.empty(M, shapes[n][0], device="cuda", dtype=torch.float16) for n in names]
    for gl in groups:
        ops.linear_fwd_multi(x, gl, outs)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            for gl in groups:
                ops.linear_fwd_multi(x, gl, outs)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(groups))
    nbytes = sum(shapes[n][0] * K * 0.5 + 2 * (shapes[n][0] * K // 64) * 2 for n in names)
    print(f"fused {'+'.join(names)} K={K} M={M}: {us:.2f} us  {nbytes / us / 1e3:.0f} GB/s", flush=True)
