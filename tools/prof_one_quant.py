"""One 14336x4096 fp16 matrix through hqq_b200_quantize a few times (for ncu captures of the solver / round+pack kernels)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import ops

torch.manual_seed(0)
N, K = 14336, 4096
W = (torch.randn(N, K, device="cuda") * 0.02).half()
for _ in range(4):
    out = ops.quantize(W, 4, 64, 1, True, True)
torch.cuda.synchronize()
print("ok", int(out[0].sum()))
