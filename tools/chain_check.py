"""GPU: hqq_b200_decode_linear_chain (o_proj -> [add+RMSNorm -> gate/up -> SiLU*mul] -> down_proj as one launch with grid barriers) against
the same three launches one by one, Llama-3-8B block shapes, many repetitions (the barrier words are reused).  Run under `timeout`:
a barrier bug shows as a hang."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import ops
from hqq_b200.core.quantize import BaseQuantizeConfig, HQQLinear

dev = torch.device("cuda", 0)
H, I = 4096, 14336
torch.manual_seed(0)
cfg = BaseQuantizeConfig(nbits=4, group_size=64, axis=1)
mk = lambda n, k: HQQLinear.from_weights((torch.randn(n, k, device=dev) * 0.02).half(), None, cfg, compute_dtype=torch.float16, device=dev)  # noqa: E731
o, g, u, d = mk(H, H), mk(I, H), mk(I, H), mk(H, I)
w = torch.rand(H, device=dev).half()
bar = torch.zeros(2, dtype=torch.int32, device=dev)
z = lambda n: torch.empty(1, n, device=dev, dtype=torch.float16)  # noqa: E731
ok = True
for rep in range(50):
    a = torch.randn(1, H, device=dev).half()
    h = torch.randn(1, H, device=dev).half()
    ref = [z(H), z(I), z(I), z(H), z(H)]
    assert ops.decode_linear_fwd(a, (o,), [ref[0]])
    assert ops.decode_linear_fwd(h, (g, u), [ref[1], ref[2]], 1 | ops.YOP_SILU_MUL_PAIR, ref[0], w, ref[4], 1e-5)
    assert ops.decode_linear_fwd(ref[1], (d,), [ref[3]])
    got = [z(H), z(I), z(I), z(H), z(H)]
    assert ops.decode_linear_chain([(a, (o,), [got[0]]), (h, (g, u), [got[1], got[2]], 1 | ops.YOP_SILU_MUL_PAIR, got[0], w, got[4], 1e-5),
                                    (got[1], (d,), [got[3]])], bar)
    torch.cuda.synchronize()
    same = all(torch.equal(x, y) for x, y in zip((ref[0], ref[1], ref[3], ref[4]), (got[0], got[1], got[3], got[4])))
    ok &= same
print("CHAIN", "IDENTICAL" if ok else "MISMATCH", "barrier words", bar.tolist())
