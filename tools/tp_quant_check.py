"""2+ GPU check of sharded quantisation: every rank quantises ITS rows of a layer with the early stop taken from the
all-reduced error sums (ops.quantize_sharded) and must get exactly the shard that hqq_b200/models/tp.py cuts out of the unsharded
quantisation -- column-parallel (rows split) and row-parallel (groups of every row split), 4- and 2-bit.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_quant_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import ops
from hqq_b200.core.quantize import BaseQuantizeConfig, HQQLinear
from hqq_b200.models.tp import shard_bounds, shard_hqq_linear

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ok = True
for nbits in (4, 2):
    for N, K in ((1024, 2048), (4096, 4096)):
        g = torch.Generator(device=dev).manual_seed(100 + nbits + N)
        W = (torch.randn(N, K, device=dev, generator=g) * 0.02).half()          # the same full matrix on every rank
        full = HQQLinear.from_weights(W.clone(), None, BaseQuantizeConfig(nbits=nbits, group_size=64, axis=1), compute_dtype=torch.float16, device=str(dev))
        _, _, _, tr_full = ops.quantize(W, nbits, 64, 1, nbits == 4, True, want_trace=True)
        for parallel in ("column", "row"):
            ref = shard_hqq_linear(full, world, rank, parallel)
            if parallel == "column":
                n0, n1 = shard_bounds(N, world, rank)
                Ws = W[n0:n1].contiguous()
            else:
                j0, j1 = shard_bounds(K // 64, world, rank)
                Ws = W[:, j0 * 64:j1 * 64].contiguous()
            W_q, scale, zero, tr = ops.quantize_sharded(Ws, nbits, 64, 1, nbits == 4, want_trace=True)
            same = (torch.equal(W_q, ref.W_q.data) and torch.equal(scale.half().view(-1), ref.meta["scale"].view(-1))
                    and torch.equal(zero.half().view(-1), ref.meta["zero"].view(-1)) and int(tr["info"][0]) == int(tr_full["info"][0]))
            flag = torch.tensor([1 if same else 0], device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            ok &= bool(flag.item())
            if rank == 0:
                print(f"nbits={nbits} {N}x{K} {parallel}: {'IDENTICAL' if flag.item() else 'MISMATCH'} (iterations {int(tr['info'][0])} / unsharded {int(tr_full['info'][0])})", flush=True)
if rank == 0:
    print("SHARDED-QUANT", "ALL-IDENTICAL" if ok else "MISMATCH", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
