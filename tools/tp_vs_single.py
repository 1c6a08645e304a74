"""2+ GPU check: the tensor-parallel decode against the ONE-GPU model on the very same quantised weights.

Every rank builds the model with shard_from_full=True (full matrices from the shared generator, quantised unsharded, shards cut out
of the quantised tensors by hqq_b200/models/tp.py); rank 0 also builds the tp = 1 model from the same seed.  Both compute the same
function -- only the summation order of the row-parallel partials differs -- so the greedy tokens agree (a late near-tie may flip).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_vs_single.py
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import harness

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
shape = harness.LlamaShape(hidden=2048, inter=4096, n_layers=3, n_heads=16, n_kv_heads=8, vocab=4096)
STEPS = 16


def run(m):
    m.capture()
    m.reset_state(3)
    toks = []
    for _ in range(STEPS):
        m.decode()
        toks.append(int(m.next_tok))
    torch.cuda.synchronize()
    return toks


tp_model = harness.DecodeModel(shape, dtype=torch.float16, device=dev, cache_len=32, tp=world, rank=rank, process_group=dist.group.WORLD, seed=7,
                               shard_from_full=True)
tp_toks = run(tp_model)
dist.barrier()
if rank == 0:
    single = harness.DecodeModel(shape, dtype=torch.float16, device=dev, cache_len=32, tp=1, rank=0, seed=7, shard_from_full=True)
    one = run(single)
    agree = sum(int(a == b) for a, b in zip(one, tp_toks))
    print("single", one, flush=True)
    print(f"tp{world}   ", tp_toks, flush=True)
    print("AGREE", agree, "of", STEPS, "first4", one[:4] == tp_toks[:4], flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
