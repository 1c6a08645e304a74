"""2+ GPU check of the tensor-parallel decode: peer-memory fused exchange vs NCCL all-reduce (same shards, same seeds)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import harness

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
shape = harness.LlamaShape(hidden=2048, inter=4096, n_layers=3, n_heads=16, n_kv_heads=8, vocab=4096)
res = {}
for mode in ("nccl", "p2p"):
    m = harness.DecodeModel(shape, dtype=torch.float16, device=dev, cache_len=32, tp=world, rank=rank, process_group=dist.group.WORLD, seed=7, tp_mode=mode)
    m.capture()
    m.tok.fill_(3); m.pos.zero_()
    for blk in m.blocks:
        blk["k_cache"].zero_(); blk["v_cache"].zero_()
    toks = []
    torch.cuda.synchronize(); dist.barrier()
    t0 = time.perf_counter()
    for _ in range(16):
        m.decode()
        toks.append(int(m.next_tok))
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 16
    res[mode] = (m.tp_mode, toks, dt)
    if rank == 0:
        print(mode, "->", m.tp_mode, toks, f"{dt * 1e6:.0f} us/token", flush=True)
if rank == 0:
    a, b = res["nccl"][1], res["p2p"][1]
    agree = sum(int(x == y) for x, y in zip(a, b))
    print("AGREE", agree, "of", len(a), "first4", a[:4] == b[:4], flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
