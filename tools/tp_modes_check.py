"""2+ GPUs: bench.py's own N > 1 check (`tokens_agree`: the fused NVLink exchange against NCCL all-reduces, teacher-forced, with the
logit differences) on shapes whose PER-RANK matrices are those of larger tensor-parallel degrees -- e.g. Llama-3-8B at tp = 8 has
o_proj K = 512 and down_proj K = 1792 per rank (not a multiple of 512: the one-token kernel's register-metadata variant), which two
GPUs reproduce with hidden 1024 / 8 heads / inter 3584.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/tp_modes_check.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hqq_b200 import harness  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
S = harness.LlamaShape
SHAPES = {
    "per-rank shapes of 8B at tp=8 (o K=512, down K=1792)": S(hidden=512 * world, inter=1792 * world, n_layers=8, n_heads=4 * world, n_kv_heads=world, vocab=16032 * world),
    "per-rank shapes of 8B at tp=4 (o K=1024, down K=3584)": S(hidden=1024 * world, inter=3584 * world, n_layers=8, n_heads=8 * world, n_kv_heads=2 * world, vocab=16032 * world),
    "Llama-3-8B, 6 blocks": S(hidden=4096, inter=14336, n_layers=6, n_heads=32, n_kv_heads=8, vocab=128256),
}
bad = 0
for name, shape in SHAPES.items():
    for mode in ("p2p", "nccl"):
        m = harness.DecodeModel(shape, dtype=torch.float16, device=dev, cache_len=64, tp=world, rank=rank, process_group=dist.group.WORLD, seed=11, tp_mode=mode)
        m.capture(warmup=2)
        agree, detail = bench.tokens_agree(m, torch, n_tokens=24)
        bad += 0 if agree else 1
        if rank == 0:
            print(json.dumps({"shape": name, "timed_mode": mode, "agree": agree, **detail}), flush=True)
        del m
        torch.cuda.empty_cache()
        dist.barrier()
if rank == 0:
    print("TP_MODES_CHECK", "OK" if bad == 0 else f"{bad} FAILED", flush=True)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0 if bad == 0 else 1)
