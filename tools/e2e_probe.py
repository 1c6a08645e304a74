"""2+ GPUs: where does the end-to-end loop (host token in, host token out, one host sync per step) lose time against the device
loop?  Per-step host timestamps for the fused exchange with the key exchange in the argmax launch, the same with an NCCL all-reduce
of the keys, with and without nvidia-smi sampling beside it.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/e2e_probe.py
"""
import os
import statistics
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from hqq_b200 import harness  # noqa: E402

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
STEPS = 200
layers = int(sys.argv[1]) if len(sys.argv) > 1 else 32
model = harness.DecodeModel(harness.LLAMA3_8B, dtype=torch.float16, device=dev, cache_len=256, tp=world, rank=rank, process_group=dist.group.WORLD,
                            n_layers=layers)
stream = torch.cuda.current_stream(dev)
h_in = torch.ones(1, dtype=torch.long).pin_memory()
h_out = torch.zeros(1, dtype=torch.long).pin_memory()


def loops(tag, sample):
    model.reset_state(1)
    for _ in range(5):
        model.decode()
    model.pos.zero_()
    torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(STEPS):
        model.decode()
    e1.record(stream)
    torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)
    dev_ms = e0.elapsed_time(e1) / STEPS
    sampler = bench.ClockSampler(lr) if (sample and rank == 0) else None
    if sampler:
        sampler.start()
    model.pos.zero_()
    for _ in range(3):
        model.tok.copy_(h_in, non_blocking=True); model.graph.replay(); h_out.copy_(model.next_tok, non_blocking=True); torch.cuda.synchronize(dev)
    model.pos.zero_()
    torch.cuda.synchronize(dev); dist.barrier(); torch.cuda.synchronize(dev)
    stamps = [time.perf_counter()]
    launch = []
    for _ in range(STEPS):
        a = time.perf_counter()
        model.tok.copy_(h_in, non_blocking=True)
        model.graph.replay()
        h_out.copy_(model.next_tok, non_blocking=True)
        launch.append(time.perf_counter() - a)
        stream.synchronize()
        h_in.copy_(h_out)
        stamps.append(time.perf_counter())
    if sampler:
        sampler.stop()
    d = sorted((b - a) * 1e3 for a, b in zip(stamps, stamps[1:]))
    print(f"[rank {rank}] {tag:34s} device loop {dev_ms:6.3f} ms/step | e2e per step: median {statistics.median(d):6.3f} mean {sum(d) / len(d):6.3f} "
          f"p90 {d[int(0.9 * len(d))]:6.3f} max {d[-1]:7.3f} ms, {sum(1 for v in d if v > 1.5 * statistics.median(d))} steps > 1.5 x median; "
          f"host launch part median {statistics.median(launch) * 1e3:6.3f} ms", flush=True)
    dist.barrier()


for head in ("p2p", "nccl"):
    os.environ["HQQ_B200_HEAD_EXCHANGE"] = head
    model.graph = None
    model.capture(warmup=3)
    for sample in (False, True):
        loops(f"head={head} nvidia-smi={'on' if sample else 'off'}", sample)
torch.cuda.synchronize()
sys.stdout.flush()
os._exit(0)
