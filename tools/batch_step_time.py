"""1 GPU: the lock-step batch decode step (bs = 32 leg of BASELINE configs[4]) on the batched glue kernels against the same step on
framework ops, CUDA-graph replays, Llama-3-8B-sized blocks (and the per-rank shapes of the 70B model at tp = 8 need 8 GPUs: bench.py)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import harness

layers = int(sys.argv[1]) if len(sys.argv) > 1 else 8
for batch in (8, 32):
    for fused in (True, False):
        m = harness.DecodeModel(harness.LLAMA3_8B, dtype=torch.float16, device="cuda", cache_len=128, fused=fused, batch=batch, n_layers=layers)
        m.capture()
        m.reset_state(1)
        for _ in range(5):
            m.decode()
        torch.cuda.synchronize()
        m.pos.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            m.decode()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 50
        print(f"batch {batch:2d} {'batched glue kernels' if fused else 'framework ops       '}: {ms:7.3f} ms/step with {layers} blocks = "
              f"{(ms * 1e3) / layers:6.1f} us per block (lm_head included), {batch / ms * 1e3:8.0f} tok/s", flush=True)
        del m
        torch.cuda.empty_cache()
