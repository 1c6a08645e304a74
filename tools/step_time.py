"""Print the captured decode step time (us) of the Llama-3-8B-shaped model under the current env knobs -- one line, for sweeps."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import harness

dev = torch.device("cuda", 0)
model = harness.DecodeModel(harness.LLAMA3_8B, nbits=4, group_size=64, dtype=torch.float16, device=dev, cache_len=256,
                            n_layers=int(os.environ.get("LAYERS", "32")))
model.capture(warmup=3)
model.pos.fill_(100)
for _ in range(5):
    model.decode()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = int(os.environ.get("REPS", "100"))
model.pos.fill_(50)
e0.record()
for _ in range(reps):
    model.decode()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / reps
knobs = {k: v for k, v in os.environ.items() if k.startswith("HQQ_B200_")}
print(f"step {us:9.1f} us  {1e6 / us:7.1f} tok/s  next_tok {int(model.next_tok)}  {knobs}", flush=True)
