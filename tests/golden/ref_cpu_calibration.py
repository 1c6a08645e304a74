"""Build container only (needs /root/reference): time the REAL reference's CPU path (HQQLinear, HQQBackend.PYTORCH, float32,
torch.set_num_threads(all cores)) beside the C/OpenMP oracle port on the same layer, so that the port which bench.py times as the
CPU baseline on the GPU box (where the reference does not exist) is calibrated against the thing it stands in for.

    python tests/golden/ref_cpu_calibration.py            # prints one JSON line
"""
import json
import os
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("HQQ_REFERENCE", "/root/reference")
shim = tempfile.mkdtemp()
with open(os.path.join(shim, "termcolor.py"), "w") as f:
    f.write("def colored(s, *a, **k):\n    return s\n")
sys.path.insert(0, shim)
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import torch  # noqa: E402
from hqq.core.quantize import BaseQuantizeConfig, HQQBackend, HQQLinear  # noqa: E402

from oracle import hqq_oracle_c as C  # noqa: E402

torch.set_num_threads(os.cpu_count())


def best(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


def main():
    out = {"cores": os.cpu_count(), "torch_threads": torch.get_num_threads(), "c_threads": C.threads()}
    HQQLinear.set_backend(HQQBackend.PYTORCH)
    for n, k in ((4096, 4096), (14336, 4096)):
        torch.manual_seed(0)
        lin = torch.nn.Linear(k, n, bias=False)
        lin.weight.data = torch.randn(n, k) * 0.02
        W = lin.weight.data.numpy().copy()
        t0 = time.perf_counter()
        layer = HQQLinear(lin, BaseQuantizeConfig(nbits=4, group_size=64, axis=1), compute_dtype=torch.float32, device="cpu")
        t_ref_q = time.perf_counter() - t0
        t0 = time.perf_counter()
        Wq_c, meta_c = C.quantize(W, nbits=4, group_size=64, axis=1, round_zero=True)
        t_c_q = time.perf_counter() - t0
        x = torch.randn(1, k)
        with torch.no_grad():
            t_ref_f = best(lambda: layer(x), 5)
            y_ref = layer(x).numpy()
        meta = {"nbits": 4, "group_size": 64, "shape": (n, k), "axis": 1, "packing": "4bit_u8",
                "scale": layer.meta["scale"].float().numpy(), "zero": layer.meta["zero"].float().numpy()}
        fwd = C.Forward(layer.W_q.data.numpy(), meta)
        t_c_f = best(lambda: fwd(x.numpy()), 5)
        y_c = fwd(x.numpy())
        a, b = np.asarray(layer.W_q.data.numpy()), Wq_c
        out[f"{n}x{k}"] = {"reference_forward_ms": t_ref_f * 1e3, "c_port_forward_ms": t_c_f * 1e3, "forward_ratio_ref_over_port": t_ref_f / t_c_f,
                           "forward_rel_err": float(np.linalg.norm(y_c - y_ref) / np.linalg.norm(y_ref)),
                           "reference_quantize_s": t_ref_q, "c_port_quantize_s": t_c_q, "quantize_ratio_ref_over_port": t_ref_q / t_c_q,
                           "packed_bytes_differing": float((a != b).mean())}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
