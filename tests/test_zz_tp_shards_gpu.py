"""GPU: hqq_b200.models.tp.shard_hqq_linear -- tensor-parallel shards cut out of an already quantised HQQLinear with this package's
unpack / pack kernels.  Written after round 1's GPU budget was spent (the index arithmetic is covered on the CPU through the oracle,
tests/test_tp_shards_cpu.py)."""
import pytest
import torch

from hqq_b200.core.quantize import BaseQuantizeConfig, HQQLinear
from hqq_b200.models.tp import shard_bounds, shard_hqq_linear

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nbits", (4, 3, 2))
def test_shards_dequantise_to_slices_and_recombine(nbits):
    dev = "cuda:0"
    torch.manual_seed(nbits)
    N, K, tp = 512, 1024, 2
    lin = torch.nn.Linear(K, N, bias=True)
    layer = HQQLinear(lin, BaseQuantizeConfig(nbits=nbits, group_size=64, axis=1), compute_dtype=torch.float16, device=dev)
    full = layer.dequantize()
    x = torch.randn(1, K, device=dev).half()
    y_full = layer(x)
    cols, rows = [], []
    for rank in range(tp):
        c = shard_hqq_linear(layer, tp, rank, "column")
        n0, n1 = shard_bounds(N, tp, rank)
        assert torch.equal(c.dequantize(), full[n0:n1]) and torch.equal(c.bias, layer.bias[n0:n1])
        cols.append(c(x))
        r = shard_hqq_linear(layer, tp, rank, "row")
        k0, k1 = (v * 64 for v in shard_bounds(K // 64, tp, rank))
        assert torch.equal(r.dequantize(), full[:, k0:k1]) and (r.bias is not None) == (rank == 0)
        rows.append(r(x[:, k0:k1].contiguous()).float())
    # same rows, same kernel -- but a row's level sits in another bit field of the shard's bytes (the slab step changes), and the one-token
    # kernel plants the fields into fp16 lanes at different scales (1024 + q vs 1024 + 16 q): equal up to fp32 accumulation order
    y_col = torch.cat(cols, dim=1)
    assert (y_col.float() - y_full.float()).norm() / y_full.float().norm() <= 1e-3
    y_row = (rows[0] + rows[1]).half()
    assert (y_row.float() - y_full.float()).norm() / y_full.float().norm() <= 2e-3
    sd = shard_hqq_linear(layer, tp, 1, "column").state_dict()             # a shard serialises like any HQQLinear
    assert tuple(int(v) for v in sd["shape"]) == (N // tp, K)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_tensor_parallel_model_equals_the_one_gpu_model_on_the_same_quantised_weights():
    """tools/tp_vs_single.py: shards cut out of the unsharded quantisation -> TP = 2 decodes the one-GPU model's tokens."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29547", os.path.join(root, "tools", "tp_vs_single.py")], capture_output=True, text=True, timeout=300)
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("AGREE")]
    assert line, out.stdout[-2000:] + out.stderr[-2000:]
    agree = int(line[-1].split()[1])
    assert agree >= 14 and "first4 True" in line[-1], line[-1]
