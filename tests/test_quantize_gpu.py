"""GPU: hqq_b200_quantize (fused init + proximal solver + pack) against the reference (golden fixtures) and the
oracle.

Stated tolerances (the solver is a fixed-point iteration whose float32 rounding order decides
round-half ties, so bit-exactness is only required where the arithmetic is order-free):
  * optimize=False (init, round-half-even, clamp, packing, 1/scale): BIT-EXACT.
  * optimize=True, the reference's own fixtures: scale bit-exact, the SAME iteration count and EVERY level identical (what round 1's
    B200 run demonstrated on all 14 configurations); zero-points to 2e-6; error trajectory to 5e-5.  Comparisons against the
    oracle on other random matrices keep a +-1 iteration / <= 1 level / <= 5e-4 bound (the early stop compares float32 means that
    can tie to the last bit, and every implementation sums them in its own order).
"""
import hashlib

import numpy as np
import pytest
import torch

from hqq_b200 import ops
from hqq_b200.core.quantize import BaseQuantizeConfig, HQQLinear, Quantizer
from hqq_b200.core.optimize import optimize_weights_proximal

pytestmark = pytest.mark.gpu
DEV = "cuda"
COMBOS = [(nbits, axis, gs) for nbits in (8, 4, 3, 2, 1) for axis in (0, 1) for gs in ((64,) if nbits != 4 else (64, 32, 128))]


def _unpacked(oracle, nbits, W_q):
    return oracle.UNPACK[oracle.BIT_TO_PACKING[nbits]](W_q).astype(np.int32)


@pytest.mark.parametrize("nbits,axis,gs", COMBOS)
@pytest.mark.parametrize("src", [torch.float32])
def test_noopt_bit_exact_vs_reference(golden, nbits, axis, gs, src):
    q = golden.quant
    key = f"b{nbits}_a{axis}_g{gs}/noopt"
    W = torch.from_numpy(q["W"]).to(DEV).to(src)
    W_q, meta = Quantizer.quantize(W, nbits=nbits, group_size=gs, axis=axis, round_zero=(nbits == 4), optimize=False)
    assert np.array_equal(W_q.cpu().numpy(), q[key + "/W_q"])
    assert np.array_equal(meta["scale"].cpu().numpy(), q[key + "/scale"])
    assert np.array_equal(meta["zero"].cpu().numpy(), q[key + "/zero"])


@pytest.mark.parametrize("nbits,axis,gs", COMBOS)
def test_solver_vs_reference(golden, oracle, nbits, axis, gs):
    q = golden.quant
    key = f"b{nbits}_a{axis}_g{gs}"
    W = torch.from_numpy(q["W"]).to(DEV)
    W2d = W
    W_q, scale, zero, tr = ops.quantize(W2d, nbits=nbits, group_size=gs, axis=axis, round_zero=(nbits == 4), optimize=True, want_trace=True)
    iters = int(tr["info"][0])
    assert iters == int(q[key + "/iters"])
    np.testing.assert_allclose(tr["errors"].cpu().numpy()[:iters], q[key + "/errors"][:iters], rtol=5e-5)
    rows = q["W"].size // gs if axis == 1 else gs
    a = _unpacked(oracle, nbits, W_q.cpu().numpy())[:rows]
    b = _unpacked(oracle, nbits, q[key + "/W_q"])[:rows]
    assert np.array_equal(a, b)  # every level of the reference's W_q
    assert np.array_equal(scale.cpu().numpy().reshape(q[key + "/scale"].shape), q[key + "/scale"])
    zr = q[key + "/zero"]
    dz = np.abs(zero.cpu().numpy().reshape(zr.shape) - zr) / np.maximum(np.abs(zr), 1.0)
    assert np.median(dz) <= 2e-6


@pytest.mark.parametrize("src", [torch.float16, torch.bfloat16])
def test_half_sources_are_widened_exactly(oracle, src):
    """Model weights arrive as fp16/bf16; `tensor.float()` (quantize.py:102) is exact, so quantising the half tensor
    must equal quantising its float32 copy."""
    torch.manual_seed(3)
    W = (torch.randn(256, 512, device=DEV) * 0.02).to(src)
    a = ops.quantize(W, 4, 64, 1, True, True)
    b = ops.quantize(W.float(), 4, 64, 1, True, True)
    for x, y in zip(a[:3], b[:3]):
        assert torch.equal(x, y)


@pytest.mark.parametrize("axis,gs,shape", [(1, 24, (16, 96)), (0, 24, (48, 32)), (1, 512, (8, 1024)), (0, 128, (256, 16)), (1, 64, (2, 64))])
def test_generic_group_sizes_vs_oracle(oracle, axis, gs, shape):
    """Group sizes outside the register-resident fast paths (warp-per-group kernel), tiny tensors."""
    rng = np.random.RandomState(11)
    W = (rng.randn(*shape) * 0.05).astype(np.float32)
    for nbits in (4, 2):
        R = (W.size // gs) if axis == 1 else gs
        if R % (8 // nbits):
            continue
        Wq_o, meta_o, tr_o = oracle.quantize(W, nbits=nbits, group_size=gs, axis=axis, round_zero=(nbits == 4), return_trace=True)
        W_q, scale, zero, tr = ops.quantize(torch.from_numpy(W).to(DEV), nbits, gs, axis, nbits == 4, True, want_trace=True)
        assert abs(int(tr["info"][0]) - tr_o["iters"]) <= 1
        a, b = _unpacked(oracle, nbits, W_q.cpu().numpy()), _unpacked(oracle, nbits, Wq_o)
        assert np.abs(a - b).max() <= 1 and (a != b).mean() <= 2e-3
        assert np.array_equal(scale.cpu().numpy(), meta_o["scale"].reshape(-1))


def test_degenerate_groups(oracle):
    """Constant groups (|max-min| <= 1e-4 -> scale 1), huge ranges (scale clamp 2e4), zeros -- quantize.py:126-131."""
    W = np.zeros((8, 64), dtype=np.float32)
    W[1] = 3.25
    W[2] = np.linspace(-1e-6, 1e-6, 64)
    W[3] = np.linspace(-1e4, 1e4, 64)
    W[4, ::2] = 1e-3
    W[5] = -7.0
    W[6] = np.linspace(0, 1, 64)
    W[7, 0] = 100.0
    for optimize in (False, True):
        Wq_o, meta_o = oracle.quantize(W, nbits=4, group_size=64, axis=1, round_zero=True, optimize=optimize)
        W_q, scale, zero, _ = ops.quantize(torch.from_numpy(W).to(DEV), 4, 64, 1, True, optimize)
        assert np.array_equal(scale.cpu().numpy(), meta_o["scale"].reshape(-1))
        a, b = _unpacked(oracle, 4, W_q.cpu().numpy()), _unpacked(oracle, 4, Wq_o)
        assert np.abs(a - b).max() <= (0 if not optimize else 1)
        assert np.isfinite(zero.cpu().numpy()).all()


def test_optimize_weights_seam(golden, oracle):
    """Quantizer.optimize_weights contract (quantize.py:137-145): grouped fp32 W + inverse scale + zero in,
    (W_q float levels, scale, zero) out; result equals what the fused quantize path stores."""
    q = golden.quant
    W = q["W"]
    Wg, s0, z0, mm = oracle.quantize_init(W, 4, 64, 1, True)
    W_q, s, z = optimize_weights_proximal(tensor=torch.from_numpy(Wg).to(DEV), scale=torch.from_numpy(s0).to(DEV),
                                          zero=torch.from_numpy(z0).to(DEV), min_max=mm, axis=1, device="cuda")
    assert W_q.dtype == torch.float32 and tuple(W_q.shape) == Wg.shape and tuple(z.shape) == z0.shape
    assert torch.equal(s.cpu(), torch.from_numpy(s0))
    ref = _unpacked(oracle, 4, q["b4_a1_g64/W_q"])
    got = W_q.cpu().numpy().astype(np.int32)
    assert np.abs(got - ref).max() <= 1 and (got != ref).mean() <= 5e-4
    # a user-supplied solver is honoured (the seam is a class attribute, quantize.py:38)
    calls = []

    def my_solver(tensor, scale, zero, min_max, axis, device):
        calls.append(tensor.shape)
        return torch.round(tensor * scale + zero).clamp(min_max[0], min_max[1]), scale, zero

    old = Quantizer.optimize_weights
    try:
        Quantizer.optimize_weights = my_solver
        Wq2, meta2 = Quantizer.quantize(torch.from_numpy(W).to(DEV), nbits=4, group_size=64, axis=1, round_zero=True)
    finally:
        Quantizer.optimize_weights = old
    assert calls == [torch.Size([512, 64])]
    assert np.array_equal(Wq2.cpu().numpy(), q["b4_a1_g64/noopt/W_q"])


def test_config1_hqqlinear_vs_reference(golden, oracle):
    """BASELINE config 0: HQQLinear(nn.Linear(1024,1024), nbits=4, gs=64, axis=1); W from the reference tests' seed 42."""
    c = golden.config1
    torch.manual_seed(42)
    lin = torch.nn.Linear(1024, 1024)
    if hashlib.sha256(lin.weight.data.numpy().tobytes()).digest() != c["W_sha256"].tobytes():
        pytest.skip("torch RNG stream differs from the build container's")
    layer = HQQLinear(lin, BaseQuantizeConfig(nbits=4, group_size=64, axis=1), compute_dtype=torch.float16, device="cuda")
    assert layer.W_q.dtype == torch.uint8 and tuple(layer.W_q.shape) == (8192, 64) and layer.W_q.is_cuda
    assert layer.meta["scale"].dtype == torch.float16 and tuple(layer.meta["scale"].shape) == (16384, 1)
    a = _unpacked(oracle, 4, layer.W_q.data.cpu().numpy())
    b = _unpacked(oracle, 4, c["W_q"])
    assert np.abs(a - b).max() <= 1 and (a != b).mean() <= 1e-4
    assert np.array_equal(layer.meta["scale"].cpu().numpy(), c["scale"].astype(np.float16))
    x = torch.from_numpy(c["x"]).to(DEV).half()
    y = layer(x).float().cpu().numpy()
    ref = c["y/float16"]
    assert np.linalg.norm(y - ref) / np.linalg.norm(ref) <= 2e-3


@pytest.mark.parametrize("shape", [(4096, 4096), (14336, 4096)])
def test_full_size_properties(shape):
    """BASELINE sizes, size-independent properties: levels in range, reconstruction error at the level the reference
    reports for N(0,0.02^2) weights (1.5e-3 for 4-bit), determinism, solver beats plain rounding."""
    torch.manual_seed(0)
    W = (torch.randn(*shape, device=DEV) * 0.02).half()
    W_q, meta = Quantizer.quantize(W, nbits=4, group_size=64, axis=1, round_zero=True, compute_dtype=torch.float16)
    meta["compute_dtype"] = torch.float32
    W_r = Quantizer.dequantize(W_q, meta)
    err = (W.float() - W_r).abs().mean().item()
    assert 1.2e-3 < err < 1.8e-3
    W_q0, meta0 = Quantizer.quantize(W, nbits=4, group_size=64, axis=1, round_zero=True, optimize=False)
    meta0["compute_dtype"] = torch.float32
    lp = lambda a: (W.float() - a).abs().pow(0.7).mean().item()
    assert lp(W_r) < lp(Quantizer.dequantize(W_q0, meta0))  # the solver minimises the lp<1 norm it is built for
    W_q2, meta2 = Quantizer.quantize(W, nbits=4, group_size=64, axis=1, round_zero=True)
    assert torch.equal(W_q, W_q2) and torch.equal(meta["zero"], meta2["zero"])  # bit-reproducible (no atomics)


@pytest.mark.parametrize("nbits,gs,shape,std,src", [(4, 64, (1024, 4096), 0.02, torch.float16), (4, 64, (512, 1024), 1.0, torch.float16),
                                                    (2, 64, (512, 2048), 0.02, torch.bfloat16), (8, 128, (256, 1024), 0.05, torch.float16),
                                                    (3, 64, (250, 1024), 0.02, torch.float32), (1, 32, (512, 512), 0.02, torch.float16)])
@pytest.mark.parametrize("axis", [0, 1])
def test_register_solver_equals_plain_loop(monkeypatch, nbits, gs, shape, std, src, axis):
    """The shipped solver kernels (weights in registers, exact zero-shrinkage shortcut, exit at the fixed point) against the plain
    20-iteration loop of solver_generic_kernel (HQQ_B200_PLAIN_SOLVER=1) on the same GPU: same levels, scale, zero-points and
    iteration count.  (std 1.0: errors above the shrinkage threshold, the full formula runs.)  The two kernels sum the per-iteration
    error means in different fixed orders: should two consecutive means tie to the last float32 bit the stop may land one iteration
    apart -- then that, and nothing else, is asserted."""
    torch.manual_seed(nbits * 100 + gs + axis)
    W = (torch.randn(*shape, device=DEV) * std).to(src)
    fast = ops.quantize(W, nbits, gs, axis, nbits == 4, True, want_trace=True)
    monkeypatch.setenv("HQQ_B200_PLAIN_SOLVER", "1")
    plain = ops.quantize(W, nbits, gs, axis, nbits == 4, True, want_trace=True)
    monkeypatch.delenv("HQQ_B200_PLAIN_SOLVER")
    assert torch.equal(fast[1], plain[1])  # scale
    fi, pi = int(fast[3]["info"][0]), int(plain[3]["info"][0])
    torch.testing.assert_close(fast[3]["errors"], plain[3]["errors"], rtol=2e-6, atol=0)
    if fi == pi:
        assert torch.equal(fast[0], plain[0]) and torch.equal(fast[2], plain[2])
    else:
        assert abs(fi - pi) == 1
