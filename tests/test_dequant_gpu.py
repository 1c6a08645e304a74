"""GPU: hqq_b200_dequantize against the reference's dequantised tensors (golden) and the oracle -- bit-exact in
float32, float16 and bfloat16 (the kernel keeps the reference's two roundings, quantize.py:198)."""
import numpy as np
import pytest
import torch

from hqq_b200 import ops
from hqq_b200.core.quantize import Quantizer

pytestmark = pytest.mark.gpu
DEV = "cuda"
DT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


def _to_np(t):
    return t.float().cpu().numpy()


@pytest.mark.parametrize("nbits", [8, 4, 3, 2, 1])
@pytest.mark.parametrize("axis", [0, 1])
@pytest.mark.parametrize("dtype", ["float32", "float16", "bfloat16"])
def test_golden_bit_exact(golden, nbits, axis, dtype):
    q = golden.quant
    key = f"b{nbits}_a{axis}_g64"
    dt = DT[dtype]
    W_q = torch.from_numpy(q[key + "/W_q"]).to(DEV)
    meta = {"nbits": nbits, "group_size": 64, "shape": torch.Size([128, 256]), "axis": axis,
            "packing": Quantizer.bit_to_packing[nbits], "view_as_float": False, "compute_dtype": dt,
            "scale": torch.from_numpy(q[key + "/scale"]).to(DEV).to(dt), "zero": torch.from_numpy(q[key + "/zero"]).to(DEV).to(dt)}
    W_r = Quantizer.dequantize(W_q, meta)
    assert W_r.dtype == dt and tuple(W_r.shape) == (128, 256)
    assert np.array_equal(_to_np(W_r), q[f"{key}/W_r/{dtype}"])


@pytest.mark.parametrize("nbits", [8, 4, 3, 2, 1])
@pytest.mark.parametrize("axis,gs,shape", [(1, 8, (24, 40)), (0, 8, (24, 40)), (1, 64, (40, 192)), (0, 64, (64, 104)), (1, 40, (16, 80))])
def test_oracle_ragged(oracle, nbits, axis, gs, shape):
    """Shapes that exercise the scalar fall-back of the vector kernel and the 3-bit padding."""
    rng = np.random.RandomState(7)
    N, K = shape
    total = N * K
    G = total // gs
    R, C = (G, gs) if axis == 1 else (gs, G)
    f = {8: 1, 4: 2, 3: 1, 2: 4, 1: 8}[nbits]
    if R % f:
        pytest.skip("reference cannot pack this many rows")
    levels = rng.randint(0, 2 ** nbits, size=(R, C))
    packing = oracle.BIT_TO_PACKING[nbits]
    W_q = oracle.PACK[packing](levels)
    mshape = (G, 1) if axis == 1 else (1, G)
    scale = (rng.rand(*mshape) * 0.01 + 1e-3).astype(np.float32)
    zero = (rng.rand(*mshape) * (2 ** nbits - 1)).astype(np.float32)
    for dname, dt in DT.items():
        meta_o = {"nbits": nbits, "group_size": gs, "shape": (N, K), "axis": axis, "packing": packing, "scale": scale, "zero": zero}
        ref = oracle.dequantize(W_q, meta_o, dname)
        out = ops.dequantize(torch.from_numpy(W_q).to(DEV), torch.from_numpy(scale).to(DEV), torch.from_numpy(zero).to(DEV),
                             (N, K), gs, nbits, axis, dt)
        assert np.array_equal(_to_np(out), ref), (dname,)


def test_view_as_float_is_the_same_bytes():
    """quantize.py:170-173,187-188: W_q may be stored viewed as compute_dtype; kernels reinterpret the pointer."""
    torch.manual_seed(0)
    W = torch.randn(128, 256, device=DEV) * 0.02
    for nbits in [8, 4, 3, 2, 1]:
        a, ma = Quantizer.quantize(W, nbits=nbits, group_size=64, axis=1, view_as_float=False)
        b, mb = Quantizer.quantize(W, nbits=nbits, group_size=64, axis=1, view_as_float=True, compute_dtype=torch.float16)
        assert b.dtype == torch.float16
        assert torch.equal(a, b.view(ma["unpack_view_dtype"]))
        assert torch.equal(Quantizer.dequantize(a, ma), Quantizer.dequantize(b, mb))


def test_full_size_matches_unpack_then_affine():
    """4096x4096 4-bit (BASELINE sweep size): fused dequantize == unpack kernel followed by torch's two-step affine map."""
    torch.manual_seed(1)
    N = K = 4096
    W_q = torch.randint(0, 256, (N * K // 64 // 2, 64), dtype=torch.uint8, device=DEV)
    scale = (torch.rand(N * K // 64, 1, device=DEV) * 0.01 + 1e-3).half()
    zero = (torch.rand(N * K // 64, 1, device=DEV) * 15).half()
    out = ops.dequantize(W_q, scale, zero, (N, K), 64, 4, 1, torch.float16)
    from hqq_b200.core.bitpack import BitPack
    ref = ((BitPack.unpack_4bit_u8(W_q, dtype=torch.float16) - zero) * scale).reshape(N, K)
    assert torch.equal(out, ref)


def test_hqq_aten_module_surface(oracle):
    """hqq/kernels/hqq_aten_cuda.cpp:57-73 names, served by the same library (tests/test_bitpack.py:37-47 pattern:
    python pack -> native unpack must agree)."""
    import hqq_b200.hqq_aten as hqq_aten
    from hqq_b200.core.bitpack import BitPack
    torch.manual_seed(42)
    for nbits, pack, unpack in ((4, BitPack.pack_4bit_u8, hqq_aten.unpack_4bit_u8), (2, BitPack.pack_2bit_u8, hqq_aten.unpack_2bit_u8),
                                (1, BitPack.pack_1bit_u8, hqq_aten.unpack_1bit_u8), (3, BitPack.pack_3bit_32, hqq_aten.unpack_3bit_32)):
        W = torch.randint(0, 2 ** nbits, (128, 256), device=DEV)
        assert torch.equal(W, unpack(pack(W))[: len(W)].to(W.dtype))
    W = torch.randn(128, 256, device=DEV) * 0.02
    for axis in (0, 1):
        W_q, meta = Quantizer.quantize(W, nbits=4, group_size=64, axis=axis, compute_dtype=torch.float16)
        s, z = meta["scale"].half(), meta["zero"].half()
        out = hqq_aten.dequantize(W_q, s, z, 128, 256, 64, 4, axis, "4bit_u8")
        meta["compute_dtype"] = torch.float16
        assert out.dtype == torch.float16 and torch.equal(out, Quantizer.dequantize(W_q, dict(meta, scale=s, zero=z)))
    Wg = torch.randint(0, 16, (64, 512), device=DEV)  # grouped [gs, C] matrix, axis-0 meta [1, C]
    s = torch.rand(1, 512, device=DEV).half(); z = (torch.rand(1, 512, device=DEV) * 15).half()
    out = hqq_aten.dequantize_4bit_u8(BitPack.pack_4bit_u8(Wg), s, z)
    assert torch.equal(out, (Wg.half() - z) * s)
