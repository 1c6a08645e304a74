"""GPU: BitPack through the C ABI.  Mirrors the reference's tests/test_bitpack.py (pack -> unpack identity over
its 11 shapes, 3 unpack dtypes, seed 42) and adds what the reference never pins: bit-exact agreement of the
packed bytes with an independent implementation (the oracle) and with fixtures produced by the reference."""
import numpy as np
import pytest
import torch

from hqq_b200.core.bitpack import BitPack

pytestmark = pytest.mark.gpu
DEV = "cuda"
SHAPES = [[32, 32], [128, 256], [256, 256], [512, 512], [1024, 1024], [2048, 2048], [4096, 4096], [8192, 8192], [8192, 4096],
          [8192, 128], [32, 4096]]
CASES = {8: (BitPack.pack_8bit_u8, BitPack.unpack_8bit_u8, "8bit_u8"), 4: (BitPack.pack_4bit_u8, BitPack.unpack_4bit_u8, "4bit_u8"),
         3: (BitPack.pack_3bit_32, BitPack.unpack_3bit_32, "3bit_32"), 2: (BitPack.pack_2bit_u8, BitPack.unpack_2bit_u8, "2bit_u8"),
         1: (BitPack.pack_1bit_u8, BitPack.unpack_1bit_u8, "1bit_u8")}


@pytest.mark.parametrize("nbits", [8, 4, 3, 2, 1])
def test_roundtrip_reference_loop(nbits):
    """tests/test_bitpack.py:24-34 -- W == unpack(pack(W))[:len(W)] for every shape / dtype (3 repeats)."""
    torch.manual_seed(42)
    pack, unpack, _ = CASES[nbits]
    for dtype in [torch.float16, torch.bfloat16, torch.float32]:
        for shape in SHAPES:
            for _ in range(3):
                W = torch.randint(0, 2 ** nbits, shape, device=DEV).contiguous()
                W_r = unpack(pack(W), dtype=dtype)
                assert W_r.dtype == dtype
                assert torch.equal(W, W_r[: len(W)].to(W.dtype))


@pytest.mark.parametrize("nbits", [8, 4, 3, 2, 1])
def test_golden_bytes(golden, nbits):
    """Bit-exact against the reference's own output (tests/golden/bitpack.npz)."""
    pack, unpack, name = CASES[nbits]
    g = golden.bitpack
    n = 0
    for k in g.files:
        if k.startswith(name) and k.endswith("/W"):
            si = k.split("/")[1]
            W = torch.from_numpy(g[k]).to(DEV)
            packed = pack(W)
            assert packed.dtype == (torch.int32 if nbits == 3 else torch.uint8)
            assert np.array_equal(packed.cpu().numpy(), g[f"{name}/{si}/packed"])
            assert np.array_equal(unpack(packed).cpu().numpy(), g[f"{name}/{si}/unpacked"])
            n += 1
    assert n >= 2


@pytest.mark.parametrize("nbits", [8, 4, 3, 2, 1])
@pytest.mark.parametrize("in_dtype", [torch.uint8, torch.int32, torch.int64, torch.float32, torch.float16, torch.bfloat16])
def test_against_oracle_ragged_and_dtypes(oracle, nbits, in_dtype):
    """Layout agreement with an independent implementation, incl. shapes that defeat the vector path
    (odd column counts, the 3-bit zero padding) and every input dtype Quantizer may hand to pack()."""
    pack, unpack, name = CASES[nbits]
    rng = np.random.RandomState(nbits)
    f = {8: 1, 4: 2, 3: 1, 2: 4, 1: 8}[nbits]
    for rows, cols in [(8 * f, 7), (16 * f, 64), (24 * f + (3 if nbits == 3 else 0), 33), (40 * f, 128)]:
        W = rng.randint(0, 2 ** nbits, size=(rows, cols))
        Wt = torch.from_numpy(W).to(DEV).to(in_dtype)
        packed = pack(Wt)
        ref = oracle.PACK[name](W)
        assert np.array_equal(packed.cpu().numpy(), ref), (rows, cols)
        assert np.array_equal(unpack(packed).cpu().numpy(), oracle.UNPACK[name](ref))


def test_empty_and_errors():
    assert BitPack.pack_4bit_u8(torch.zeros(0, 16, dtype=torch.uint8, device=DEV)).shape == (0, 16)
    assert BitPack.unpack_4bit_u8(torch.zeros(0, 16, dtype=torch.uint8, device=DEV)).shape == (0, 16)
    with pytest.raises(RuntimeError):  # ragged slab split fails in the reference too (bitpack.py:26-28)
        BitPack.pack_4bit_u8(torch.zeros(7, 16, dtype=torch.uint8, device=DEV))
    with pytest.raises(TypeError):
        BitPack.unpack_3bit_32(torch.zeros(4, 16, dtype=torch.uint8, device=DEV))


def test_cpu_tensors_are_staged_through_the_gpu():
    W = torch.randint(0, 16, (64, 32))
    out = BitPack.unpack_4bit_u8(BitPack.pack_4bit_u8(W))
    assert out.device.type == "cpu" and torch.equal(out.to(W.dtype), W)
