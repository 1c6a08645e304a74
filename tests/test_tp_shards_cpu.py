"""CPU: tensor-parallel shards of an already quantised layer (hqq_b200/models/tp.py) through the oracle's pack / unpack: the
shard's dequantised matrix IS the corresponding slice of the unsharded one (exactly), for every width incl. the padded 3-bit
packing, and column / row shards recombine to the unsharded forward."""
import numpy as np
import pytest

from hqq_b200.models import tp as TP


def _pack(oracle):
    return lambda lv, nbits: oracle.PACK[oracle.BIT_TO_PACKING[nbits]](lv)


def _unpack(oracle):
    return lambda wq, nbits: oracle.UNPACK[oracle.BIT_TO_PACKING[nbits]](wq)


@pytest.mark.parametrize("nbits", (8, 4, 3, 2, 1))
@pytest.mark.parametrize("parallel", ("column", "row"))
@pytest.mark.parametrize("tp", (1, 2, 4))
def test_shard_is_a_slice_of_the_unsharded_quantisation(oracle, nbits, parallel, tp):
    rng = np.random.default_rng(nbits * 10 + tp)
    N, K, gs = 64, 512, 64
    W = (rng.standard_normal((N, K)) * 0.02).astype(np.float32)
    W_q, meta = oracle.quantize(W, nbits=nbits, group_size=gs, axis=1, round_zero=(nbits == 4))
    full = oracle.dequantize(W_q, meta, "float32")
    x = rng.standard_normal((3, K)).astype(np.float32)
    y_full = oracle.linear_forward(x, W_q, meta, None, "float32")
    parts = []
    for rank in range(tp):
        Wq_s, meta_s = TP.shard_quantized(W_q, meta, tp, rank, parallel, _pack(oracle), _unpack(oracle))
        d = oracle.dequantize(Wq_s, meta_s, "float32")
        if parallel == "column":
            n0, n1 = TP.shard_bounds(N, tp, rank)
            assert meta_s["shape"] == (n1 - n0, K) and np.array_equal(d, full[n0:n1])
            parts.append(oracle.linear_forward(x, Wq_s, meta_s, None, "float32"))
        else:
            k0, k1 = (v * gs for v in TP.shard_bounds(K // gs, tp, rank))
            assert meta_s["shape"] == (N, k1 - k0) and np.array_equal(d, full[:, k0:k1])
            parts.append(oracle.linear_forward(x[:, k0:k1], Wq_s, meta_s, None, "float32"))
        assert meta_s["nbits"] == nbits and meta_s["packing"] == meta["packing"] and meta is not meta_s
    y = np.concatenate(parts, axis=1) if parallel == "column" else np.sum(parts, axis=0)
    if parallel == "column":
        assert np.array_equal(y, y_full)  # column shards: the very same dot products
    else:
        assert np.linalg.norm(y - y_full) / np.linalg.norm(y_full) <= 1e-6  # row shards: the all-reduce's summation order


def test_shard_argument_checks(oracle):
    W_q, meta = oracle.quantize(np.ones((8, 128), np.float32), nbits=4, group_size=64, axis=1)
    pk, up = _pack(oracle), _unpack(oracle)
    with pytest.raises(ValueError):
        TP.shard_quantized(W_q, meta, 3, 0, "column", pk, up)       # 8 rows do not split three ways
    with pytest.raises(ValueError):
        TP.shard_quantized(W_q, meta, 2, 2, "column", pk, up)       # rank out of range
    with pytest.raises(ValueError):
        TP.shard_quantized(W_q, meta, 4, 0, "row", pk, up)          # two groups per row do not split four ways
    with pytest.raises(ValueError):
        TP.shard_quantized(W_q, meta, 2, 0, "diagonal", pk, up)
    with pytest.raises(ValueError):
        TP.shard_quantized(W_q, dict(meta, axis=0), 2, 0, "column", pk, up)
    W_q2, meta2 = oracle.quantize(np.ones((8, 128), np.float32), nbits=2, group_size=64, axis=1)
    with pytest.raises(ValueError):
        TP.shard_quantized(W_q2, meta2, 8, 0, "column", pk, up)     # one row = two groups: not a whole byte of four 2-bit fields
    assert TP.shard_bounds(4096, 8, 3) == (1536, 2048)
