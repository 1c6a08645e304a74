"""CPU: the work decomposition of the fused 3-bit one-token kernel (hqq_b200/csrc/linear3.cu) is a partition of the matrix.

3-bit packing puts ten slabs of rows in one int32 (bitpack.py:69-91): field f of packed row i is unpacked row i + f*step of the
[R, 64] group view, and step = ceil(R / 10) is not a multiple of the groups per output row, so the slabs cut output rows at ten
different offsets.  The kernel walks the packed words ONCE, in chunks of K words (one output row's worth per field); inside a
chunk every field contributes the tail of one output row and the head of the next, and every output row ends up as at most three
pieces (first / last / middle) that are written to three private fp32 slots and summed in slot order -- no atomics, one writer per
slot.  This test executes that decomposition in numpy on the oracle's packed layout and checks that it reproduces the oracle's
forward and never writes a slot twice."""
import numpy as np
import pytest

from oracle import hqq_oracle as O


def fused3_emulate(Wq, scale, zero, x, N, K, gs=64):
    Gk = K // gs
    R = N * Gk
    step = Wq.shape[0]
    assert step == -(-R // 10)
    y3 = np.zeros((N, 3), dtype=np.float64)
    written = np.zeros((N, 3), dtype=bool)
    X = x.reshape(Gk, gs).sum(1)
    for i0 in range(0, step, Gk):          # one chunk = Gk packed rows = K words
        for f in range(10):
            r0 = i0 + f * step
            nrows = min(Gk, step - i0, R - r0)
            if nrows <= 0:
                continue
            kbA, nA = r0 % Gk, r0 // Gk
            b = Gk - kbA                     # rows j >= b belong to output row nA + 1
            acc, cnt = [0.0, 0.0], [0, 0]
            for j in range(nrows):
                p = int(j >= b)
                kb = kbA + j if p == 0 else j - b
                r = r0 + j
                q = ((Wq[i0 + j] >> (27 - 3 * f)) & 7).astype(np.float64)
                acc[p] += scale[r] * float(q @ x[kb * gs:(kb + 1) * gs]) - scale[r] * zero[r] * X[kb]
                cnt[p] += 1
            for p in (0, 1):
                if not cnt[p]:
                    continue
                n = nA + p
                if p == 1 or kbA == 0:
                    slot = 0                 # the piece holding the row's first group
                elif kbA + cnt[0] == Gk:
                    slot = 1                 # the piece holding its last group
                else:
                    slot = 2                 # cut on both sides (slab boundary inside the row)
                assert not written[n, slot], (n, slot)
                written[n, slot] = True
                y3[n, slot] = acc[p]
    assert written[:, 0].all()               # every output row has a first piece
    return y3[:, 0] + y3[:, 1] + y3[:, 2], written


@pytest.mark.parametrize("N,K", [(16, 64), (10, 128), (33, 256), (7, 640), (128, 512), (50, 1024), (3, 64), (1, 128)])
def test_chunk_piece_slot_decomposition_matches_the_oracle(N, K):
    rng = np.random.default_rng(N * 1000 + K)
    gs, Gk = 64, K // 64
    R = N * Gk
    levels = rng.integers(0, 8, size=(R, gs))
    Wq = O.pack_3bit_32(levels)
    scale = (rng.random((R, 1)) * 0.01 + 2e-3).astype(np.float32)
    zero = (rng.random((R, 1)) * 7).astype(np.float32)
    x = rng.standard_normal(K).astype(np.float32)
    y, written = fused3_emulate(Wq, scale[:, 0].astype(np.float64), zero[:, 0].astype(np.float64), x.astype(np.float64), N, K)
    meta = {"nbits": 3, "group_size": gs, "shape": (N, K), "axis": 1, "packing": "3bit_32", "scale": scale, "zero": zero}
    ref = O.linear_forward(x[None], Wq, meta, None, "float32")[0]
    assert np.allclose(y, ref, rtol=1e-4, atol=1e-4)
    assert written.sum(1).max() <= 3 and written.sum() >= N
