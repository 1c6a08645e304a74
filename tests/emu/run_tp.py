"""TEST INFRASTRUCTURE ONLY: the tensor-parallel tagged-word exchange of the one-token kernel (csrc/linear_small.cu,
hqq_b200_decode_linear_fwd_desc) on the emulator.  `tp` ranks live in ONE process as `tp` sets of host buffers; per exchange all
producers run first (row-parallel shard -> tagged words scattered into every rank's buffer), then all consumers (residual add of
the reduced partials + RMSNorm -> column-parallel shard), so no kernel ever polls for a word that is still to be written.  What is
checked is the data path the 8-GPU run relies on and no GPU test has exercised at tp = 8: slot indexing [parity][rank][n] on every
peer, tags/parities across consecutive exchanges, the fp32 reduction order, and the values fed to the next linear."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import build_emu  # noqa: E402
import run_small as R  # noqa: E402
from hqq_b200._lib import DecodeDesc  # noqa: E402  (the ctypes mirror of hqq_b200_decode_desc; importing it loads nothing)
from oracle import hqq_oracle as O  # noqa: E402

VP = ctypes.c_void_p


def ptrs(arrs):
    a = (VP * len(arrs))(*[x.ctypes.data if x is not None else None for x in arrs])
    return a, ctypes.cast(a, VP)


def run(tp, out):
    lib = ctypes.CDLL(build_emu.build())
    lib.hqq_b200_last_error.restype = ctypes.c_char_p
    H, Kr, Ng, nbits, steps, nblocks = 256, 256, 64, 4, 3, 2
    rng = np.random.default_rng(tp)
    f16 = lambda v: np.asarray(v, dtype=np.float32).astype(np.float16)  # noqa: E731
    row = [[R.make_layer(rng, H, Kr, nbits, 64) for _ in range(tp)] for _ in range(nblocks)]   # row-parallel shards [H, K/tp]
    col = [[R.make_layer(rng, Ng, H, nbits, 64) for _ in range(tp)] for _ in range(nblocks)]   # column-parallel shards [N/tp, H]
    norm_w = R.dev(rng.random(H).astype(np.float16))
    bufs = [R.aligned((2, tp, H), np.uint32) for _ in range(tp)]
    for b in bufs:
        b[...] = 0xFFFFFFFF
    peer_keep, peer_ptr = ptrs(bufs)
    step_ctr = R.aligned((1,), np.int32)
    h = [R.dev(f16(rng.standard_normal((1, H)))) for _ in range(tp)]      # every rank holds the same residual stream
    for r in range(1, tp):
        h[r][...] = h[0]
    res = {}
    for step in range(steps):
        step_ctr[0] = step
        for blk in range(nblocks):
            xs = [R.dev(f16(rng.standard_normal((1, Kr)))) for _ in range(tp)]
            part = [R.aligned((1, H), np.float16) for _ in range(tp)]
            # ---- producers: rank r scatters its partial of the row-parallel linear to every peer
            for r in range(tp):
                L = row[blk][r]
                keep = [ptrs([L["Wq"]]), ptrs([L["scale"]]), ptrs([L["zero"]]), ptrs([None]), ptrs([part[r]])]
                N = (ctypes.c_int64 * 1)(H)
                d = DecodeDesc(x=xs[r].ctypes.data, x_op=0, count=1, W_q=keep[0][1], scale=keep[1][1], zero=keep[2][1], bias=keep[3][1], y=keep[4][1],
                               N=ctypes.cast(N, VP), K=Kr, group_size=64, nbits=nbits, dtype=R.F16, tp=tp, rank=r, peer_data=peer_ptr,
                               step_ctr=step_ctr.ctypes.data, x_index=blk + 1, x_per_step=nblocks)
                # weight-prefetch spans for the launch that follows (a pure hint: the emulator ignores it, the descriptor must parse)
                nxt = col[blk][r]["Wq"]
                d.pf_ptr = (VP * 4)(nxt.ctypes.data, None, None, None)
                d.pf_bytes = (ctypes.c_int64 * 4)(nxt.nbytes & ~127, 0, 0, 0)
                assert lib.hqq_b200_decode_linear_fwd_desc(ctypes.byref(d), None) == 0, lib.hqq_b200_last_error()
            ex = step * nblocks + blk + 1
            tag, par = ex & 0xFFFF, ex & 1
            for r in range(tp):   # every rank's buffer now holds every rank's partial, tagged with this exchange
                for src in range(tp):
                    w = bufs[r][par, src]
                    assert np.all((w >> 16) == tag), (tp, step, blk, r, src)
                    assert np.array_equal((w & 0xFFFF).astype(np.uint16), part[src][0].view(np.uint16)), (tp, step, blk, r, src)
            # ---- consumers: delta = sum of the partials (fp32, rank order, rounded once), residual add, RMSNorm, next linear
            acc = np.zeros(H, dtype=np.float32)
            for src in range(tp):
                acc = (acc + part[src][0].astype(np.float32)).astype(np.float32)
            delta = f16(acc)
            t = f16(h[0][0].astype(np.float32) + delta.astype(np.float32))
            inv = np.float32(1.0 / np.sqrt(np.mean(t.astype(np.float32) ** 2, dtype=np.float32) + np.float32(1e-5)))
            xn = f16(f16(t.astype(np.float32) * inv).astype(np.float32) * norm_w.astype(np.float32))
            for r in range(tp):
                L = col[blk][r]
                y, hout = R.aligned((1, Ng), np.float16), R.aligned((1, H), np.float16)
                keep = [ptrs([L["Wq"]]), ptrs([L["scale"]]), ptrs([L["zero"]]), ptrs([None]), ptrs([y])]
                N = (ctypes.c_int64 * 1)(Ng)
                d = DecodeDesc(x=h[r].ctypes.data, x_op=1, x_weight=norm_w.ctypes.data, h_out=hout.ctypes.data, eps=1e-5, count=1, W_q=keep[0][1],
                               scale=keep[1][1], zero=keep[2][1], bias=keep[3][1], y=keep[4][1], N=ctypes.cast(N, VP), K=H, group_size=64, nbits=nbits,
                               dtype=R.F16, tp=tp, rank=r, red_data=bufs[r].ctypes.data, step_ctr=step_ctr.ctypes.data, x_index=blk + 1,
                               x_per_step=nblocks)
                assert lib.hqq_b200_decode_linear_fwd_desc(ctypes.byref(d), None) == 0, lib.hqq_b200_last_error()
                assert np.array_equal(hout[0], t), (tp, step, blk, r)      # the residual stream every rank carries forward
                ref = O.linear_forward(xn[None].astype(np.float32), L["Wq_host"], L["meta"], None, "float16")[0]
                err = float(np.linalg.norm(y[0].astype(np.float64) - ref) / np.linalg.norm(ref))
                assert err <= 3e-3, (tp, step, blk, r, err)
                res[f"tp{tp}_s{step}_b{blk}_r{r}"] = y.copy()
                h[r][...] = hout
    np.savez(out, **res)


if __name__ == "__main__":
    run(int(sys.argv[1]), sys.argv[2])
