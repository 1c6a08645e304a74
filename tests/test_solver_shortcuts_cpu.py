"""CPU: the two shortcuts of the fast solver kernel (hqq_b200/csrc/quantize.cu, solver_axis1_fast_kernel) are EXACT.

The kernel itself only runs on a GPU; what can be pinned here, on the oracle's float32 arithmetic, is the mathematics it relies on:
  (1) shrink_lp_op(x) == 0 exactly for |x| < thr, thr = 0.9 * beta^(-1/(2-p)) (p < 1) or 1/beta (p == 1)  (optimize.py:96-108);
  (2) one solver iteration is a function of the group's zero-point alone, so a bitwise-repeated zero means every later iteration
      repeats -- emulating the kernel's control flow (per-warp exit once all four groups are fixed, tail slots filled with the last
      zero / error) reproduces the plain `iters`-iteration trajectory and error sums bit for bit.
"""
import numpy as np
import pytest

from oracle import hqq_oracle as O

f32 = np.float32


def kernel_thr(beta, p):
    inv_beta = f32(1.0 / beta)
    if p == 1:
        return inv_beta
    if p < 1:
        return f32(0.9) * f32(float(inv_beta) ** (1.0 / (2.0 - p)))
    return f32(0.0)


@pytest.mark.parametrize("beta,p", [(10.0, 0.7), (10.0, 1.0), (2.0, 0.5), (100.0, 0.9), (1.0, 0.1), (10.0 * 1.01 ** 20, 0.7)])
def test_shrinkage_is_exactly_zero_below_threshold(beta, p):
    thr = kernel_thr(beta, p)
    assert thr > 0
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 200000).astype(f32) * thr,            # everything strictly inside
                        np.nextafter(thr, f32(0)) * np.array([1, -1], dtype=f32),  # the largest admissible magnitude
                        np.array([0.0, -0.0, 1e-45, -1e-45, 1e-38, 1e-30], dtype=f32)])
    x = x[np.abs(x) < thr]
    out = O.shrink_lp_op(x, beta, p)
    assert np.all(out == 0)
    # and the shortcut is not vacuous: the operator is non-zero a little above the root
    root = (1.0 / beta) ** (1.0 / (2.0 - p)) if p < 1 else 1.0 / beta
    assert O.shrink_lp_op(np.array([1.2 * root], dtype=f32), beta, p)[0] > 0


def plain_trajectory(W, s, z, maxv, beta, p, iters):
    """iters full iterations (what solver_axis1_kernel executes): zero-point history [iters+1, G] and per-group error sums."""
    hist, errs = [z.copy()], []
    for _ in range(iters):
        W_r, _, z = O.proximal_step(W, s, z, [0, maxv], beta, p, 1)
        errs.append(np.abs(W - W_r).astype(f32).sum(axis=1, dtype=f32))
        hist.append(z.copy())
    return np.stack(hist)[..., 0], np.stack(errs)


def fast_trajectory(W, s, z, maxv, beta, p, iters, groups_per_warp=4):
    """The kernel's control flow: per warp, iterate until every group's zero repeats bitwise, then fill the tail."""
    G = W.shape[0]
    hist = np.zeros((iters + 1, G), dtype=f32)
    errs = np.zeros((iters, G), dtype=f32)
    thr = kernel_thr(beta, p)
    executed = 0
    for g0 in range(0, G, groups_per_warp):
        sl = slice(g0, min(g0 + groups_per_warp, G))
        Ww, sw, zw = W[sl], s[sl], z[sl].copy()
        ws = (Ww * sw).astype(f32)
        hist[0, sl] = zw[:, 0]
        it, e_last = 0, None
        while it < iters:
            q = np.clip(np.round(ws + zw), f32(0), f32(maxv)).astype(f32)
            wr = ((q - zw) / sw).astype(f32)
            ad = np.abs(Ww - wr).astype(f32)
            if not (ad.max() < thr):   # warp-uniform fallback to the full formula
                _, _, znew = O.proximal_step(Ww, sw, zw, [0, maxv], beta, p, 1)
            else:
                znew = np.mean((q - ws).astype(f32), axis=1, keepdims=True, dtype=np.float64).astype(f32)  # the oracle's group mean
            e_last = ad.sum(axis=1, dtype=f32)
            errs[it, sl] = e_last
            hist[it + 1, sl] = znew[:, 0]
            fixed = np.all(znew.view(np.uint32) == zw.view(np.uint32))
            zw = znew
            it += 1
            executed += 1
            if fixed:
                break
        hist[it + 1:, sl] = zw[:, 0]
        errs[it:, sl] = e_last
    return hist, errs, executed / max(1, -(-G // groups_per_warp))


@pytest.mark.parametrize("nbits,std,p", [(4, 0.02, 0.7), (2, 0.02, 0.7), (8, 0.05, 0.7), (1, 0.02, 0.7), (3, 0.02, 0.7),
                                         (4, 1.0, 0.7),   # errors above the threshold: the fallback path runs
                                         (2, 2.0, 0.7), (4, 0.02, 1.0), (4, 1.0, 1.0)])
def test_fixed_point_exit_reproduces_the_full_trajectory(nbits, std, p):
    rng = np.random.default_rng(nbits * 7 + int(std * 100))
    Wt = (rng.standard_normal((96, 256)) * std).astype(f32)
    W, s, z, mm = O.quantize_init(Wt, nbits, 64, 1, round_zero=(nbits == 4))
    iters, beta = 20, 10.0
    h0, e0 = plain_trajectory(W, s, z, mm[1], beta, p, iters)
    h1, e1, mean_iters = fast_trajectory(W, s, z, mm[1], beta, p, iters)
    assert np.array_equal(h0.view(np.uint32), h1.view(np.uint32))
    assert np.array_equal(e0.view(np.uint32), e1.view(np.uint32))
    if std <= 0.05:
        assert mean_iters < 12  # the point of the exercise: far fewer than 20 iterations per warp on weight-like data
