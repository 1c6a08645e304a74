"""Generate the golden fixtures in this directory by running the REAL reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

It imports mobiusml/hqq from /root/reference (read-only) with a two-line `termcolor`
stand-in (the reference imports termcolor at hqq/core/quantize.py:13; it is not installed
here), runs the reference's CPU / float32 / HQQBackend.PYTORCH path on seeded inputs and
stores inputs + outputs as compressed .npz files.  Nothing in tests/ imports the reference
at run time -- only these files.
"""
import hashlib
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HQQ_REFERENCE", "/root/reference")

shim = tempfile.mkdtemp()
with open(os.path.join(shim, "termcolor.py"), "w") as f:
    f.write("def colored(s, *a, **k):\n    return s\n")
sys.path.insert(0, shim)
sys.path.insert(0, REF)

import torch  # noqa: E402
from hqq.core.bitpack import BitPack  # noqa: E402
from hqq.core.quantize import Quantizer, HQQLinear, HQQBackend, BaseQuantizeConfig  # noqa: E402
from hqq.core import optimize as ref_opt  # noqa: E402

torch.set_num_threads(os.cpu_count())
DT = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}


def npy(t):
    if isinstance(t, torch.Tensor):
        if t.dtype == torch.bfloat16:
            return t.float().numpy()
        return t.detach().cpu().numpy()
    return np.asarray(t)


def solver_trace(W, nbits, group_size, axis, round_zero):
    """Re-run the reference loop (optimize.py:237-247) with its own step function to
    record the iteration count and the error trajectory."""
    Wf = W.float()
    Wg = Wf.reshape([-1, group_size]) if axis == 1 else Wf.reshape([group_size, -1])
    _min = Wg.min(axis=axis, keepdim=True)[0]
    _max = Wg.max(axis=axis, keepdim=True)[0]
    max_v = round(2 ** nbits - 1)
    denom = _max - _min
    scale = max_v / denom
    scale = torch.where(denom.abs() <= 1e-4, torch.full_like(scale, 1.0), scale)
    scale = scale.clamp(max=2e4)
    zero = -_min * scale
    if round_zero:
        zero = torch.round(zero)
    best = torch.tensor(torch.inf)
    errs = []
    for _ in range(20):
        W_r, W_q, zero, scale = ref_opt.optimize_weights_proximal_legacy_step(
            Wg, scale, zero, [0, max_v], 1e1, 0.7, axis)
        e = torch.abs(Wg - W_r).mean().float()
        errs.append(float(e))
        if e < best:
            best = e
        else:
            break
    return len(errs), np.asarray(errs, dtype=np.float64), npy(zero)


def gen_bitpack():
    rng = np.random.RandomState(42)
    out = {}
    cases = {"8bit_u8": (8, [(32, 32), (7, 24)]),
             "4bit_u8": (4, [(32, 32), (40, 16)]),
             "2bit_u8": (2, [(32, 32), (40, 16)]),
             "1bit_u8": (1, [(32, 32), (40, 16)]),
             "3bit_32": (3, [(32, 32), (40, 16), (23, 16)])}
    for name, (nbits, shapes) in cases.items():
        for si, shape in enumerate(shapes):
            W = rng.randint(0, 2 ** nbits, size=shape).astype(np.int32)
            packed = getattr(BitPack, "pack_" + name)(torch.from_numpy(W))
            unpacked = getattr(BitPack, "unpack_" + name)(packed)
            out[f"{name}/{si}/W"] = W
            out[f"{name}/{si}/packed"] = npy(packed)
            out[f"{name}/{si}/unpacked"] = npy(unpacked)
    np.savez_compressed(os.path.join(HERE, "bitpack.npz"), **out)


def gen_quantize_small():
    rng = np.random.RandomState(0)
    W = (rng.randn(128, 256) * 0.02).astype(np.float32)
    x = rng.randn(4, 256).astype(np.float32)
    out = {"W": W, "x": x}
    for nbits in [8, 4, 3, 2, 1]:
        for axis in [0, 1]:
            for gs in ([64] if nbits != 4 else [64, 32, 128]):
                rz = nbits == 4  # BaseQuantizeConfig: round_zero = (nbits == 4), quantize.py:1097
                W_q, meta = Quantizer.quantize(torch.from_numpy(W), nbits=nbits, group_size=gs, axis=axis,
                                               round_zero=rz, optimize=True, device="cpu")
                key = f"b{nbits}_a{axis}_g{gs}"
                iters, errs, zero_chk = solver_trace(torch.from_numpy(W), nbits, gs, axis, rz)
                assert np.array_equal(zero_chk, npy(meta["zero"])), key
                out[key + "/W_q"] = npy(W_q)
                out[key + "/scale"] = npy(meta["scale"])
                out[key + "/zero"] = npy(meta["zero"])
                out[key + "/iters"] = np.int32(iters)
                out[key + "/errors"] = errs
                # no-optimize variant pins the init + rounding path alone
                W_q0, meta0 = Quantizer.quantize(torch.from_numpy(W), nbits=nbits, group_size=gs, axis=axis,
                                                 round_zero=rz, optimize=False, device="cpu")
                out[key + "/noopt/W_q"] = npy(W_q0)
                out[key + "/noopt/scale"] = npy(meta0["scale"])
                out[key + "/noopt/zero"] = npy(meta0["zero"])
                if gs == 64:
                    for dname, dt in DT.items():
                        m = dict(meta)
                        m["compute_dtype"] = dt
                        Wq_d, m = Quantizer.to_ooplace(W_q, m, "cpu")
                        W_r = Quantizer.dequantize(Wq_d, m)
                        out[f"{key}/W_r/{dname}"] = npy(W_r)
                        if axis == 1:
                            y = torch.matmul(torch.from_numpy(x).to(dt), W_r.t())
                            out[f"{key}/y/{dname}"] = npy(y)
    np.savez_compressed(os.path.join(HERE, "quantize_small.npz"), **out)


def gen_quantize_heavy():
    """Heavy-tailed weights: |W - W_r| exceeds the shrinkage threshold beta^(-1/(2-p)) = 0.170, so the |x|^(p-1) branch of
    shrink_lp_op (optimize.py:96-108) -- identically zero on weight-like data such as quantize_small -- shapes the zero-points.
    Also known-answer vectors of shrink_lp_op itself (p = 0.7 and the p = 1 branch, with 0, tiny, threshold-sized, large inputs)."""
    rng = np.random.RandomState(5)
    W = (rng.standard_t(3, size=(64, 256)) * 1.5).astype(np.float32)
    out = {"W": W}
    for nbits in [4, 2, 1]:
        for axis in [0, 1]:
            rz = nbits == 4
            W_q, meta = Quantizer.quantize(torch.from_numpy(W), nbits=nbits, group_size=64, axis=axis, round_zero=rz, optimize=True, device="cpu")
            key = f"b{nbits}_a{axis}_g64"
            iters, errs, zero_chk = solver_trace(torch.from_numpy(W), nbits, 64, axis, rz)
            assert np.array_equal(zero_chk, npy(meta["zero"])), key
            out[key + "/W_q"] = npy(W_q)
            out[key + "/scale"] = npy(meta["scale"])
            out[key + "/zero"] = npy(meta["zero"])
            out[key + "/iters"] = np.int32(iters)
            out[key + "/errors"] = errs
    x = np.concatenate([[0.0, -0.0, 1e-30, -1e-30, 1e-6, 0.05, -0.05, 0.1699, 0.17, 0.1701, -0.1702, 0.2, -0.3, 1.0, -7.5, 100.0, 1e6],
                        rng.randn(200) * 0.2, rng.randn(200) * 3.0]).astype(np.float32)
    out["shrink/x"] = x
    out["shrink/p0.7_beta10"] = npy(ref_opt.shrink_lp_op(torch.from_numpy(x), 1e1, 0.7))
    out["shrink/p1_beta10"] = npy(ref_opt.shrink_lp_op(torch.from_numpy(x), 1e1, 1))
    out["shrink/p0.5_beta4"] = npy(ref_opt.shrink_lp_op(torch.from_numpy(x), 4.0, 0.5))
    np.savez_compressed(os.path.join(HERE, "quantize_heavy.npz"), **out)


def gen_quantize_degenerate():
    """Groups that hit the guards of the init (quantize.py:126-131): constant groups (|max - min| <= 1e-4 -> scale 1), ranges around
    the 1e-4 boundary, a range so small that the inverse scale is clamped to 2e4, huge ranges, all-zero groups, single outliers."""
    W = np.zeros((12, 64), dtype=np.float32)
    W[1] = 3.25
    W[2] = np.linspace(-1e-6, 1e-6, 64)
    W[3] = np.linspace(-1e4, 1e4, 64)
    W[4, ::2] = 1e-3
    W[5] = -7.0
    W[6] = np.linspace(0, 1, 64)
    W[7, 0] = 100.0
    W[8] = np.linspace(0, 1.0e-4, 64)        # denom == 1e-4 (float32): still "constant"
    W[9] = np.linspace(0, 1.01e-4, 64)       # just above: scale = 15 / 1.01e-4 > 2e4 -> clamped
    W[10] = np.linspace(-3e-4, 5e-4, 64)     # scale 18750: below the clamp
    W[11, 5] = -1e-30
    out = {"W": W}
    for nbits in (4, 2, 8):
        for optimize in (False, True):
            W_q, meta = Quantizer.quantize(torch.from_numpy(W), nbits=nbits, group_size=64, axis=1, round_zero=(nbits == 4), optimize=optimize,
                                           device="cpu")
            key = f"b{nbits}_opt{int(optimize)}"
            out[key + "/W_q"] = npy(W_q)
            out[key + "/scale"] = npy(meta["scale"])
            out[key + "/zero"] = npy(meta["zero"])
    np.savez_compressed(os.path.join(HERE, "quantize_degenerate.npz"), **out)


def gen_config1():
    """BASELINE config 0: single HQQLinear 1024x1024 nbits=4 gs=64 axis=1, PYTORCH backend, CPU."""
    torch.manual_seed(42)  # the reference tests' seed, tests/test_quantize.py:22
    lin = torch.nn.Linear(1024, 1024)
    W = lin.weight.data.clone()
    b = lin.bias.data.clone()
    HQQLinear.set_backend(HQQBackend.PYTORCH)
    out = {"W_sha256": np.frombuffer(hashlib.sha256(W.numpy().tobytes()).digest(), dtype=np.uint8),
           "W_probe": W[:4, :8].numpy().copy(), "bias": b.numpy()}
    torch.manual_seed(1)
    x = torch.randn(2, 1024)
    out["x"] = x.numpy()
    iters, errs, _ = solver_trace(W, 4, 64, 1, True)
    out["iters"], out["errors"] = np.int32(iters), errs
    for dname, dt in DT.items():
        layer = HQQLinear(torch.nn.Linear(1024, 1024), None, initialize=False) if False else None
        lin2 = torch.nn.Linear(1024, 1024)
        lin2.weight.data = W.clone()
        lin2.bias.data = b.clone()
        layer = HQQLinear(lin2, BaseQuantizeConfig(nbits=4, group_size=64, axis=1), compute_dtype=dt, device="cpu")
        if dname == "float32":
            out["W_q"] = npy(layer.W_q.data)
            out["scale"] = npy(layer.meta["scale"])
            out["zero"] = npy(layer.meta["zero"])
        with torch.no_grad():
            y = layer(x.to(dt))
        out["y/" + dname] = npy(y)
    np.savez_compressed(os.path.join(HERE, "config1.npz"), **out)


def gen_state_dict_keys():
    lin = torch.nn.Linear(64, 64)
    layer = HQQLinear(lin, BaseQuantizeConfig(nbits=4, group_size=64, axis=1), compute_dtype=torch.float32, device="cpu")
    sd = layer.state_dict()
    out = {}
    for k, v in sd.items():
        out["sd/" + k] = npy(v) if isinstance(v, torch.Tensor) else np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "state_dict.npz"), **out)


if __name__ == "__main__":
    gens = {"bitpack": gen_bitpack, "quantize_small": gen_quantize_small, "quantize_heavy": gen_quantize_heavy, "quantize_degenerate": gen_quantize_degenerate,
            "config1": gen_config1,
            "state_dict": gen_state_dict_keys}
    for name in (sys.argv[1:] or list(gens)):  # python make_golden.py [name ...] regenerates only the named fixtures
        gens[name]()
    for fn in sorted(os.listdir(HERE)):
        if fn.endswith(".npz"):
            print(fn, os.path.getsize(os.path.join(HERE, fn)))
