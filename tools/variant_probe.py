"""One experimental-kernel probe per process: run a fixed seeded workload under the environment knobs of THIS process, print
ONE JSON line {"what", "knobs", "us", "digest", ...} and optionally save the outputs for a tolerance comparison.

    python tools/variant_probe.py quant|l3|gemm|gemm_mid|decode [--save PATH]

bench.py launches it (default knobs, then a knob) in subprocesses with a timeout and compares digests / outputs; the kernels it
exercises were written after round 1's GPU budget was spent, so a crash or hang here must never reach the bench process.
`digest` is a sha256 over the raw output bytes: equal digests = bit-identical results."""
import hashlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hqq_b200 import ops  # noqa: E402
from hqq_b200.core.quantize import BaseQuantizeConfig, HQQLinear  # noqa: E402

DEV = torch.device("cuda", 0)


def digest(tensors):
    h = hashlib.sha256()
    for t in tensors:
        h.update(t.detach().contiguous().cpu().view(torch.uint8).numpy().tobytes())
    return h.hexdigest()[:16]


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def graph_timed(fn, reps):
    """Launch-bound pieces: capture `reps` calls in a CUDA graph, time one replay."""
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def probe_quant():
    """Quantizer.quantize on two Llama-3-8B matrices (fp16 source, 4-bit gs 64) + small odd cases for the digest."""
    g = torch.Generator(device=DEV)
    g.manual_seed(11)
    big = [(torch.randn(n, k, device=DEV, generator=g, dtype=torch.float32) * 0.02).half() for n, k in ((4096, 4096), (14336, 4096))]
    small = [((torch.randn(n, k, device=DEV, generator=g, dtype=torch.float32) * s).to(dt), nb, gs)
             for n, k, s, dt, nb, gs in ((1000, 512, 1.0, torch.float16, 4, 64), (330, 640, 0.02, torch.bfloat16, 3, 64),
                                         (256, 512, 2.0, torch.float32, 2, 32), (38, 64, 0.02, torch.float16, 4, 64))]
    outs = []
    for W in big:
        Wq, s, z, tr = ops.quantize(W, 4, 64, 1, True, True, want_trace=True)
        outs += [Wq, s, z, tr["info"], tr["errors"]]
    for W, nb, gs in small:
        Wq, s, z, tr = ops.quantize(W, nb, gs, 1, nb == 4, True, want_trace=True)
        outs += [Wq, s, z, tr["info"], tr["errors"]]
    us = timed(lambda: [ops.quantize(W, 4, 64, 1, True, True) for W in big], 3)
    weights = sum(W.numel() for W in big)
    return {"us": us, "digest": digest(outs), "gweights_per_s": weights / us / 1e3, "algorithmic_GBps": weights * 2.5625 / us / 1e3}, outs


def probe_l3():
    """3-bit layers at M = 1: HQQLinear.forward (route 0 = dequantize + GEMM, route 3 = csrc/linear3.cu)."""
    cfg = BaseQuantizeConfig(nbits=3, group_size=64, axis=1)
    torch.manual_seed(5)
    layers = [HQQLinear.from_weights((torch.randn(n, k, device=DEV) * 0.05).half(), None, cfg, compute_dtype=torch.float16, device=DEV)
              for n, k in ((4096, 4096), (11008, 4096), (4096, 11008), (1000, 1024), (33, 256))]
    xs = [torch.randn(1, l.meta["shape"][1], device=DEV).half() for l in layers]
    with torch.no_grad():
        outs = [l(x).clone() for l, x in zip(layers, xs)]
        us = graph_timed(lambda: [l(x) for l, x in zip(layers[:3], xs[:3])], 5) / 3
    return {"us": us, "digest": digest(outs), "route": ops.linear_route(1, 4096, 4096, 64, 3, 1, torch.float16)}, outs


def _gemm(Ms, shapes, reps):
    cfg = BaseQuantizeConfig(nbits=4, group_size=64, axis=1)
    torch.manual_seed(3)
    outs, per = [], {}
    for N, K in shapes:
        lin = HQQLinear.from_weights((torch.randn(N, K, device=DEV) * 0.02).half(), (torch.randn(N, device=DEV) * 0.1).half(), cfg,
                                     compute_dtype=torch.float16, device=DEV)
        for M in Ms:
            x = torch.randn(M, K, device=DEV).half()
            y = torch.empty(M, N, device=DEV, dtype=torch.float16)

            def run():
                return ops.linear_fwd(x, lin.W_q, lin.meta["scale"], lin.meta["zero"], lin.bias, N, K, 64, 4, 1, out=y)

            run()
            torch.cuda.synchronize()
            outs.append(y.clone())
            us = timed(run, reps)
            per[f"{N}x{K}xM{M}"] = {"us": round(us, 1), "TFLOPs": round(2.0 * M * N * K / us / 1e6, 1)}
    return per, outs


def probe_gemm():
    per, outs = _gemm([4096, 1000], [(4096, 4096), (11008, 4096)], 10)
    return {"us": per["4096x4096xM4096"]["us"], "per": per, "digest": digest(outs)}, outs


def probe_gemm_mid():
    per, outs = _gemm([64, 128, 256], [(4096, 4096), (4096, 11008)], 20)
    return {"us": per["4096x4096xM128"]["us"], "per": per, "digest": digest(outs)}, outs


def probe_decode():
    """Captured decode step of an 8-block Llama-3-8B-shaped stack (same kernels and launch order as bench.py)."""
    from hqq_b200 import harness
    m = harness.DecodeModel(harness.LLAMA3_8B, nbits=4, group_size=64, dtype=torch.float16, device=DEV, cache_len=64, n_layers=8)
    m.capture(warmup=3)
    m.tok.fill_(1)
    m.pos.zero_()
    for blk in m.blocks:
        blk["k_cache"].zero_(); blk["v_cache"].zero_()
    toks = []
    for _ in range(16):
        m.decode()
        toks.append(m.next_tok.clone())
    torch.cuda.synchronize()
    m.pos.fill_(20)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        m.decode()
    e1.record()
    torch.cuda.synchronize()
    outs = [torch.stack(toks)]
    return {"us": e0.elapsed_time(e1) * 1e3 / 30, "digest": digest(outs), "layers": 8}, outs


def probe_gemm_sweep():
    """BASELINE configs[2]: per-linear dequant-GEMM over (4096x4096, 11008x4096, 4096x11008) x nbits {8,4,3,2,1}, gs 64, fp16, at
    M = 4096 (tensor leg) and M = 128, default kernels.  `route` 2 = fused tcgen05 kernel, 0 = dequantize kernel + library GEMM
    (3-bit).  TFLOP/s = 2 M N K / time; `cublas_dequantised` times the library GEMM alone on the pre-dequantised fp16 matrix."""
    torch.manual_seed(0)
    per, outs = {}, []
    for nbits in (8, 4, 3, 2, 1):
        cfg = BaseQuantizeConfig(nbits=nbits, group_size=64, axis=1)
        for N, K in ((4096, 4096), (11008, 4096), (4096, 11008)):
            lin = HQQLinear.from_weights((torch.randn(N, K, device=DEV) * 0.02).half(), None, cfg, compute_dtype=torch.float16, device=DEV)
            Wd = lin.dequantize()
            for M in (4096, 128):
                x = torch.randn(M, K, device=DEV).half()
                y = torch.empty(M, N, device=DEV, dtype=torch.float16)
                route = ops.linear_route(M, N, K, 64, nbits, 1, torch.float16)

                def run():
                    if route != 0:
                        return ops.linear_fwd(x, lin.W_q, lin.meta["scale"], lin.meta["zero"], None, N, K, 64, nbits, 1, out=y)
                    with torch.no_grad():
                        return lin(x)

                r = run()
                torch.cuda.synchronize()
                ref = torch.matmul(x, Wd.t())
                err = float((r.float() - ref.float()).norm() / ref.float().norm())
                us = timed(run, 10)
                e = {"us": round(us, 1), "TFLOPs": round(2.0 * M * N * K / us / 1e6, 1), "route": route, "rel_err_vs_dequant_matmul": err}
                if M == 4096:
                    e["cublas_dequantised_TFLOPs"] = round(2.0 * M * N * K / timed(lambda: torch.matmul(x, Wd.t(), out=y), 10) / 1e6, 1)
                per[f"b{nbits}_{N}x{K}_M{M}"] = e
            del lin, Wd
    head = per["b4_4096x4096_M4096"]
    return {"us": head["us"], "per": per, "digest": "n/a"}, outs


def probe_bitpack():
    """pack / unpack / dequantize are HBM-bound streaming kernels.  One 14336 x 4096 matrix (58.7 M weights, gs 64, axis 1)
    per call through the C ABI with preallocated outputs, cycling over four input copies so that no call finds its input in the
    126 MB L2; algorithmic bytes = input + output (+ scale/zero) per call; event-timed over 24 back-to-back launches."""
    from hqq_b200._lib import DTYPE_CODE, check, load, ptr, stream_ptr
    lib, st = load(), stream_ptr(DEV)
    N, K, gs = 14336, 4096, 64
    R = N * K // gs
    g = torch.Generator(device=DEV)
    g.manual_seed(9)
    per, outs = {}, []
    for nbits in (4, 2, 8, 3, 1):
        f = 10 if nbits == 3 else 8 // nbits
        prow = -(-R // 10) if nbits == 3 else R // f
        levels = [torch.randint(0, 2 ** nbits, (R, gs), device=DEV, dtype=torch.uint8, generator=g) for _ in range(4)]
        packed = [torch.empty((prow, gs), device=DEV, dtype=torch.int32 if nbits == 3 else torch.uint8) for _ in range(4)]
        scale = (torch.rand(R, device=DEV, generator=g) * 0.01 + 1e-3).half()
        zero = (torch.rand(R, device=DEV, generator=g) * (2 ** nbits - 1)).half()
        unp = torch.empty((prow * f, gs), device=DEV, dtype=torch.uint8)
        deq = torch.empty((N, K), device=DEV, dtype=torch.float16)
        i = [0]

        def do_pack():
            j = i[0] = (i[0] + 1) % 4
            check(lib.hqq_b200_pack(nbits, ptr(levels[j]), DTYPE_CODE[torch.uint8], ptr(packed[j]), R, gs, st))

        def do_unpack():
            j = i[0] = (i[0] + 1) % 4
            check(lib.hqq_b200_unpack(nbits, ptr(packed[j]), ptr(unp), DTYPE_CODE[torch.uint8], prow, gs, st))

        def do_deq():
            j = i[0] = (i[0] + 1) % 4
            check(lib.hqq_b200_dequantize(ptr(packed[j]), ptr(scale), ptr(zero), ptr(deq), N, K, gs, nbits, 1, DTYPE_CODE[torch.float16], st))

        for _ in range(4):
            do_pack()
        torch.cuda.synchronize()
        assert torch.equal(ops.unpack(packed[0], nbits)[:R], levels[0])  # round trip through the public ops
        pbytes = packed[0].numel() * packed[0].element_size()
        for name, fn, nbytes in (("pack", do_pack, R * gs + pbytes), ("unpack", do_unpack, pbytes + prow * f * gs),
                                 ("dequantize_f16", do_deq, pbytes + 2 * N * K + 4 * R)):
            us = timed(fn, 24)
            per[f"b{nbits}_{name}"] = {"us": round(us, 2), "GBps": round(nbytes / us / 1e3, 1), "bytes": nbytes}
        del levels, packed, unp, deq
    return {"us": per["b4_dequantize_f16"]["us"], "per": per, "digest": "n/a", "shape": [N, K]}, outs


PROBES = {"bitpack": probe_bitpack, "gemm_sweep": probe_gemm_sweep, "quant": probe_quant, "l3": probe_l3, "gemm": probe_gemm, "gemm_mid": probe_gemm_mid, "decode": probe_decode}


def compare(ref, got):
    """max over outputs of ||a - b|| / ||a|| (0.0 = bit-identical everywhere)."""
    worst = 0.0
    for a, b in zip(ref, got):
        if a.shape != b.shape:
            return float("inf")
        if torch.equal(a, b):
            continue
        a, b = a.double(), b.double()
        worst = max(worst, float((a - b).norm() / a.norm().clamp_min(1e-30)))
    return worst


def main():
    what = sys.argv[1]
    res, outs = PROBES[what]()
    res["what"] = what
    res["knobs"] = {k: v for k, v in os.environ.items() if k.startswith("HQQ_B200_")}
    if "--both" in sys.argv:
        # knobs that are read on every call (HQQ_B200_SOLVER_VARIANT, HQQ_B200_FUSED_3BIT): default and variant in one process
        key, val = sys.argv[sys.argv.index("--both") + 1].split("=", 1)
        os.environ[key] = val
        try:
            res2, outs2 = PROBES[what]()
        finally:
            os.environ.pop(key, None)
        res = {"what": what, "knob": f"{key}={val}", "default": res, "variant": res2, "bit_identical": res["digest"] == res2["digest"],
               "rel_err": compare(outs, outs2), "speedup": res["us"] / res2["us"]}
    if "--save" in sys.argv:
        torch.save([o.cpu() for o in outs], sys.argv[sys.argv.index("--save") + 1])
    def finite(o):
        if isinstance(o, float):
            return o if o == o and abs(o) != float("inf") else None
        if isinstance(o, dict):
            return {k: finite(v) for k, v in o.items()}
        return o

    print("PROBE " + json.dumps(finite(res)), flush=True)


if __name__ == "__main__":
    main()
