#!/usr/bin/env python
"""bench.py -- Llama-3-8B-shaped 4-bit (gs=64, axis=1) decode tokens/s on B200 through hqq_b200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

A "step" is one decoded token (bs=1, seq=1) through all 32 blocks + lm_head of a random-init Llama-3-8B-shaped
stack whose 224 block linears are HQQLinear layers quantised on the GPU by this package (synthetic data, BASELINE.json
configs[1]).  `value` is tokens/s with the token fed back on the device (inputs resident in HBM); `e2e` is the same loop
driven from the host through the public API: every step copies the input token from pinned host memory, replays the
decode graph and reads the produced token back.  For N > 1 the same model is tensor-parallel over N GPUs (column-sharded
q/k/v/gate/up, row-sharded o/down whose partial sums are exchanged inside the kernels over NVLink peer memory -- NCCL
all-reduce with HQQ_B200_TP_MODE=nccl), i.e. strong scaling.

`--impl reference` times the reference algorithm's CPU implementation (the oracle port of HQQBackend.PYTORCH:
dequantise -> matmul per linear) on this box's host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "llama3_8b_4bit_gs64_decode_tokens_per_s"
UNIT = "tokens/s"
# BASELINE.json configs[1]; both arms print exactly this string as config.workload
WORKLOAD = "Llama-3-8B-shaped decode bs=1 seq=1, 32 blocks x 7 HQQLinear 4-bit gs=64 axis=1, fp16 lm_head (BASELINE configs[1])"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": float(d["hbm_gbs"]), "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------- CPU baseline (oracle port)
def _cpu_oracle():
    """(forward factory, cores, label): the C/OpenMP restatement (oracle/hqq_oracle_c.c, all host threads) when it builds here,
    else the numpy port (element-wise passes single-threaded, BLAS matmul).  bench.py's cpu_baseline / --impl reference legs are
    the only product-side places that may execute oracle/ (it is the thing timed here, never the thing shipped)."""
    try:
        from oracle import hqq_oracle_c as c
        cores = c.threads()
        return (lambda W_q, meta: c.Forward(W_q, meta)), cores, f"C/OpenMP port (oracle/hqq_oracle_c.c, {cores} threads)", c
    except Exception:  # noqa: BLE001 -- no C compiler / no OpenMP: the numpy port
        from oracle import hqq_oracle as o
        return (lambda W_q, meta: (lambda x: o.linear_forward_f32_fast(x, W_q, meta))), 1, "numpy port (oracle/hqq_oracle.py; BLAS matmul may use more threads)", None


class CpuReference:
    """HQQBackend.PYTORCH on the host: per linear, dequantise the whole matrix (unpack, subtract, multiply: three passes over an
    N x K float32 matrix) then matmul (quantize.py:184-199, 880-898), float32 compute dtype (the reference's CPU path), through the
    oracle port.  One `sample()` = a bounded number of passes over ONE of the 32 blocks (7 linears, bs=1) plus 1/8 of the fp32
    lm_head GEMV; the per-token figure extrapolates x32 blocks."""

    SHAPES = {"q": (4096, 4096), "k": (1024, 4096), "v": (1024, 4096), "o": (4096, 4096), "gate": (14336, 4096), "up": (14336, 4096),
              "down": (4096, 14336)}

    def __init__(self):
        import numpy as np
        make, self.cores, self.label, _ = _cpu_oracle()
        rng = np.random.RandomState(0)
        self.layers = []
        for name, (n, k) in self.SHAPES.items():
            R = n * k // 64
            W_q = rng.randint(0, 256, size=(R // 2, 64)).astype(np.uint8)
            meta = {"nbits": 4, "group_size": 64, "shape": (n, k), "axis": 1, "packing": "4bit_u8",
                    "scale": (rng.rand(R, 1) * 0.01 + 1e-3).astype(np.float32), "zero": (rng.rand(R, 1) * 15).astype(np.float32)}
            self.layers.append((make(W_q, meta), rng.randn(1, k).astype(np.float32)))
        self.lm = rng.randn(16032, 4096).astype(np.float32)  # 1/8 of the 128256-row fp32 lm_head
        self.xv = rng.randn(4096).astype(np.float32)
        self.block()  # untimed pass: page the scratch matrices in

    def block(self):
        for f, x in self.layers:
            f(x)

    def sample(self, budget_s: float = 6.0):
        t0 = time.perf_counter()
        reps = 0
        while True:
            self.block()
            reps += 1
            el = time.perf_counter() - t0
            if el > budget_s or reps >= 64 or (reps >= 3 and el > min(budget_s, 4.0)):
                break
        block_s = (time.perf_counter() - t0) / reps
        t1 = time.perf_counter()
        for _ in range(3):
            self.lm @ self.xv
        lm_s = (time.perf_counter() - t1) / 3 * 8
        tok_s = 1.0 / (block_s * 32 + lm_s)
        return tok_s, {"block_s": block_s, "lm_head_s": lm_s, "cores": self.cores, "port": self.label,
                       "sample": f"{reps}x one block (7 HQQ linears, dequantise+matmul, fp32, bs=1) + 1/8 lm_head; x32 blocks extrapolated; {self.label}"}


def cpu_reference_tokens_per_s(budget_s: float = 15.0):
    return CpuReference().sample(budget_s)


def cpu_quantizer_baseline():
    """Quantizer.quantize on the host cores through the C/OpenMP oracle port: ONE 4096 x 4096 matrix of the quantizer object's
    workload (same distribution, float32 as the reference's CPU path computes), in G weights/s like `quantizer.gweights_per_s`."""
    import numpy as np
    _, cores, label, c = _cpu_oracle()
    if c is None:
        return {"error": "C oracle unavailable (no compiler)"}
    W = (np.random.RandomState(7).randn(4096, 4096) * 0.02).astype(np.float16).astype(np.float32)
    t0 = time.perf_counter()
    _, _, tr = c.quantize(W, nbits=4, group_size=64, axis=1, round_zero=True, optimize=True, return_trace=True)
    dt = time.perf_counter() - t0
    return {"value": W.size / dt / 1e9, "unit": "Gweights/s", "cores": cores, "kind": "port", "seconds": dt, "solver_iterations": tr["iters"],
            "sample": f"one 4096x4096 matrix (of the block's seven), min/max + proximal solver + round + pack; {label}"}


def run_reference(args, rank, world):
    """--impl reference: rank 0 times the reference's CPU path (the oracle port, every host thread it can use) on bounded samples
    of the workload; at most 1 warm-up and 3 timed samples so that the default --steps/--warmup finish within a minute or two."""
    if rank != 0:
        return
    ref = CpuReference()
    warm, steps = min(args.warmup, 1), max(1, min(args.steps, 3))
    vals, info = [], None
    for i in range(warm + steps):
        v, info = ref.sample(budget_s=6.0)
        if i >= warm:
            vals.append(v)
    value = statistics.median(vals)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals),
            "warmup": warm, "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "path": "HQQBackend.PYTORCH data flow (dequantise + matmul) on the host cores: " + info["port"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": info["cores"], "host_cores": os.cpu_count() or 1, "kind": "port",
                             "sample": info["sample"]},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- GPU arm
def kernel_roofline(model, torch, peaks, reps=4):
    """Average duration of the fused forward kernel for each of the four launch groups a decode step issues per block
    (q+k+v, o, gate+up, down -- the matrices that share an activation go out in ONE launch).  All launches of one group over
    all layers (x reps) are captured into a CUDA graph so the measurement is not bound by Python launch overhead, replayed,
    and timed with CUDA events on the launching stream.  Cycling through every layer's weights means each launch streams cold
    weights (per-group footprint x 32 layers exceeds the 126 MB L2).
    achieved = algorithmic bytes of the 128 launches (224 matrices) of one step / their summed average durations."""
    from hqq_b200 import ops
    dev = model.device
    groups = [("qkv", ("q", "k", "v")), ("o", ("o",)), ("gate_up", ("gate", "up")), ("down", ("down",))]
    per = {}
    tot_bytes = tot_ms = 0.0
    for gname, names in groups:
        sets = [[blk[n] for n in names] for blk in model.blocks]
        K = sets[0][0].meta["shape"][1]
        Ns = [l.meta["shape"][0] for l in sets[0]]
        x = torch.randn(1, K, device=dev).to(model.dtype)
        outs = [torch.empty(1, N, device=dev, dtype=model.dtype) for N in Ns]

        # the MLP launch ships with the silu*mul epilogue (gate and up rows paired per tile, one output vector)
        x_op = ops.YOP_SILU_MUL_PAIR if (gname == "gate_up" and model.pair_silu and model.nbits < 8) else 0

        def run_all():
            for ls in sets:
                if not ops.decode_linear_fwd(x, ls, outs, x_op):
                    ops.linear_fwd_multi(x, ls, outs)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            run_all()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(reps):
                run_all()
        g.replay()
        torch.cuda.synchronize(dev)
        stream = torch.cuda.current_stream(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        g.replay()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1) / (reps * len(sets))
        nbytes = sum(N * K * 0.5 + 2 * (N * K // 64) * 2 for N in Ns) + (Ns[0] if x_op else sum(Ns)) * 2 + K * 2
        per[gname] = {"N": Ns, "K": K, "us": round(ms * 1e3, 3), "GBps": round(nbytes / ms / 1e6, 1)}
        tot_bytes += nbytes
        tot_ms += ms
    achieved = tot_bytes / tot_ms / 1e6
    # dram bytes per launch (average over the four launch groups) from the committed `ncu --set full` capture of these kernels
    traffic = None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r1_decode1_traffic.json")) as fh:
            traffic = json.load(fh)["traffic_bytes_per_launch_avg"]
    except (OSError, KeyError, ValueError):
        pass
    return {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "traffic": traffic, "algorithmic_bytes_per_launch": tot_bytes / len(groups), "kernel": "hqq::linear_decode1_kernel<half,4,64> (4 launches/block, 128/step; D1_VARIANT=" + os.environ.get("HQQ_B200_D1_VARIANT", "0") + ")", "peak_source": peaks["source"],
            "per_launch_group": per, "linear_us_per_step": round(tot_ms * 1e3 * len(model.blocks), 1),
            "note": "event-timed graph replay of back-to-back launches over all layers (cold weights); traffic: profiles/r1_decode1_traffic.json (ncu dram bytes)"}


def quantizer_roofline(torch, peaks, dev, reps=3, fast_ok=False):
    """North-star path (a): Quantizer.quantize (min/max init + proximal solver + round + pack) of ONE Llama-3-8B block's seven
    matrices (218 M weights, fp16 source, 4-bit gs=64 axis=1, 20 iterations max), timed with CUDA events on the launching stream.
    Algorithmic bytes: N*K*(2 + 0.5) + 2*(N*K/64)*4 per matrix.  Reported next to the decode roofline; the solver is
    instruction-bound by construction, so the HBM fraction is a ceiling statement, not a tuning target."""
    from hqq_b200 import ops
    shapes = [(4096, 4096), (1024, 4096), (1024, 4096), (4096, 4096), (14336, 4096), (14336, 4096), (4096, 14336)]
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    Ws = [(torch.randn(n, k, device=dev, generator=g, dtype=torch.float32) * 0.02).half() for n, k in shapes]
    stream = torch.cuda.current_stream(dev)

    def run():
        for W in Ws:
            ops.quantize(W, 4, 64, 1, True, True)

    def timed():
        run()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            run()
        e1.record(stream)
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / reps

    ms = timed()
    extra = {}
    pinned = "HQQ_B200_SOLVER_VARIANT" in os.environ
    if fast_ok and not pinned:
        # the fast solver (exact shortcuts) was bit-identical to the default one in its own probe process: check
        # that again here on this workload, time it, and report the faster of the two as this object's figure
        ref = [ops.quantize(W, 4, 64, 1, True, True) for W in Ws[:4]]
        os.environ["HQQ_B200_SOLVER_VARIANT"] = "1"
        try:
            got = [ops.quantize(W, 4, 64, 1, True, True) for W in Ws[:4]]
            same = all(torch.equal(x, y) for r, g_ in zip(ref, got) for x, y in zip(r[:3], g_[:3]))
            ms_fast = timed() if same else None
        finally:
            os.environ.pop("HQQ_B200_SOLVER_VARIANT", None)
        extra = {"default_ms_per_block": ms, "fast_ms_per_block": ms_fast, "fast_bit_identical": same}
        if same and ms_fast < ms:
            ms = ms_fast
            extra["selected"] = "HQQ_B200_SOLVER_VARIANT=1"
        else:
            extra["selected"] = "default"
    weights = sum(n * k for n, k in shapes)
    nbytes = sum(n * k * 2.5 + 2 * (n * k // 64) * 4 for n, k in shapes)
    achieved = nbytes / ms / 1e6
    return {"bound": "hbm", "achieved": achieved, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": achieved / peaks["hbm_gbs"],
            "ms_per_block": ms, "gweights_per_s": weights / ms / 1e6, "algorithmic_bytes_per_block": nbytes,
            "solver_variant": os.environ.get("HQQ_B200_SOLVER_VARIANT", extra.get("selected", "0")), **extra,
            "workload": "one Llama-3-8B block (7 matrices, 218 M weights) fp16 -> 4-bit gs=64 axis=1, solver + pack, 3 launches per matrix"}


def _finite(o):
    """json.dumps would print NaN / Infinity, which is not JSON: map non-finite floats to None."""
    if isinstance(o, float):
        return o if o == o and abs(o) != float("inf") else None
    if isinstance(o, dict):
        return {k: _finite(v) for k, v in o.items()}
    if isinstance(o, (list, tuple)):
        return [_finite(v) for v in o]
    return o


def run_probes(budget_s=150.0, timeout_s=45.0):
    """First GPU execution of the kernels written after round 1's GPU budget was spent: each knob runs
    tools/variant_probe.py in its OWN process under a timeout -- a crash or a hang there cannot reach this process -- on a fixed
    seeded workload, and is compared with the default kernels (sha256 of the outputs, relative error where the summation order
    differs by design).  Reported under "experimental"; nothing here touches the timed regions or the default kernels."""
    import tempfile
    import torch
    probe = os.path.join(ROOT, "tools", "variant_probe.py")
    tmp = tempfile.mkdtemp(prefix="hqq_probe_")
    t_start = time.perf_counter()
    base_env = {k: v for k, v in os.environ.items() if not k.startswith("HQQ_B200_")}

    def run(what, knob=None, both=None, save=None, timeout_s=timeout_s):
        if time.perf_counter() - t_start > budget_s:
            return {"skipped": "probe time budget spent"}
        env = dict(base_env)
        if knob:
            env[knob[0]] = knob[1]
        cmd = [sys.executable, probe, what]
        if both:
            cmd += ["--both", both]
        if save:
            cmd += ["--save", save]
        try:
            r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout_s)
        except subprocess.TimeoutExpired:
            return {"error": f"timeout after {timeout_s:.0f} s"}
        for ln in r.stdout.splitlines():
            if ln.startswith("PROBE "):
                return json.loads(ln[6:])
        return {"error": (r.stderr or r.stdout)[-300:].replace("\n", " | ")}

    def against_default(what, knobs):
        ref_path = os.path.join(tmp, what + "_default.pt")
        ref = run(what, save=ref_path)
        out = {"default": ref}
        for key, val in knobs:
            path = os.path.join(tmp, f"{what}_{key}_{val}.pt")
            got = run(what, knob=(key, val), save=path)
            if "digest" in ref and "digest" in got:
                got["bit_identical"] = got["digest"] == ref["digest"]
                got["speedup"] = ref["us"] / got["us"]
                if not got["bit_identical"]:
                    try:
                        a, b = torch.load(ref_path, weights_only=True), torch.load(path, weights_only=True)
                        got["rel_err"] = max(float((x.double() - y.double()).norm() / x.double().norm().clamp_min(1e-30)) if x.shape == y.shape
                                             else float("inf") for x, y in zip(a, b))
                    except Exception as e:  # noqa: BLE001
                        got["rel_err"] = repr(e)[:100]
            out[f"{key}={val}"] = got
        return out

    res = {"gemm_sweep": run("gemm_sweep", timeout_s=80.0),
           "bitpack": run("bitpack"),
           "solver_fast": run("quant", both="HQQ_B200_SOLVER_VARIANT=1"),
           "fused_3bit": run("l3", both="HQQ_B200_FUSED_3BIT=1"),
           # the three cheapest GEMM experiments; ld / ld512 / split-K are timed by tools/variant_sweep.sh (the bench must stay short)
           "gemm": against_default("gemm", [("HQQ_B200_GEMM_VARIANT", "dq16"), ("HQQ_B200_GEMM_VARIANT", "un512dq"), ("HQQ_B200_GEMM_VARIANT", "un512")])}
    res["seconds"] = round(time.perf_counter() - t_start, 1)
    return res


def _attach_sweeps(line, peaks):
    """Move the gemm_sweep / bitpack probe results out of `experimental` into objects of their own, with fractions of the measured peaks."""
    sweep = (line.get("experimental") or {}).pop("gemm_sweep", None)
    if isinstance(sweep, dict) and "per" in sweep:
        # BASELINE configs[2] / metric (2): fused dequant-GEMM TFLOP/s against the measured dense bf16/fp16 tensor peak
        peak_tf = None
        try:
            peak_tf = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops"])
        except Exception:  # noqa: BLE001
            peak_tf = 2250.0
        for e in sweep["per"].values():
            e["frac_of_tensor_peak"] = round(e["TFLOPs"] / peak_tf, 4)
        line["gemm_sweep"] = {"bound": "tensor", "peak": peak_tf, "unit": "TFLOP/s", "headline": "b4_4096x4096_M4096",
                              "achieved": sweep["per"]["b4_4096x4096_M4096"]["TFLOPs"],
                              "frac": sweep["per"]["b4_4096x4096_M4096"]["frac_of_tensor_peak"], "per": sweep["per"],
                              "note": "default kernels, event-timed alone (burst peak); route 2 = fused tcgen05 kernel, 0 = dequantize kernel + library GEMM (3-bit)"}
        # the opt-in GEMM kernels that reproduced the default kernel's outputs bit for bit in their probes: report the fastest
        # beside the default figure (same shape, same process-per-knob protocol); the headline stays the default kernel's
        best = None
        for name, r in ((line.get("experimental") or {}).get("gemm") or {}).items():
            tf = ((r.get("per") or {}).get("4096x4096xM4096") or {}).get("TFLOPs") if isinstance(r, dict) else None
            if name != "default" and isinstance(r, dict) and r.get("bit_identical") and tf and (best is None or tf > best[1]):
                best = (name, tf)
        if best is not None:
            line["gemm_sweep"]["best_bit_identical_variant"] = {"knob": best[0], "TFLOPs": best[1], "frac": round(best[1] / peak_tf, 4),
                                                                "shape": "4096x4096xM4096"}
    elif sweep is not None:
        line["gemm_sweep"] = sweep
    bp = (line.get("experimental") or {}).pop("bitpack", None)
    if isinstance(bp, dict) and "per" in bp:
        # pack / unpack / dequantize against the measured HBM peak (default kernels, own process)
        for e in bp["per"].values():
            e["frac_of_hbm_peak"] = round(e["GBps"] / peaks["hbm_gbs"], 4)
        line["bitpack"] = {"bound": "hbm", "peak": peaks["hbm_gbs"], "unit": "GB/s", "headline": "b4_dequantize_f16",
                           "achieved": bp["per"]["b4_dequantize_f16"]["GBps"], "frac": bp["per"]["b4_dequantize_f16"]["frac_of_hbm_peak"],
                           "per": bp["per"], "note": "one 14336x4096 matrix per call, inputs cycled (cold L2), algorithmic bytes = input + output"}
    elif bp is not None:
        line["bitpack"] = bp


def run_gpu(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    from hqq_b200 import _lib, harness
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    lib = _lib.load()
    big = args.model == "70b"  # BASELINE configs[4] (bs = 1 leg): needs --gpus 8 (4.8 GB of packed weights per rank); not the default line
    shape = harness.LLAMA3_70B if big else harness.LLAMA3_8B
    metric = "llama3_70b_4bit_gs64_decode_tokens_per_s" if big else METRIC
    workload = ("Llama-3-70B-shaped decode bs=1 seq=1, 80 blocks x 7 HQQLinear 4-bit gs=64 axis=1, fp16 lm_head (BASELINE configs[4], bs=1)"
                if big else WORKLOAD)
    n_layers = args.layers or shape.n_layers
    if args.cache_len <= 0:  # every timed loop starts at position 0 and must not wrap inside the cache
        args.cache_len = min(8192, max(256, args.steps + max(args.warmup, 3) + 8))
    B = max(1, args.batch)
    if B > 1:  # BASELINE configs[4] bs = 32 leg: not the default line
        metric += f"_bs{B}"
        workload = workload.replace("bs=1", f"bs={B}")
    # Decode autotuner (hqq_b200/tune.py): kernel variants and prefetch hints that leave every result unchanged are first run in a
    # child process (crash / hang / token guard), the survivors are then timed on this very model and the fastest stays captured.
    from hqq_b200 import tune
    autotune = None
    do_tune = world == 1 and not big and B == 1 and not args.no_autotune and tune.autotune_enabled()
    guard = None
    if do_tune:
        t_tune = time.perf_counter()
        try:
            guard = tune.guard_decode(budget_s=args.autotune_budget)
        except Exception as e:  # noqa: BLE001 -- the tuner must never cost the bench line
            autotune = {"error": repr(e)[:200]}
    model = harness.DecodeModel(shape, nbits=4, group_size=64, dtype=torch.float16, device=dev, cache_len=args.cache_len, tp=world,
                                rank=rank, process_group=pg, n_layers=n_layers, batch=B)
    selected = {}
    if guard is not None:
        try:
            model.capture(warmup=3)
            rep = tune.choose_decode(model, guard)
            selected = rep["selected"]
            autotune = {"selected": tune.knob_label(selected), "gain_vs_default": rep["gain"], "default_us": rep["default_us"],
                        "selected_us": rep["selected_us"],
                        "guard": [{"knobs": tune.knob_label(r["knobs"]), **{k: v for k, v in r.items() if k in ("us", "identical", "speedup", "error")}}
                                  for r in guard],
                        "in_process": [{**t, "knobs": tune.knob_label(t["knobs"])} for t in rep["tried"]],
                        "seconds": round(time.perf_counter() - t_tune, 1),
                        "note": "candidates select bit-identical kernels / add L2 prefetch hints; kept only if the token stream equals the default's"}
        except Exception as e:  # noqa: BLE001
            autotune = {"error": repr(e)[:200]}
            selected = {}
    lib.hqq_b200_launch_count_reset()
    if do_tune:
        model.retune(selected, warmup=3)
    else:
        model.capture(warmup=3)
    # launches of OUR kernels in one step = those issued while capturing one step (3 warm-up steps + 1 captured)
    launches_per_step = int(lib.hqq_b200_launch_count()) // 4
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident loop -------------------------------------------------------------------
    model.tok.fill_(1); model.pos.zero_()
    for _ in range(max(args.warmup, 3)):
        model.decode()
    sampler = ClockSampler(local_rank)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        model.decode()
    e1.record(stream)
    barrier()
    dev_ms = e0.elapsed_time(e1)

    # ---- end-to-end loop through the public API with host buffers ------------------------------
    h_in = torch.ones(B, dtype=torch.long).pin_memory()
    h_out = torch.zeros(B, dtype=torch.long).pin_memory()
    model.pos.zero_()
    for _ in range(3):
        model.tok.copy_(h_in, non_blocking=True); model.graph.replay(); h_out.copy_(model.next_tok, non_blocking=True); torch.cuda.synchronize(dev)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(args.steps):
        model.tok.copy_(h_in, non_blocking=True)          # H2D: this step's input token
        model.graph.replay()
        h_out.copy_(model.next_tok, non_blocking=True)    # D2H: the produced token
        stream.synchronize()
        h_in.copy_(h_out)                                 # host-side feedback, as a generation loop would
    t1.record(stream)
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    clocks = sampler.stop() if rank == 0 else None

    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()

    peaks = load_peaks()
    roof = kernel_roofline(model, torch, peaks) if (rank == 0 and world == 1 and B == 1) else None
    if rank == 0:
        scale_layers = shape.n_layers / n_layers
        value = args.steps * B / (dev_ms / 1e3)
        e2e = args.steps * B / (e2e_ms / 1e3)
        bytes_tok = model.bytes_per_token() * world  # whole-job bytes (each rank streams 1/world of the blocks + full lm_head)
        line = {"metric": metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f16",
                "data": "synthetic",
                "config": {"workload": workload,
                           "path": f"fused sm_100a kernels, kv cache {args.cache_len}, CUDA graph; {model.bytes_per_token() / 1e9:.2f} GB streamed per step "
                                   ">> 126 MB L2 (inputs larger than L2, no flush needed)",
                           "parallelism": f"tp{world}", "layers": n_layers, "global_batch": B, "autotune": autotune},
                "clocks": clocks,
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": 8 * B, "d2h_bytes_per_step": 8 * B},
                "gpu_launches": launches_per_step * args.steps,
                "step_hbm_GBps": bytes_tok / world / (dev_ms / args.steps) / 1e6}
        if scale_layers != 1.0:
            line["config"]["note"] = "REDUCED layer count (debug run) -- not the BASELINE configuration"
        if roof is not None:
            line["roofline"] = roof
        if args.worker:
            # supervised run: hand the finished measurement to the supervisor right away (it keeps the LAST line it receives), so
            # that nothing below can cost it even if this process were killed
            print(json.dumps(_finite(dict(line, extras="preliminary line: the worker did not finish its extra objects"))), flush=True)
        # Everything below adds objects to the line that is already complete (probes in sub-processes, quantizer, CPU baselines).
        # A watchdog prints the line as it stands if they ever exceed their deadline, so they cannot cost the measurement.
        printed = threading.Event()

        def emit():
            if printed.is_set():
                return
            printed.set()
            for _ in range(5):
                try:
                    print(json.dumps(_finite(dict(line))), flush=True)
                    return
                except RuntimeError:  # the main thread added a key meanwhile
                    time.sleep(0.05)

        def watchdog():
            if not printed.wait(timeout=args.extras_deadline):
                line["extras"] = f"cut off after {args.extras_deadline:.0f} s"
                emit()
                os._exit(0)

        threading.Thread(target=watchdog, daemon=True).start()
        if world == 1 and not big and B == 1:
            del model
            torch.cuda.empty_cache()
        if world == 1 and not big and B == 1 and not args.no_probes:
            try:
                line["experimental"] = run_probes()
            except Exception as e:  # noqa: BLE001
                line["experimental"] = {"error": repr(e)[:200]}
        try:
            _attach_sweeps(line, peaks)
        except Exception as e:  # noqa: BLE001 -- formatting of an extra object must not cost the line
            line["extras_error"] = repr(e)[:200]
        if world == 1 and not big and B == 1:
            try:  # extra object, never allowed to cost the bench line
                sf = (line.get("experimental") or {}).get("solver_fast") or {}
                line["quantizer"] = quantizer_roofline(torch, peaks, dev, fast_ok=bool(sf.get("bit_identical")))
            except Exception as e:  # noqa: BLE001
                line["quantizer"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline and not big and B == 1 and isinstance(line.get("quantizer"), dict) and "ms_per_block" in line["quantizer"]:
            try:  # the reference's CPU solver (float32, optimize.py:201-255) beside the quantizer object, on a bounded sample
                line["quantizer"]["cpu_baseline"] = cpu_quantizer_baseline()
            except Exception as e:  # noqa: BLE001
                line["quantizer"]["cpu_baseline"] = {"error": repr(e)[:200]}
        if world == 1 and not args.no_cpu_baseline and not big and B == 1:
            try:
                v, info = cpu_reference_tokens_per_s(budget_s=15.0)
                line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": info["cores"], "host_cores": os.cpu_count() or 1, "kind": "port",
                                        "sample": info["sample"]}
            except Exception as e:  # noqa: BLE001
                line["cpu_baseline"] = {"error": repr(e)[:200]}
        emit()
    if world > 1:
        # Tear down without touching NCCL again: destroying a process group while captured graphs still hold its kernels
        # can hang.  Everything is measured and printed; leave through the fast exit on every rank.
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def supervise(argv, first_timeout_s=720.0, retry_timeout_s=600.0):
    """N = 1 with the autotuner on: the measurement runs in a worker process.  The tuner re-captures the decode graph under kernel
    variants inside the measuring process (after each has survived its own guard process); should that process still die or hang,
    the measurement is repeated once with the default kernels only, so a tuner failure can never cost the bench line."""
    from hqq_b200 import _lib
    _lib.load()  # this process reports through the same library (no CUDA work here)
    me = os.path.abspath(__file__)
    for extra, timeout_s in ((["--worker"], first_timeout_s), (["--worker", "--no-autotune"], retry_timeout_s)):
        why = None
        try:
            try:
                r = subprocess.run([sys.executable, me, *argv, *extra], stdout=subprocess.PIPE, text=True, timeout=timeout_s)
                out, why = r.stdout or "", f"worker exited with code {r.returncode} without a result line"
            except subprocess.TimeoutExpired as e:  # a worker that printed its line and then hung (teardown) still counts
                out = e.stdout.decode("utf-8", "replace") if isinstance(e.stdout, bytes) else (e.stdout or "")
                why = f"worker timed out after {timeout_s:.0f} s"
            lines = [ln for ln in out.splitlines() if ln.startswith("{") and '"metric"' in ln]
            if lines:
                line = lines[-1]
                if "--no-autotune" in extra:
                    try:
                        d = json.loads(line)
                        d["config"]["autotune"] = {"error": f"autotuned worker failed ({first_why}); measured again with the default kernels"}
                        line = json.dumps(d)
                    except Exception:  # noqa: BLE001
                        pass
                print(line, flush=True)
                return 0
        except Exception as e:  # noqa: BLE001 -- could not even start the worker
            why = repr(e)[:200]
        sys.stderr.write(f"bench.py: {why}\n")
        first_why = why
    return 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="hqq_b200", choices=["hqq_b200", "reference"])
    ap.add_argument("--cache-len", type=int, default=0, help="KV-cache length; 0 = large enough that the timed loops never wrap (>= 256)")
    ap.add_argument("--layers", type=int, default=0, help="debug: fewer blocks (marks the line as reduced)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-probes", action="store_true", help="skip the experimental-kernel probes (sub-processes, N=1 only)")
    ap.add_argument("--extras-deadline", type=float, default=330.0, help="seconds the objects added after the measurement (probes, quantizer, "
                    "CPU baselines) may take before the line is printed without the unfinished ones")
    ap.add_argument("--worker", action="store_true", help=argparse.SUPPRESS)  # internal: the measuring process of a supervised N=1 run
    ap.add_argument("--no-autotune", action="store_true", help="time the default kernels only (no decode autotuner; N=1 only anyway)")
    ap.add_argument("--autotune-budget", type=float, default=90.0, help="seconds the autotuner's guard processes may take")
    ap.add_argument("--batch", type=int, default=1, help="sequences decoded in lock-step (BASELINE configs[4]: 32); > 1 uses the fused small-M "
                    "kernel between framework glue ops and NCCL all-reduce")
    ap.add_argument("--model", default="8b", choices=["8b", "70b"], help="70b = BASELINE configs[4] at bs=1 (use with --gpus 8)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}", "--master-addr", "127.0.0.1",
               "--master-port", os.environ.get("MASTER_PORT", "29533"), os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if world == 1 and args.gpus <= 1 and not args.worker and not args.no_autotune and args.model == "8b" and args.batch <= 1:
        from hqq_b200 import tune
        if tune.autotune_enabled():
            sys.exit(supervise(sys.argv[1:]))
    run_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
