"""Decode-path autotuner: pick, on the GPU the process runs on, among kernel variants that produce IDENTICAL results.

The one-token path has tuning knobs that never change a result (`DecodeModel.TUNABLE`): `HQQ_B200_D1_VARIANT` selects among
bit-identical instantiations of `linear_decode1_kernel` (meta through the cp.async ring, evict-first weight stream, L2 prefetch
under the dependency wait) and `HQQ_B200_WPF_MB/_AHEAD` adds pure L2 prefetch hints for the following launches' weights.  Which of
them pays depends on the box (HBM clocks, L2 behaviour), so the choice is measured, in two steps:

1. **guard** (`guard_decode`): every candidate first runs in a *child process* (`python -m hqq_b200.tune --child ...`) on an
   8-block model of the same shapes: a crash, a hang or a wrong token there cannot reach the caller.  A candidate survives only if
   the 16 tokens it decodes from a fixed state are identical to the default kernels' tokens.
2. **choose** (`choose_decode`): the survivors that were faster in the child are re-captured on the caller's own model
   (`DecodeModel.retune`), must reproduce the default token stream there as well, and are timed with CUDA events; the fastest
   configuration that beats the default by `min_gain` stays captured -- otherwise the default kernels do.

No CPU arithmetic, no second backend: every candidate is this package's sm_100a code.  bench.py reports the outcome in
`config.autotune`; `HQQ_B200_AUTOTUNE=0` (or any TUNABLE knob set by hand) switches the tuner off.
"""
from __future__ import annotations

import hashlib
import json
import os
import queue
import subprocess
import sys
import threading
import time

# Candidates beyond the default kernels ({}): D1 variants alone, weight prefetch alone, and prefetch on top of the full variant.
DECODE_CANDIDATES = [
    {"HQQ_B200_D1_VARIANT": "1042"},
    {"HQQ_B200_D1_VARIANT": "2042"},
    {"HQQ_B200_D1_VARIANT": "4042"},
    {"HQQ_B200_D1_VARIANT": "7042"},
    {"HQQ_B200_D1_VARIANT": "3042"},
    {"HQQ_B200_D1_VARIANT": "7033"},
    # weight prefetch from every linear launch (each covers the next launch / the next two)
    {"HQQ_B200_WPF_MB": "8"},
    {"HQQ_B200_WPF_MB": "24"},
    {"HQQ_B200_WPF_MB": "48", "HQQ_B200_WPF_AHEAD": "1"},
    {"HQQ_B200_WPF_MB": "48", "HQQ_B200_WPF_AHEAD": "2"},
    # only from the launches whose prologue runs while HBM idles: o (under attention) pulls gate/up, gate/up (under o) pulls down
    {"HQQ_B200_WPF_MB": "48", "HQQ_B200_WPF_AHEAD": "1", "HQQ_B200_WPF_FROM": "o"},
    {"HQQ_B200_WPF_MB": "64", "HQQ_B200_WPF_AHEAD": "1", "HQQ_B200_WPF_FROM": "o"},
    {"HQQ_B200_WPF_MB": "64", "HQQ_B200_WPF_AHEAD": "1", "HQQ_B200_WPF_FROM": "o,gu"},
    {"HQQ_B200_WPF_MB": "88", "HQQ_B200_WPF_AHEAD": "2", "HQQ_B200_WPF_FROM": "o"},
    # the same volumes through the bulk-copy unit (one cp.async.bulk.prefetch.L2 per 16 / 64 KiB instead of one prefetch per line)
    {"HQQ_B200_WPF_MB": "24", "HQQ_B200_WPF_BULK": "16"},
    {"HQQ_B200_WPF_MB": "64", "HQQ_B200_WPF_AHEAD": "1", "HQQ_B200_WPF_FROM": "o,gu", "HQQ_B200_WPF_BULK": "16"},
    {"HQQ_B200_WPF_MB": "64", "HQQ_B200_WPF_AHEAD": "1", "HQQ_B200_WPF_FROM": "o,gu", "HQQ_B200_WPF_BULK": "64"},
    # prefetch on top of the kernel variants
    {"HQQ_B200_D1_VARIANT": "1042", "HQQ_B200_WPF_MB": "24"},
    {"HQQ_B200_D1_VARIANT": "7042", "HQQ_B200_WPF_MB": "24"},
    {"HQQ_B200_D1_VARIANT": "7042", "HQQ_B200_WPF_MB": "64", "HQQ_B200_WPF_AHEAD": "1", "HQQ_B200_WPF_FROM": "o,gu"},
]

N_CHECK_TOKENS = 16


def knob_label(knobs: dict) -> str:
    return ",".join(f"{k.replace('HQQ_B200_', '')}={v}" for k, v in sorted(knobs.items())) or "default"


def token_digest(tokens) -> str:
    """sha256 over the raw bytes of a [steps, batch] int64 token tensor (equal digests = the same token stream)."""
    return hashlib.sha256(tokens.detach().contiguous().cpu().numpy().tobytes()).hexdigest()[:16]


def measure(model, steps: int = 30, rounds: int = 2, start_pos: int = 20):
    """(tokens [N_CHECK_TOKENS, batch] decoded from the reset state, best µs per step over `rounds` timed loops at `start_pos`)."""
    import torch
    model.reset_state()
    toks = []
    for _ in range(N_CHECK_TOKENS):
        model.decode()
        toks.append(model.next_tok.clone())
    torch.cuda.synchronize(model.device)
    stream = torch.cuda.current_stream(model.device)
    best = float("inf")
    for _ in range(rounds):
        model.pos.fill_(start_pos)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            model.decode()
        e1.record(stream)
        torch.cuda.synchronize(model.device)
        best = min(best, e0.elapsed_time(e1) * 1e3 / steps)
    return torch.stack(toks), best


# ---------------------------------------------------------------------------------------------------------- guard (child process)
def _child(candidates, layers: int, shape: dict | None = None, nbits: int = 4, group_size: int = 64):
    import torch
    from . import harness
    dev = torch.device("cuda", 0)
    shp = harness.LlamaShape(**shape) if shape else harness.LLAMA3_8B
    m = harness.DecodeModel(shp, nbits=nbits, group_size=group_size, dtype=torch.float16, device=dev, cache_len=64, n_layers=layers)
    for i, knobs in enumerate(candidates):
        print("TRY " + json.dumps({"i": i}), flush=True)
        m.retune(knobs, warmup=2)
        toks, us = measure(m, steps=30, rounds=2, start_pos=20)
        print("CAND " + json.dumps({"i": i, "knobs": knobs, "us": us, "digest": token_digest(toks)}), flush=True)
    print("DONE", flush=True)


def _run_child(candidates, layers, first_line_s, per_line_s, deadline, model_args=None):
    """One child over `candidates`.  Returns ({position: result}, running, why): `why` is None when the child finished the list;
    otherwise `running` is the position it had announced (TRY) and not completed, or None if it died outside a candidate."""
    env = {k: v for k, v in os.environ.items() if not k.startswith("HQQ_B200_")}
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env["PYTHONPATH"] = root + os.pathsep + env.get("PYTHONPATH", "")
    cmd = [sys.executable, "-m", "hqq_b200.tune", "--child", json.dumps(candidates), "--layers", str(layers)]
    if model_args:
        cmd += ["--model", json.dumps(model_args)]
    proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, env=env, cwd=root)
    q: queue.Queue = queue.Queue()

    def reader():
        for ln in proc.stdout:
            q.put(ln)
        q.put(None)

    threading.Thread(target=reader, daemon=True).start()
    done, running, why, finished = {}, None, None, False
    wait = first_line_s
    while True:
        remaining = deadline - time.perf_counter()
        if remaining <= 0:
            why = "time budget spent"
            break
        try:
            ln = q.get(timeout=min(wait, remaining))
        except queue.Empty:
            why = "time budget spent" if remaining < wait else f"no progress for {wait:.0f} s"
            break
        if ln is None:  # end of the child's output
            why = None if finished else f"child exited (code {proc.wait()})"
            break
        if ln.startswith("TRY "):
            running = json.loads(ln[4:])["i"]
            wait = per_line_s
        elif ln.startswith("CAND "):
            r = json.loads(ln[5:])
            done[r["i"]] = r
            running = None
            wait = per_line_s
        elif ln.startswith("DONE"):
            finished = True
    if proc.poll() is None:
        proc.kill()
    try:
        proc.wait(timeout=10)
    except subprocess.TimeoutExpired:
        pass
    return done, running, why


def guard_decode(candidates=None, layers: int = 8, budget_s: float = 90.0, first_line_s: float = 75.0, per_line_s: float = 20.0,
                 run_child=None, model_args=None):
    """Run the default kernels and every candidate in child processes.  Returns one entry per configuration, entry 0 being the
    default kernels: {"knobs", "us", "digest", "identical", "speedup"} for a candidate that ran, {"knobs", "error"} for one that
    crashed, hung or was not reached.  `model_args` = {"shape": LlamaShape fields, "nbits", "group_size"} makes the child build another
    shape than Llama-3-8B.  A candidate that kills its child is dropped and the rest continue in a new child (which
    starts with the default kernels again, so speed-ups are always relative to the same process).  `run_child` is the seam the
    CPU tests use."""
    if run_child is None:
        run_child = (lambda *a: _run_child(*a, model_args=model_args)) if model_args else _run_child
    cands = [{}] + [dict(c) for c in (DECODE_CANDIDATES if candidates is None else candidates)]
    out = [None] * len(cands)
    deadline = time.perf_counter() + budget_s
    pending = list(range(len(cands)))

    def fail(ids, why):
        for i in ids:
            if out[i] is None:
                out[i] = {"knobs": cands[i], "error": why}

    while pending:
        if time.perf_counter() >= deadline:
            fail(pending, "time budget spent")
            break
        batch = pending if pending[0] == 0 else [0] + pending  # the default configuration leads every child
        done, running, why = run_child([cands[i] for i in batch], layers, first_line_s, per_line_s, deadline)
        ref = done.get(0)
        for j, r in sorted(done.items()):
            i = batch[j]
            if out[i] is None:
                out[i] = {"knobs": cands[i], "us": r["us"], "digest": r["digest"]}
                if i != 0 and ref is not None:
                    out[i]["identical"] = r["digest"] == ref["digest"]
                    out[i]["speedup"] = ref["us"] / r["us"]
        pending = [i for i in pending if out[i] is None]
        if why is None:
            fail(pending, "not reported by the child")  # cannot happen with a well-formed child
            break
        if running is None or running >= len(batch) or batch[running] == 0 or ref is None:
            fail(pending, why if running is None or ref is not None else f"default configuration failed in the guard: {why}")
            break
        bad = batch[running]
        fail([bad], why)
        pending = [i for i in pending if i != bad]
    return out


# ---------------------------------------------------------------------------------------------------------- choose (in process)
def choose_decode(model, guard, top: int = 4, steps: int = 40, min_gain: float = 1.02, measure_fn=None):
    """Re-capture `model` under the guard's best survivors, keep the fastest one that reproduces the default token stream and beats
    the default by `min_gain`; otherwise the default kernels stay.  Returns a report dict; `model` is left captured under
    report["selected"]."""
    import torch
    measure = measure_fn or globals()["measure"]  # the seam the CPU tests use
    report = {"selected": {}, "tried": []}
    ok = [r for r in guard[1:] if r.get("identical") and r.get("speedup", 0.0) > 1.0]
    ok.sort(key=lambda r: -r["speedup"])
    model.retune({})
    measure(model, steps=steps)  # clocks and caches settle on a throw-away pass: the default must not look slow for being first
    ref_toks, ref_us = measure(model, steps=steps)
    report["default_us"] = ref_us
    best_knobs, best_us = {}, ref_us
    for r in ok[:top]:
        knobs = r["knobs"]
        try:
            model.retune(knobs)
            toks, us = measure(model, steps=steps)
        except Exception as e:  # noqa: BLE001 -- a launch the guard accepted on 8 blocks was refused here: keep the default
            report["tried"].append({"knobs": knobs, "error": repr(e)[:160]})
            continue
        same = bool(torch.equal(toks, ref_toks))
        report["tried"].append({"knobs": knobs, "us": us, "identical": same, "guard_speedup": r["speedup"]})
        if same and us * min_gain < ref_us and us < best_us:
            best_knobs, best_us = knobs, us
    model.retune(best_knobs)
    if best_knobs:  # the winner must still hold once it is the captured graph
        toks, us = measure(model, steps=steps)
        if not torch.equal(toks, ref_toks) or us * min_gain >= ref_us:
            best_knobs, us = {}, ref_us
            model.retune({})
        best_us = us
    report["selected"] = best_knobs
    report["selected_us"] = best_us
    report["gain"] = ref_us / best_us
    return report


def autotune_enabled() -> bool:
    if os.environ.get("HQQ_B200_AUTOTUNE", "1") == "0":
        return False
    from .harness import DecodeModel
    return not any(k in os.environ for k in DecodeModel.TUNABLE)


if __name__ == "__main__":
    if "--child" in sys.argv:
        margs = json.loads(sys.argv[sys.argv.index("--model") + 1]) if "--model" in sys.argv else {}
        _child(json.loads(sys.argv[sys.argv.index("--child") + 1]), int(sys.argv[sys.argv.index("--layers") + 1]), **margs)
