"""Host logic of the decode autotuner (hqq_b200/tune.py) without a GPU: the guard's bookkeeping over child processes that finish,
crash, hang or run out of time is driven through its `run_child` seam, and the real child runner is exercised against a stand-in
child script."""
import json
import os
import sys
import time

import pytest

from hqq_b200 import tune
from hqq_b200.harness import DecodeModel

CANDS = [{"HQQ_B200_D1_VARIANT": "1042"}, {"HQQ_B200_WPF_MB": "8"}, {"HQQ_B200_D1_VARIANT": "7042"}]


def fake_child(script):
    """script: knobs-label -> ("ok", us, digest) | ("crash",) | ("hang",); the default ("default") must be listed."""
    calls = []

    def run_child(cands, layers, first_line_s, per_line_s, deadline):
        calls.append([tune.knob_label(c) for c in cands])
        done = {}
        for j, c in enumerate(cands):
            what = script[tune.knob_label(c)]
            if what[0] == "ok":
                done[j] = {"i": j, "knobs": c, "us": what[1], "digest": what[2]}
            elif what[0] == "crash":
                return done, j, "child exited (code -6)"
            else:
                return done, j, "no progress for 20 s"
        return done, None, None

    run_child.calls = calls
    return run_child


def test_guard_all_candidates_finish():
    rc = fake_child({"default": ("ok", 100.0, "aa"), "D1_VARIANT=1042": ("ok", 80.0, "aa"), "WPF_MB=8": ("ok", 125.0, "aa"),
                     "D1_VARIANT=7042": ("ok", 50.0, "bb")})
    out = tune.guard_decode(CANDS, run_child=rc)
    assert len(rc.calls) == 1 and rc.calls[0][0] == "default"
    assert out[0]["knobs"] == {} and out[0]["us"] == 100.0
    assert out[1]["identical"] and out[1]["speedup"] == pytest.approx(1.25)
    assert out[2]["identical"] and out[2]["speedup"] == pytest.approx(0.8)
    assert out[3]["identical"] is False  # fast but different tokens: never a survivor


def test_guard_drops_a_crashing_candidate_and_continues_in_a_new_child():
    rc = fake_child({"default": ("ok", 100.0, "aa"), "D1_VARIANT=1042": ("crash",), "WPF_MB=8": ("ok", 90.0, "aa"),
                     "D1_VARIANT=7042": ("hang",)})
    out = tune.guard_decode(CANDS, run_child=rc)
    assert rc.calls == [["default", "D1_VARIANT=1042", "WPF_MB=8", "D1_VARIANT=7042"], ["default", "WPF_MB=8", "D1_VARIANT=7042"]]
    assert "exited" in out[1]["error"]
    assert out[2]["identical"] and out[2]["speedup"] == pytest.approx(100.0 / 90.0)
    assert "no progress" in out[3]["error"]


def test_guard_stops_when_the_default_kernels_fail_in_the_child():
    rc = fake_child({"default": ("crash",), "D1_VARIANT=1042": ("ok", 1.0, "aa"), "WPF_MB=8": ("ok", 1.0, "aa"), "D1_VARIANT=7042": ("ok", 1.0, "aa")})
    out = tune.guard_decode(CANDS, run_child=rc)
    assert len(rc.calls) == 1
    assert all("error" in r for r in out)


def test_guard_respects_its_time_budget():
    def slow(cands, layers, first_line_s, per_line_s, deadline):
        time.sleep(0.05)
        return {0: {"i": 0, "knobs": {}, "us": 1.0, "digest": "aa"}}, 1, "time budget spent"

    out = tune.guard_decode(CANDS, budget_s=0.01, run_child=slow)
    assert out[0]["us"] == 1.0
    assert all(r["error"] == "time budget spent" for r in out[1:])


CHILD = r'''
import json, sys, time, os
cands = json.loads(sys.argv[sys.argv.index("--child") + 1])
for i, c in enumerate(cands):
    print("TRY " + json.dumps({"i": i}), flush=True)
    v = c.get("HQQ_B200_D1_VARIANT")
    if v == "1042":
        os._exit(134)
    if v == "7042":
        time.sleep(60)
    print("CAND " + json.dumps({"i": i, "knobs": c, "us": 10.0 + i, "digest": "aa"}), flush=True)
print("DONE", flush=True)
'''


def test_real_child_runner_detects_exit_and_hang(tmp_path, monkeypatch):
    """_run_child against a stand-in for `python -m hqq_b200.tune --child`: a fake package of the same name placed in a directory
    the runner is pointed at."""
    pkg = tmp_path / "hqq_b200"
    pkg.mkdir()
    (pkg / "__init__.py").write_text("")
    (pkg / "tune.py").write_text(CHILD)
    monkeypatch.setattr(tune, "__file__", str(pkg / "tune.py"))
    t0 = time.perf_counter()
    done, running, why = tune._run_child([{}, {"HQQ_B200_WPF_MB": "8"}], 8, 30.0, 5.0, time.perf_counter() + 60)
    assert why is None and running is None and sorted(done) == [0, 1]
    done, running, why = tune._run_child([{}, {"HQQ_B200_D1_VARIANT": "1042"}, {"HQQ_B200_WPF_MB": "8"}], 8, 30.0, 5.0, time.perf_counter() + 60)
    assert sorted(done) == [0] and running == 1 and "exited" in why
    done, running, why = tune._run_child([{}, {"HQQ_B200_D1_VARIANT": "7042"}], 8, 30.0, 1.0, time.perf_counter() + 60)
    assert sorted(done) == [0] and running == 1 and "no progress" in why
    assert time.perf_counter() - t0 < 30


def test_candidates_only_use_tunable_knobs():
    for c in tune.DECODE_CANDIDATES:
        assert set(c) <= set(DecodeModel.TUNABLE)
    assert tune.knob_label({}) == "default"


def test_autotune_is_off_when_a_knob_is_pinned_by_hand(monkeypatch):
    for k in DecodeModel.TUNABLE + ("HQQ_B200_AUTOTUNE",):
        monkeypatch.delenv(k, raising=False)
    assert tune.autotune_enabled()
    monkeypatch.setenv("HQQ_B200_WPF_MB", "16")
    assert not tune.autotune_enabled()
    monkeypatch.delenv("HQQ_B200_WPF_MB")
    monkeypatch.setenv("HQQ_B200_AUTOTUNE", "0")
    assert not tune.autotune_enabled()


class FakeModel:
    def __init__(self):
        self.knobs, self.history = None, []

    def retune(self, knobs=None, warmup=2):
        self.knobs = dict(knobs or {})
        self.history.append(tune.knob_label(self.knobs))


def fake_measure(table):
    import torch

    def measure(model, steps=30, rounds=2, start_pos=20):
        us, tok = table[tune.knob_label(model.knobs)]
        return torch.full((tune.N_CHECK_TOKENS, 1), tok, dtype=torch.long), us

    return measure


GUARD = [{"knobs": {}, "us": 100.0, "digest": "aa"},
         {"knobs": {"HQQ_B200_D1_VARIANT": "1042"}, "us": 90.0, "digest": "aa", "identical": True, "speedup": 1.11},
         {"knobs": {"HQQ_B200_WPF_MB": "8"}, "us": 70.0, "digest": "aa", "identical": True, "speedup": 1.43},
         {"knobs": {"HQQ_B200_D1_VARIANT": "7042"}, "us": 50.0, "digest": "bb", "identical": False, "speedup": 2.0},
         {"knobs": {"HQQ_B200_D1_VARIANT": "2042"}, "us": 120.0, "digest": "aa", "identical": True, "speedup": 0.83},
         {"knobs": {"HQQ_B200_D1_VARIANT": "4042"}, "error": "child exited (code -11)"}]


def test_choose_keeps_the_fastest_identical_candidate():
    m = FakeModel()
    rep = tune.choose_decode(m, GUARD, measure_fn=fake_measure({"default": (400.0, 5), "WPF_MB=8": (300.0, 5), "D1_VARIANT=1042": (350.0, 5)}))
    assert rep["selected"] == {"HQQ_B200_WPF_MB": "8"} and m.knobs == rep["selected"]
    assert rep["gain"] == pytest.approx(400.0 / 300.0)
    # only guard survivors that were identical AND faster are ever captured in this process, best first
    assert m.history == ["default", "WPF_MB=8", "D1_VARIANT=1042", "WPF_MB=8"]


def test_choose_rejects_a_candidate_whose_tokens_differ_on_the_real_model():
    m = FakeModel()
    rep = tune.choose_decode(m, GUARD, measure_fn=fake_measure({"default": (400.0, 5), "WPF_MB=8": (100.0, 6), "D1_VARIANT=1042": (399.0, 5)}))
    assert rep["selected"] == {} and m.knobs == {}  # WPF_MB=8: wrong tokens; 1042: inside the noise margin
    assert rep["gain"] == 1.0
    assert [t["identical"] for t in rep["tried"]] == [False, True]


def test_choose_with_nothing_to_try_leaves_the_default_captured():
    m = FakeModel()
    rep = tune.choose_decode(m, [GUARD[0], GUARD[3], GUARD[5]], measure_fn=fake_measure({"default": (400.0, 5)}))
    assert rep["selected"] == {} and rep["tried"] == [] and m.history == ["default", "default"]


def _fake_run(script):
    """subprocess.run stand-in for bench.supervise: `script` is a list of outcomes, one per worker launch."""
    import subprocess
    calls = []

    def run(cmd, stdout=None, text=None, timeout=None):
        calls.append(cmd)
        what = script[len(calls) - 1]
        if what == "timeout":
            raise subprocess.TimeoutExpired(cmd, timeout)
        if isinstance(what, tuple) and what[0] == "timeout_after_line":
            raise subprocess.TimeoutExpired(cmd, timeout, output=what[1].encode())
        code, out = what
        return subprocess.CompletedProcess(cmd, code, stdout=out)

    run.calls = calls
    return run


LINE = json.dumps({"metric": "m", "value": 1.0, "config": {"autotune": {"selected": "WPF_MB=8"}}})


def test_bench_supervisor_passes_the_worker_line_through(monkeypatch, capsys):
    import bench
    run = _fake_run([(0, "noise\n" + LINE + "\n")])
    monkeypatch.setattr(bench.subprocess, "run", run)
    assert bench.supervise(["--steps", "5"]) == 0
    assert capsys.readouterr().out.strip() == LINE
    assert run.calls[0][-1] == "--worker" and "--steps" in run.calls[0]


@pytest.mark.parametrize("first", [(-11, ""), (1, "Traceback ...\n"), "timeout"])
def test_bench_supervisor_measures_again_with_default_kernels_when_the_tuned_worker_fails(monkeypatch, capsys, first):
    import bench
    run = _fake_run([first, (0, LINE + "\n")])
    monkeypatch.setattr(bench.subprocess, "run", run)
    assert bench.supervise([]) == 0
    d = json.loads(capsys.readouterr().out.strip())
    assert d["value"] == 1.0 and "measured again with the default kernels" in d["config"]["autotune"]["error"]
    assert run.calls[1][-2:] == ["--worker", "--no-autotune"]


def test_bench_supervisor_reports_failure_when_both_workers_fail(monkeypatch, capsys):
    import bench
    monkeypatch.setattr(bench.subprocess, "run", _fake_run([(1, ""), (1, "")]))
    assert bench.supervise([]) == 1
    assert capsys.readouterr().out.strip() == ""


def test_retune_sets_the_environment_reloads_the_library_and_recaptures(monkeypatch):
    """DecodeModel.retune without a GPU: capture() and the library are stubbed; the knobs of the previous choice must not leak."""
    import os
    from hqq_b200 import _lib, harness

    class Lib:
        reloads = 0

        def hqq_b200_reload_env(self):
            Lib.reloads += 1

    monkeypatch.setattr(_lib, "load", lambda path=None: Lib())
    for k in DecodeModel.TUNABLE:
        monkeypatch.delenv(k, raising=False)
    m = harness.DecodeModel.__new__(harness.DecodeModel)
    captured = []
    m.capture = lambda warmup=3: captured.append((warmup, {k: os.environ.get(k) for k in DecodeModel.TUNABLE})) or "graph"
    assert m.retune({"HQQ_B200_D1_VARIANT": "7042", "HQQ_B200_WPF_MB": "48", "HQQ_B200_WPF_FROM": "o,gu"}) == "graph"
    assert (m.wpf_mb, m.wpf_ahead, m.wpf_from) == (48.0, 2, frozenset({"o", "gu"}))
    assert captured[-1][1]["HQQ_B200_D1_VARIANT"] == "7042" and Lib.reloads == 1
    m.retune({"HQQ_B200_WPF_MB": "8", "HQQ_B200_WPF_BULK": "16"}, warmup=3)
    assert captured[-1] == (3, {"HQQ_B200_D1_VARIANT": None, "HQQ_B200_WPF_MB": "8", "HQQ_B200_WPF_AHEAD": None, "HQQ_B200_WPF_FROM": None,
                                "HQQ_B200_WPF_BULK": "16"})
    assert (m.wpf_mb, m.wpf_ahead, m.wpf_from) == (8.0, 1, frozenset(harness.WPF_STAGES))
    m.retune({})
    assert all(v is None for v in captured[-1][1].values()) and m.wpf_mb == 0.0 and Lib.reloads == 3
    with pytest.raises(ValueError):
        m.retune({"HQQ_B200_PDL": "0"})
    for k in DecodeModel.TUNABLE:
        os.environ.pop(k, None)


def test_measure_runs_the_token_check_and_the_timed_loops(monkeypatch):
    """tune.measure with the CUDA calls stubbed: 16 tokens from the reset state, then `rounds` timed loops that restart at `start_pos`."""
    import torch

    class Ev:
        def __init__(self, enable_timing=True):
            pass

        def record(self, stream=None):
            pass

        def elapsed_time(self, other):
            return 3.0  # ms

    monkeypatch.setattr(torch.cuda, "synchronize", lambda d=None: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: object())
    monkeypatch.setattr(torch.cuda, "Event", Ev)

    class M:
        device = torch.device("cpu")

        def __init__(self):
            self.pos, self.next_tok, self.steps, self.resets, self.starts = torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long), 0, 0, []

        def reset_state(self):
            self.resets += 1
            self.pos.zero_()

        def decode(self):
            if int(self.pos) == 20:
                self.starts.append(self.steps)
            self.steps += 1
            self.next_tok.fill_(self.steps)
            self.pos.add_(1)

    m = M()
    toks, us = tune.measure(m, steps=30, rounds=2, start_pos=20)
    assert m.resets == 1 and m.steps == tune.N_CHECK_TOKENS + 60 and m.starts == [16, 46]
    assert toks.shape == (tune.N_CHECK_TOKENS, 1) and toks[:, 0].tolist() == list(range(1, 17))
    assert us == pytest.approx(3.0 * 1e3 / 30)


def test_bench_supervisor_accepts_the_line_of_a_worker_that_hangs_afterwards(monkeypatch, capsys):
    import bench
    run = _fake_run([("timeout_after_line", LINE + "\n")])
    monkeypatch.setattr(bench.subprocess, "run", run)
    assert bench.supervise([]) == 0 and len(run.calls) == 1
    assert capsys.readouterr().out.strip() == LINE


def test_bench_supervisor_prefers_the_final_line_and_falls_back_to_the_preliminary_one(monkeypatch, capsys):
    import bench
    pre = json.dumps({"metric": "m", "value": 1.0, "extras": "preliminary line", "config": {"autotune": None}})
    monkeypatch.setattr(bench.subprocess, "run", _fake_run([(0, pre + "\n" + LINE + "\n")]))
    assert bench.supervise([]) == 0 and capsys.readouterr().out.strip() == LINE
    monkeypatch.setattr(bench.subprocess, "run", _fake_run([("timeout_after_line", pre + "\n")]))
    assert bench.supervise([]) == 0 and json.loads(capsys.readouterr().out.strip())["extras"] == "preliminary line"
