"""CPU: bench.run_gpu's control flow with every GPU-touching piece replaced by a stand-in -- the autotune report, the bench line's
objects (roofline, gemm_sweep, bitpack, quantizer, cpu_baseline), the launch count and the watchdog / emit path are assembled by
the real code.  (The numbers are fake; this guards the plumbing that the round-end bench runs for the first time on a GPU.)"""
import argparse
import json
import types

import pytest
import torch

import bench
from hqq_b200 import harness, tune


class FakeGraph:
    def replay(self):
        pass


class FakeModel:
    built = []
    TUNABLE = harness.DecodeModel.TUNABLE

    def __init__(self, shape, **kw):
        self.kw, self.shape = kw, shape
        self.device, self.dtype, self.nbits, self.pair_silu = torch.device("cpu"), torch.float16, 4, True
        self.tok, self.pos, self.next_tok = torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long), torch.zeros(1, dtype=torch.long)
        self.graph, self.blocks, self.retuned = FakeGraph(), [], []
        FakeModel.built.append(self)

    def capture(self, warmup=3):
        return self.graph

    def retune(self, knobs=None, warmup=2):
        self.retuned.append(dict(knobs or {}))
        return self.graph

    def decode(self, feed_back=True):
        pass

    def bytes_per_token(self):
        return 4.98e9


class FakeEvent:
    def __init__(self, enable_timing=True):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 400.0


class FakeStream:
    cuda_stream = 0

    def synchronize(self):
        pass


GUARD = [{"knobs": {}, "us": 500.0, "digest": "aa"},
         {"knobs": {"HQQ_B200_WPF_MB": "48", "HQQ_B200_WPF_FROM": "o"}, "us": 400.0, "digest": "aa", "identical": True, "speedup": 1.25},
         {"knobs": {"HQQ_B200_D1_VARIANT": "7042"}, "error": "child exited (code -11)"}]
REPORT = {"selected": {"HQQ_B200_WPF_MB": "48", "HQQ_B200_WPF_FROM": "o"}, "tried": [{"knobs": {"HQQ_B200_WPF_MB": "48", "HQQ_B200_WPF_FROM": "o"}, "us": 1600.0,
          "identical": True, "guard_speedup": 1.25}], "default_us": 1900.0, "selected_us": 1600.0, "gain": 1900.0 / 1600.0}
PROBES = {"gemm_sweep": {"us": 120.0, "per": {"b4_4096x4096_M4096": {"us": 120.0, "TFLOPs": 1145.0, "route": 2}}, "digest": "n/a"},
          "bitpack": {"us": 30.0, "per": {"b4_dequantize_f16": {"us": 30.0, "GBps": 5000.0, "bytes": 150e6}}, "digest": "n/a"},
          "solver_fast": {"bit_identical": True, "speedup": 4.0},
          "gemm": {"default": {"us": 120.0, "per": {"4096x4096xM4096": {"us": 120.0, "TFLOPs": 1145.0}}, "digest": "d"},
                   "HQQ_B200_GEMM_VARIANT=dq16": {"per": {"4096x4096xM4096": {"TFLOPs": 1400.0}}, "bit_identical": True, "speedup": 1.22},
                   "HQQ_B200_GEMM_VARIANT=ld": {"per": {"4096x4096xM4096": {"TFLOPs": 1600.0}}, "bit_identical": False},
                   "HQQ_B200_GEMM_VARIANT=un512": {"error": "timeout after 45 s"}}}


@pytest.fixture
def fake_gpu(monkeypatch):
    for k in harness.DecodeModel.TUNABLE + ("HQQ_B200_AUTOTUNE",):
        monkeypatch.delenv(k, raising=False)
    FakeModel.built.clear()
    monkeypatch.setattr(harness, "DecodeModel", FakeModel)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda d=None: FakeStream())
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda d=None: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(bench, "ClockSampler", lambda i: types.SimpleNamespace(start=lambda: None, stop=lambda: {"sm_mhz": 1900.0, "sm_max_mhz": 1965.0, "reasons": []}))
    monkeypatch.setattr(bench, "kernel_roofline", lambda model, torch_, peaks: {"bound": "hbm", "achieved": 3000.0, "peak": peaks["hbm_gbs"], "frac": 0.45})
    monkeypatch.setattr(bench, "quantizer_roofline", lambda torch_, peaks, dev, reps=3, fast_ok=False: {"ms_per_block": 1.5, "fast_ok": fast_ok})
    monkeypatch.setattr(bench, "run_probes", lambda: json.loads(json.dumps(PROBES)))
    monkeypatch.setattr(bench, "cpu_quantizer_baseline", lambda: {"value": 0.01, "unit": "Gweights/s", "cores": 8, "kind": "port"})
    monkeypatch.setattr(bench, "cpu_reference_tokens_per_s", lambda budget_s=15.0: (0.3, {"cores": 8, "sample": "fake", "port": "fake"}))
    monkeypatch.setattr(tune, "guard_decode", lambda budget_s=90.0: json.loads(json.dumps(GUARD)))
    monkeypatch.setattr(tune, "choose_decode", lambda model, guard: dict(REPORT))


def _args(**kw):
    d = dict(gpus=1, steps=20, warmup=3, impl="hqq_b200", cache_len=0, layers=0, no_cpu_baseline=False, no_probes=False, no_autotune=False,
             autotune_budget=90.0, extras_deadline=60.0, worker=True, batch=1, model="8b")
    d.update(kw)
    return argparse.Namespace(**d)


def test_run_gpu_assembles_the_line_with_the_autotuner(fake_gpu, capsys):
    bench.run_gpu(_args(), 0, 1, 0)
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(out) == 2 and "preliminary" in json.loads(out[0])["extras"]   # worker mode: the measurement first, the full line last
    assert json.loads(out[0])["value"] == json.loads(out[1])["value"] and "extras" not in json.loads(out[1])
    d = json.loads(out[-1])
    assert d["metric"] == bench.METRIC and d["n_gpus"] == 1 and d["steps"] == 20 and d["higher_is_better"] is True
    assert d["value"] == pytest.approx(20 / 0.4) and d["e2e"]["value"] == pytest.approx(20 / 0.4)
    at = d["config"]["autotune"]
    assert at["selected"] == "WPF_FROM=o,WPF_MB=48" and at["gain_vs_default"] == pytest.approx(1900 / 1600)
    assert [g["knobs"] for g in at["guard"]] == ["default", "WPF_FROM=o,WPF_MB=48", "D1_VARIANT=7042"] and "error" in at["guard"][2]
    assert at["in_process"][0]["knobs"] == "WPF_FROM=o,WPF_MB=48"
    assert FakeModel.built[0].retuned[-1] == REPORT["selected"]           # the timed region runs under the selection
    assert d["gemm_sweep"]["frac"] == pytest.approx(1145.0 / d["gemm_sweep"]["peak"], rel=1e-3) and "gemm_sweep" not in d["experimental"]
    assert d["gemm_sweep"]["best_bit_identical_variant"]["knob"] == "HQQ_B200_GEMM_VARIANT=dq16"   # faster, but not identical: ignored
    assert d["bitpack"]["frac"] == pytest.approx(5000.0 / d["bitpack"]["peak"], rel=1e-3)
    assert d["quantizer"]["fast_ok"] is True and d["quantizer"]["cpu_baseline"]["kind"] == "port"
    assert d["cpu_baseline"]["cores"] == 8 and d["roofline"]["frac"] == 0.45 and d["clocks"]["reasons"] == []


def test_run_gpu_without_autotune_and_extras(fake_gpu, capsys):
    bench.run_gpu(_args(no_autotune=True, no_probes=True, no_cpu_baseline=True), 0, 1, 0)
    d = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert d["config"]["autotune"] is None and "gemm_sweep" not in d and "cpu_baseline" not in d
    assert FakeModel.built[0].retuned == []


def test_run_gpu_survives_a_failing_tuner(fake_gpu, capsys, monkeypatch):
    def boom(model, guard):
        raise RuntimeError("recapture failed")

    monkeypatch.setattr(tune, "choose_decode", boom)
    bench.run_gpu(_args(no_probes=True, no_cpu_baseline=True), 0, 1, 0)
    d = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert "recapture failed" in d["config"]["autotune"]["error"] and d["value"] > 0
    assert FakeModel.built[0].retuned[-1] == {}                           # back on the default kernels


def test_run_gpu_prints_exactly_one_line_when_it_is_not_a_supervised_worker(fake_gpu, capsys):
    bench.run_gpu(_args(worker=False, no_autotune=True, no_probes=True, no_cpu_baseline=True), 0, 1, 0)
    assert len([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]) == 1
