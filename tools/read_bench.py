"""Pretty-print a bench.py result line (or a driver BENCH_rNN.json / SCALE_rNN.json holding such lines): headline, roofline per launch
group, the decode autotuner's verdicts, the experimental probes, the GEMM / bit-packing sweeps, the quantizer and CPU baselines.

    python bench.py > line.json; python tools/read_bench.py line.json        # or: python tools/read_bench.py BENCH_r01.json
"""
import json
import sys


def lines_of(path):
    text = open(path).read()
    out = []

    def walk(o):
        if isinstance(o, dict):
            if "metric" in o and "value" in o:
                out.append(o)
            else:
                for v in o.values():
                    walk(v)
        elif isinstance(o, list):
            for v in o:
                walk(v)
        elif isinstance(o, str) and o.lstrip().startswith("{") and '"metric"' in o:
            for ln in o.splitlines():
                try:
                    walk(json.loads(ln))
                except ValueError:
                    pass

    try:
        walk(json.loads(text))
    except ValueError:
        for ln in text.splitlines():
            ln = ln.strip()
            if ln.startswith("{"):
                try:
                    walk(json.loads(ln))
                except ValueError:
                    pass
    return out


def f(v, nd=1):
    return "-" if v is None else (f"{v:.{nd}f}" if isinstance(v, (int, float)) else str(v))


def show(d):
    print(f"== {d.get('impl', 'hqq_b200')}  {d['metric']}  N={d.get('n_gpus')}  value {f(d['value'], 2)} {d.get('unit')}  "
          f"e2e {f((d.get('e2e') or {}).get('value'), 2)}  ms/step {f(d.get('ms_per_step'), 3)}  launches {d.get('gpu_launches')}")
    print(f"   clocks {d.get('clocks')}")
    at = (d.get("config") or {}).get("autotune")
    if at:
        if "error" in at:
            print(f"   autotune: ERROR {at['error']}")
        else:
            print(f"   autotune: selected {at['selected']}  gain {f(at['gain_vs_default'], 3)}x  ({f(at['default_us'])} -> {f(at['selected_us'])} us/step, {at.get('seconds')} s)")
            for g in at.get("guard", []):
                print(f"      guard  {g['knobs']:48s} " + (f"ERROR {g['error']}" if "error" in g else f"{f(g.get('us'))} us  x{f(g.get('speedup', 1.0), 3)}  identical={g.get('identical', '-')}"))
            for t in at.get("in_process", []):
                print(f"      model  {t['knobs']:48s} " + (f"ERROR {t['error']}" if "error" in t else f"{f(t.get('us'))} us  identical={t.get('identical')}"))
    r = d.get("roofline")
    if r:
        print(f"   roofline ({r.get('kernel')}): {f(r['achieved'])} / {f(r['peak'])} {r.get('unit')} = {f(100 * r['frac'])} %  traffic {r.get('traffic')}")
        for k, v in (r.get("per_launch_group") or {}).items():
            print(f"      {k:8s} {f(v['us'], 2)} us  {f(v['GBps'])} GB/s")
    for name in ("gemm_sweep", "bitpack"):
        s = d.get(name)
        if isinstance(s, dict) and "per" in s:
            print(f"   {name}: headline {s.get('headline')} {f(s.get('achieved'))} {s.get('unit')} = {f(100 * s.get('frac', 0))} % of {f(s.get('peak'))}"
                  + (f"   best identical variant: {s['best_bit_identical_variant']}" if "best_bit_identical_variant" in s else ""))
            for k, v in s["per"].items():
                extra = " ".join(f"{kk}={f(vv, 3) if isinstance(vv, float) else vv}" for kk, vv in v.items() if kk not in ("us",))
                print(f"      {k:28s} {f(v.get('us'), 1):>9s} us  {extra}")
        elif s:
            print(f"   {name}: {s}")
    q = d.get("quantizer")
    if q:
        print(f"   quantizer: {q}")
    for k, v in (d.get("experimental") or {}).items():
        if isinstance(v, dict) and "default" in v and all(isinstance(x, dict) for x in v.values()):
            for kk, vv in v.items():
                print(f"   experimental {k:12s} {kk:36s} " + ("ERROR " + str(vv.get("error") or vv.get("skipped")) if ("error" in vv or "skipped" in vv)
                      else f"{f(vv.get('us'))} us  identical={vv.get('bit_identical', '-')}  x{f(vv.get('speedup', 1.0), 3)}  rel_err={vv.get('rel_err', '-')}"))
        else:
            print(f"   experimental {k:12s} {json.dumps(v)[:300]}")
    if d.get("cpu_baseline"):
        print(f"   cpu_baseline: {d['cpu_baseline']}")
    for k in ("extras", "extras_error"):
        if k in d:
            print(f"   {k}: {d[k]}")


if __name__ == "__main__":
    for p in sys.argv[1:]:
        for d in lines_of(p):
            show(d)
