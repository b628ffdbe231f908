"""Prints the interesting parts of a bench.py JSON line: python scripts/show_bench.py gpurun_out/bench.json"""
import json
import sys

d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
for k in ("value", "ms_per_step", "dtype", "e2e", "gpu_launches", "clocks", "parity", "encoder", "cpu_baseline", "accuracy_counters"):
    print(k, json.dumps(d.get(k))[:1600])
r = d.get("roofline") or {}
print("roofline", r.get("achieved"), r.get("frac"), r.get("share_of_step"))
for k, v in (r.get("by_shape") or {}).items():
    print("  ", k, {a: round(b, 1) for a, b in v.items()})
for ds, v in (d.get("event2img") or {}).items():
    print(ds, "frac", round(v["frac"], 3), "gev", round(v["gevents_per_s"], 1),
          {k: (round(x["gevents_per_s"], 1), round(x["frac"], 3)) for k, x in v["by_stream_kind"].items()}, v.get("slowdown_vs_uniform"))
for k, v in (d.get("other_configs") or {}).items():
    print(k, json.dumps(v)[:1700])
