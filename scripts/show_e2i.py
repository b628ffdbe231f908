"""Prints the event2img block of `python bench.py --e2i-only` (stdin) as one line per sensor."""
import json
import sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
for k, v in d.items():
    kinds = {kk: (round(vv["gevents_per_s"], 1), round(vv["frac"], 3)) for kk, vv in v.get("by_stream_kind", {}).items()}
    print(k, "ms %.4f" % v["ms"], "frac", round(v["frac"], 3), "Gev/s", round(v["gevents_per_s"], 1), kinds, v.get("geometry"))
