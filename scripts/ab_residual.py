"""A/B of the residual-stream formats on one box: bench headline (short) for EC_RESIDUAL = fp16 / fp16x2 / fp32."""
import json
import os
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for rep in range(2):
    for mode in ("fp16", "fp16x2", "fp32"):
        env = dict(os.environ, EC_RESIDUAL=mode)
        r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--quick", "--steps", "15", "--warmup", "4"], env=env, capture_output=True, text=True)
        try:
            d = json.loads(r.stdout.strip().splitlines()[-1])
            print(mode, "samples/s %.0f ms/step %.3f" % (d["value"], d["ms_per_step"]), flush=True)
        except Exception as ex:
            print(mode, "failed", repr(ex), r.stderr[-800:], flush=True)
