"""A few graph-replayed steps of one BASELINE config for ncu launch lists:  python scripts/prof_config.py C3 [steps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from eventclip_b200.graph import GraphedClassifier

name = sys.argv[1] if len(sys.argv) > 1 else "C3"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
w = bench.Workload(name, dev, 0, n_batches=1, uniq=8 if bench.CONFIGS[name]["ds"] == "n_imagenet" else None)
g = GraphedClassifier(w.cls, max_events=w.max_events)
with torch.no_grad():
    for i in range(steps + 2):
        g(w.data(0))
torch.cuda.synchronize()
print("done", name)
