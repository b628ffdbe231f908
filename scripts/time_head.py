"""CUDA-event time of ec_head alone:  python scripts/time_head.py B T C n_cls"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import ops
B, T, C, K = [int(a) for a in sys.argv[1:5]]
dev = torch.device("cuda", 0)
f = torch.randn(B * T, C, device=dev); v = torch.ones(B * T, dtype=torch.uint8, device=dev); t = torch.randn(K, C, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3): ops.head(f, v, t, B, T, 100.0, 0, "mean")
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.head(f, v, t, B, T, 100.0, 0, "mean"); b.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
print(f"ec_head B={B} T={T} C={C} n_cls={K}: median {ts[5]:.1f} us (includes four torch.empty calls)")
