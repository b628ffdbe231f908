"""A few launches of the attention forward at the bench geometry (run under ncu, or alone for CUDA-event times):
    python scripts/prof_attn.py [n_img] [L] [heads] [fp16|bf16]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import ops

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 256
L = int(sys.argv[2]) if len(sys.argv) > 2 else 197
heads = int(sys.argv[3]) if len(sys.argv) > 3 else 12
dt = torch.bfloat16 if (len(sys.argv) > 4 and sys.argv[4] == "bf16") else torch.float16
dev = torch.device("cuda", 0)
torch.manual_seed(0)
d = heads * 64
qkv = (torch.randn(n_img * L, 3 * d, device=dev) * 1.0).to(dt)
out = torch.empty(n_img * L, d, device=dev, dtype=dt)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for _ in range(3):
    ops.attention(qkv, out, n_img, L, heads)
torch.cuda.synchronize()
ts = []
for _ in range(10):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ops.attention(qkv, out, n_img, L, heads)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b) * 1e3)
ts.sort()
fl = 4.0 * n_img * heads * L * L * 64
print(f"attention n_img={n_img} L={L} heads={heads} {dt}: median {ts[len(ts)//2]:.1f} us, min {ts[0]:.1f} us, "
      f"{fl / ts[len(ts)//2] / 1e6:.0f} TF/s")
