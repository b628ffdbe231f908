"""The four GEMM shapes of a ViT-B/16 block at the bench's M (batch 256 x 197 tokens), a few launches each, for ncu captures
and quick CUDA-event timings:   python scripts/prof_gemm.py [reps]
Order per rep: in_proj (LayerNorm folded, N=2304), out_proj (+stats, fp16 residual), c_fc (folded, QuickGELU, N=3072), c_proj."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import ops

dev = torch.device("cuda", 0)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dt = torch.bfloat16 if os.environ.get("EC_OPERANDS", "fp16") == "bf16" else torch.float16
M, d = 256 * 197, 768
g = torch.Generator(device="cpu").manual_seed(0)
x = (torch.randn(M, d, generator=g)).half().to(dev)
parts = ops.gemm_stats_parts(d)
stats = torch.empty((M, parts, 2), dtype=torch.float32, device=dev)
ops.row_stats_f16(x, stats, parts)
wg_in = (torch.randn(3 * d, d, generator=g) * d ** -0.5).half().to(dev)
wg_fc = (torch.randn(4 * d, d, generator=g) * d ** -0.5).half().to(dev)
w_out = (torch.randn(d, d, generator=g) * d ** -0.5).to(dt).to(dev)
w_proj = (torch.randn(d, 4 * d, generator=g) * (4 * d) ** -0.5).to(dt).to(dev)
c_in, c_fc = torch.zeros(3 * d, device=dev), torch.zeros(4 * d, device=dev)
b_o, b_p = torch.zeros(d, device=dev), torch.zeros(d, device=dev)
qkv = torch.empty((M, 3 * d), dtype=dt, device=dev)
att = (torch.randn(M, d, generator=g) * 0.3).to(dt).to(dev)
hid = torch.empty((M, 4 * d), dtype=dt, device=dev)
names = ["in_proj ln N2304 K768", "out_proj stats N768 K768", "c_fc ln qgelu N3072 K768", "c_proj stats N768 K3072"]
fl = [2.0 * M * 3 * d * d, 2.0 * M * d * d, 2.0 * M * 4 * d * d, 2.0 * M * d * 4 * d]
fns = [lambda: ops.gemm_ln(x, wg_in, None, c_in, stats, parts, "bf16", out=qkv),
       lambda: ops.gemm_bf16_stats(att, w_out, b_o, x, stats),
       lambda: ops.gemm_ln(x, wg_fc, None, c_fc, stats, parts, "bf16_qgelu", out=hid),
       lambda: ops.gemm_bf16_stats(hid, w_proj, b_p, x, stats)]
ev = [[(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in fns] for _ in range(reps)]
for f in fns:
    f()
torch.cuda.synchronize()
for r in range(reps):
    for i, f in enumerate(fns):
        ev[r][i][0].record()
        f()
        ev[r][i][1].record()
torch.cuda.synchronize()
for i, n in enumerate(names):
    us = min(ev[r][i][0].elapsed_time(ev[r][i][1]) for r in range(reps)) * 1e3
    print(f"{n:28s} {us:8.1f} us  {fl[i] / us / 1e6:7.1f} TF/s")
