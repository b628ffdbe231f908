"""CUDA-event time of ec_gemm_f32 alone (and its error against fp64):  python scripts/time_sgemm.py M N K"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import ops
M, N, K = [int(a) for a in sys.argv[1:4]]
dev = torch.device("cuda", 0)
A = torch.randn(M, K, device=dev); W = torch.randn(N, K, device=dev) * K ** -0.5; b = torch.randn(N, device=dev)
ref = torch.relu(A.double() @ W.double().t() + b.double())
out = ops.gemm_f32(A, W, b, None, 1)
err = ((out.double() - ref).abs().max() / ref.abs().max()).item()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(10):
    flush.zero_()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); ops.gemm_f32(A, W, b, None, 1); e.record(); torch.cuda.synchronize()
    ts.append(a.elapsed_time(e) * 1e3)
ts.sort()
print(f"ec_gemm_f32 M={M} N={N} K={K}: median {ts[5]:.1f} us, max rel err vs fp64 {err:.2e}")
