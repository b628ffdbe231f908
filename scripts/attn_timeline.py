import os, sys, torch
sys.path.insert(0, os.getcwd())
from eventclip_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
n_img, L, heads = 256, 197, 12
d = heads * 64
qkv = torch.randn(n_img * L, 3 * d, device=dev).to(torch.float16)
out = torch.empty(n_img * L, d, device=dev, dtype=torch.float16)
for _ in range(2):
    ops.attention(qkv, out, n_img, L, heads)
torch.cuda.synchronize()
