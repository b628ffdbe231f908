import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import clip
arch = sys.argv[1] if len(sys.argv) > 1 else "ViT-L/14"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda", 0)
m = clip.init_weights_(clip.CLIP(arch), seed=0).to(dev).eval()
x = torch.randn(n, 3, 224, 224, device=dev).to(torch.bfloat16)
with torch.no_grad():
    for _ in range(3):
        y = m.encode_image(x)
torch.cuda.synchronize()
print(y.shape)
