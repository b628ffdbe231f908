import sys, torch, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import bench
from eventclip_b200.graph import GraphedClassifier
dev = torch.device("cuda", 0)
w = bench.Workload("C2", dev, rank=0, n_batches=2)
g = GraphedClassifier(w.cls, max_events=w.max_events)
with torch.no_grad():
    for i in range(3):
        out = g(w.data(i % 2))
    torch.cuda.synchronize()
    patches = w.cls._last_patches
n = 32
P, G = 16, 14
def check(patches, b, tag):
    ref = w.oracle(b, n)
    x = ref["imgs"][ref["valid"]]
    gg = torch.round((x[:, :1].double() * 0.26862954 + 0.48145466) * 255.0).to(torch.float32)
    want = (gg / 128.0).reshape(n, G, P, G, P).permute(0, 1, 3, 2, 4).reshape(n * G * G, P * P)
    got = patches[: n * G * G, : P * P].float().cpu()
    bad = (got != want)
    print(tag, "shape", patches.shape, patches.dtype, "bad count", int(bad.sum()), "of", bad.numel())
    if bad.any():
        for r, c in bad.nonzero()[:10].tolist():
            print(r, c, got[r, c].item() * 128, want[r, c].item() * 128)
        rows = torch.unique(bad.nonzero()[:, 0])
        print("bad rows", rows.numel(), rows[:20].tolist(), "frames", torch.unique(rows // 196).tolist())
check(patches, 0, "graph vs batch0")
check(patches, 1, "graph vs batch1")
with torch.no_grad():
    w.cls(w.data(0))
check(w.cls._last_patches, 0, "eager vs batch0")
