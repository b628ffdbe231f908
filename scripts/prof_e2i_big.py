import os, sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from eventclip_b200 import ops
from eventclip_b200.datasets import Event2Image
from eventclip_b200.synth import SENSORS, synth_batch
dev = torch.device("cuda", 0)
ds, B = "n_imagenet", 144
cfg = SENSORS[ds]
q = dict(max_imgs=2, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True, count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])
e2i = Event2Image(q, cfg["shape"], cfg["max_n"])
ev1, off1 = synth_batch(ds, 8, 100, kind="uniform")
evs = np.concatenate([ev1] * (B // 8)); off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off1), B // 8))]).astype(np.int64)
evd = torch.from_numpy(evs).to(dev)
T = e2i.max_imgs
sel = np.tile(np.arange(T, dtype=np.int32), (B, 1))
frames, valid, chunks, nv = ops.plan_frames(off, e2i.N, T, sel=sel, compact=True)
fd = frames.to(dev)
out = torch.zeros((nv * 196, 768), dtype=torch.bfloat16, device=dev)
st = torch.zeros(1, dtype=torch.int32, device=dev)
for _ in range(3):
    ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="patch", patch=16, ldk=768, out_tensor=out, status=st)
torch.cuda.synchronize()
print("frames", nv)
