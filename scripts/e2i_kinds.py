"""One launch of the fused event kernel per (sensor, stream kind) at the bench's batch sizes, for ncu metric passes:

    ncu --metrics <list> --clock-control none -k regex:event2img --csv --log-file out.csv python scripts/e2i_kinds.py
Order of the launches: for each sensor (n_caltech101, n_cars, n_imagenet): uniform, clustered, hotpixel; each launched twice
(the second launch of a pair is the warm one)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import ops
from eventclip_b200.datasets import Event2Image
from eventclip_b200.synth import SENSORS, synth_batch

dev = torch.device("cuda", 0)
for ds, B in (("n_caltech101", 296), ("n_cars", 2072), ("n_imagenet", 144)):
    cfg = SENSORS[ds]
    q = dict(max_imgs=10, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True,
             count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])
    e2i = Event2Image(q, cfg["shape"], cfg["max_n"])
    T = e2i.max_imgs
    sel = np.tile(np.arange(T, dtype=np.int32), (B, 1))
    for kind in ("uniform", "clustered", "hotpixel"):
        ev1, off1 = synth_batch(ds, 8, 100, kind=kind)
        evs = np.concatenate([ev1] * (B // 8))
        off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off1), B // 8))]).astype(np.int64)
        evd = torch.from_numpy(evs).to(dev)
        frames, valid, chunks, nv = ops.plan_frames(off, e2i.N, T, sel=sel, compact=True)
        fd = frames.to(dev)
        outbuf = torch.zeros((nv * 196, 768), dtype=torch.bfloat16, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        for _ in range(2):
            ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="patch", patch=16, ldk=768,
                          out_tensor=outbuf, status=status)
        torch.cuda.synchronize()
        print(ds, kind, "frames", nv, "events", int(np.frombuffer(frames.numpy().tobytes(), dtype=[("s", "<i8"), ("n", "<i4"), ("o", "<i4")])["n"].sum()), flush=True)
        del evd, outbuf
