"""Authoring-box timing of the REAL reference front end (datasets/vis.py::events2frames + PIL/torchvision preprocess, imported
unmodified from /root/reference) next to the oracle port that bench.py's CPU arm runs (C event2img), on the bench workload.
The CLIP tower is third-party to the reference and not vendored, so both sides share the restated fp32 PyTorch tower; the
difference between the arms is the events -> frames stage.  Only runs where /root/reference exists.

    python scripts/time_reference_vs_port.py
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200.synth import SENSORS, synth_labeled_batch
from oracle import clip_oracle, ref_import
from oracle import event2img as orc

assert ref_import.available(), "needs /root/reference"
vis = ref_import.load_vis()
import torchvision.transforms as T
pre = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224), lambda im: im.convert("RGB"),
                 T.ToTensor(), T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])
from PIL import Image

cores = os.cpu_count()
torch.set_num_threads(cores)
for ds, arch, B in (("n_cars", "ViT-B/16", 32), ("n_caltech101", "ViT-B/32", 32)):
    cfg = SENSORS[ds]
    ev, off, _ = synth_labeled_batch(ds, B, 9000)
    Tn = orc.max_imgs(cfg["max_n"], cfg["N"], 10)
    clipm = clip_oracle.build_clip(arch, seed=0)

    def ref_frames():
        out = []
        for b in range(B):
            fr = vis.events2frames(ev[off[b]:off[b + 1]], split_method="event_count", convert_method="event_histogram",
                                   shape=cfg["shape"], N=cfg["N"], grayscale=True, count_non_zero=cfg["count_non_zero"],
                                   background_mask=cfg["background_mask"])
            out.append(torch.stack([pre(Image.fromarray(f)) for f in fr[:Tn]]))
        return out

    def port_frames():
        return [torch.from_numpy(orc.event2img_sample(ev[off[b]:off[b + 1]], cfg["shape"], cfg["N"], Tn, cfg["count_non_zero"],
                                                      cfg["background_mask"])[0]) for b in range(B)]

    res = {}
    for name, fn in (("reference vis.py + PIL", ref_frames), ("oracle port (C)", port_frames)):
        fn()
        t = []
        for _ in range(3):
            t0 = time.perf_counter()
            frames = fn()
            t.append(time.perf_counter() - t0)
        res[name] = min(t)
    a, b = ref_frames(), port_frames()
    same = all(torch.equal(x, y[:x.shape[0]]) for x, y in zip(a, b))
    imgs = torch.cat([x for x in a])
    with torch.no_grad():
        clipm.encode_image(imgs[:8])
        t0 = time.perf_counter()
        clipm.encode_image(imgs)
        enc = time.perf_counter() - t0
    print(f"{ds} {arch} batch {B} ({imgs.shape[0]} views), {cores} cores: frames reference {res['reference vis.py + PIL'] * 1e3:.1f} ms, "
          f"port {res['oracle port (C)'] * 1e3:.1f} ms (identical tensors: {same}); encoder fp32 {enc * 1e3:.0f} ms -> "
          f"samples/s reference front end {B / (res['reference vis.py + PIL'] + enc):.1f}, port {B / (res['oracle port (C)'] + enc):.1f}")
