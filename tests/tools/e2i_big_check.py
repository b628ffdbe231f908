"""Small driver for compute-sanitizer runs of the band-exchange event kernel: one forced 2-CTA sensor with a multi-round
exchange and one N-ImageNet-shaped frame, each checked against the oracle.  (tests/test_event2img_gpu.py holds the real tests.)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eventclip_b200 import ops
from eventclip_b200.synth import synth_events
from oracle import event2img as orc

dev = torch.device("cuda", 0)
cases = [((128, 128), 3000, "force", "256", "clustered"), ((480, 640), 70000, None, None, "uniform")]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    cases = cases[:1]
for shape, N, force, cap, kind in cases:
    for k, v in (("EC_E2I_BIG", force), ("EC_E2I_BIG_CAP", cap)):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    ev = synth_events(shape, 2 * N + 5, 3, kind)
    frames, _, _, K = ops.plan_frames([0, len(ev)], N, 4, compact=True)
    img, status, dbg = ops.event2img(torch.from_numpy(ev).to(dev), frames.to(dev), shape, K, False, True, out="f32", debug=True)
    torch.cuda.synchronize()
    oimg, _, _ = orc.event2img_sample(ev, shape, N, K, False, True)
    i0, i1 = orc.split_event_count(len(ev), N)
    ok_c = all((dbg["counts"][k].cpu().numpy() == orc.histogram(ev[i0[k]:i1[k]], shape)).all() for k in range(K))
    print(shape, ops.event2img_geometry(shape), "K", K, "status", int(status.item()), "counts", ok_c,
          "img", bool((img.cpu().numpy() == oimg[:K]).all()), flush=True)
