"""Which rounding points carry the encoder's error on event frames?  Centred (batch-mean-removed) relative L2 of the image
features against the fp32 oracle for the inference forward's switches (residual stream fp16 / fp32, LayerNorm folded or not)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from eventclip_b200 import clip
from eventclip_b200.synth import SENSORS, synth_labeled_batch
from oracle import clip_oracle

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from parity_util import rel_l2, rel_l2_centered

dev = torch.device("cuda", 0)
for ds, arch, n in (("n_cars", "ViT-B/16", 32), ("n_caltech101", "ViT-B/32", 16)):
    cfg = SENSORS[ds]
    ev, off, _ = synth_labeled_batch(ds, n, 4242, E=min(cfg["E"], 30000))
    imgs, valid = bench.oracle_frames(ev, off, cfg, 1)
    oracle = clip_oracle.build_clip(arch, seed=0)
    model = clip.CLIP(arch)
    model.load_state_dict(oracle.state_dict())
    model = model.to(dev).eval()
    x = imgs[:, 0]
    with torch.no_grad():
        ref = oracle.encode_image(x)
        for op, res, fold, split in ((torch.float16, torch.float16, True, True), (torch.float16, torch.float16, True, False),
                                     (torch.float16, torch.float16, False, False), (torch.float16, torch.float32, False, False),
                                     (torch.bfloat16, torch.float16, True, True), (torch.bfloat16, torch.float16, True, False),
                                     (torch.bfloat16, torch.float16, False, False), (torch.bfloat16, torch.float32, False, False)):
            model.visual.operand_dtype, model.visual.residual_dtype, model.visual.fold_ln = op, res, fold
            model.visual.residual_split = split
            model.visual.invalidate_packed()
            got = model.encode_image(x.to(dev)).cpu()
            print(ds, arch, "operands", str(op).split(".")[1], "residual", str(res).split(".")[1] + ("x2 (hi, lo)" if split else ""), "fold_ln", fold,
                  "rel_l2 %.3e centred %.3f" % (rel_l2(got, ref), rel_l2_centered(got, ref)), flush=True)
