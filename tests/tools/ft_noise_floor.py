"""Diagnostic (not collected by pytest): how far do half-precision gradients of the fine-tune step sit from fp32?

Runs the fp32 oracle's autograd on the GPU twice -- plain fp32, and under torch.autocast(bfloat16) -- plus the library's
FineTuner, on the inputs of tests/test_train_gpu.py::test_fused_step_vs_oracle_autograd, and prints per-tensor relative
L2 errors against fp32.  The autocast column is the noise floor any bf16 pipeline shows on these random-init weights.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eventclip_b200 import clip, train                      # noqa: E402
from eventclip_b200.models import FTCLIPClassifier          # noqa: E402
from eventclip_b200.synth import SENSORS, synth_batch       # noqa: E402
from oracle import clip_oracle, heads_oracle                # noqa: E402
from oracle import event2img as orc                         # noqa: E402
from tests.test_oracle_models import lora_from_sd           # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main(seed=None, verbose=True):
    dev = torch.device("cuda", 0)
    if seed is not None:
        torch.manual_seed(seed)                 # lora_down is drawn from the global generator at injection
    ds, arch, B = "n_caltech101", "ViT-B/16", 4
    cfg = SENSORS[ds]
    q = dict(max_imgs=2, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True,
             count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])
    ev, off = synth_batch(ds, B, 900, kind="clustered", E=30000)
    oracle = clip_oracle.build_clip(arch, seed=41)
    text = clip_oracle.synth_text_feats(cfg["n_cls"], oracle.visual.output_dim, 8)
    m = clip.CLIP(arch)
    m.load_state_dict(oracle.state_dict())
    cd = dict(clip_model=m.to(dev).eval(), prompt="a {}", class_names=None, agg_func="mean", lora="qkvo-16",
              only_conv1=False, only_bias=False, only_ln=False, text_feats=text)
    ft = FTCLIPClassifier(adapter_dict=dict(adapter_type="text-identity", residual=True), clip_dict=cd,
                          loss_dict=dict(use_logits_loss=True, use_probs_loss=False)).to(dev)
    ft.attach_event_frontend(q, cfg["shape"], cfg["max_n"])
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in ft.named_parameters():
            if "lora_up" in n:
                p.copy_((0.02 * torch.randn(p.shape, generator=gen)).to(dev))
    sd = {k: v.detach().cpu().clone() for k, v in ft.state_dict().items() if "lora" in k or k == "text_feats"}
    labels = torch.tensor([5, 17, 99, 0])
    tuner = train.FineTuner(ft.train(), lr=5e-4)
    sel = np.tile(np.arange(2, dtype=np.int32), (B, 1))
    tuner.forward_backward(torch.from_numpy(ev).to(dev), off, labels, sel=sel)
    imgs, valids = [], []
    for b in range(B):
        im, va, _ = orc.event2img_sample(ev[off[b]:off[b + 1]], cfg["shape"], cfg["N"], 2, cfg["count_non_zero"],
                                         cfg["background_mask"], sel=sel[b])
        imgs.append(im)
        valids.append(va)
    imgs, valid = torch.from_numpy(np.stack(imgs)), torch.from_numpy(np.stack(valids))
    oracle = oracle.to(dev)
    res = {}
    for mode in ("fp32", "autocast"):
        sdd = {k: v.to(dev) for k, v in sd.items()}
        lora, names = lora_from_sd(sdd, 12)
        tp = sdd["text_feats"].clone().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "autocast")):
            feats_loss, _ = heads_oracle.ft_train_loss(oracle.visual, lora, imgs[valid].to(dev), valid.to(dev), tp, labels.to(dev),
                                                       100.0, "mean")
        feats_loss.backward()
        res[mode] = ({names[k][j]: t[j].grad for k, t in lora.items() for j in (0, 1)}, feats_loss.item())
    named = dict(ft.named_parameters())
    if verbose:
        print("loss fp32 %.5f autocast %.5f" % (res["fp32"][1], res["autocast"][1]))
        print("%-70s %8s %8s" % ("tensor", "autocast", "library"))
    qk = {"autocast": [], "library": []}
    cat = {"ref": [], "autocast": [], "library": []}
    for nm, gref in res["fp32"][0].items():
        ra, rl = rel(res["autocast"][0][nm], gref), rel(tuner._grad_view(named[nm]), gref)
        if verbose:
            print("%-70s %8.4f %8.4f" % (nm[len("model.visual.transformer."):], ra, rl))
        if nm.endswith(("_q", "_k")):
            qk["autocast"].append(ra)
            qk["library"].append(rl)
            cat["ref"].append(gref.reshape(-1).float().cpu())
            cat["autocast"].append(res["autocast"][0][nm].reshape(-1).float().cpu())
            cat["library"].append(tuner._grad_view(named[nm]).reshape(-1).float().cpu())
    ref = torch.cat(cat["ref"])
    print("seed %s: q/k factor gradients vs fp32 -- autocast max %.3f all %.3f | library max %.3f all %.3f" % (
        seed, max(qk["autocast"]), rel(torch.cat(cat["autocast"]), ref), max(qk["library"]), rel(torch.cat(cat["library"]), ref)))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        for sd in range(int(sys.argv[1])):
            main(seed=sd, verbose=False)
    else:
        main(seed=0)
