"""Diagnostic: repeats FineTuner.forward_backward on the same batch with the caching allocator poisoned by NaNs in
between; reports which gradient tensors are not bitwise reproducible (uninitialised reads / races)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eventclip_b200 import clip, train                      # noqa: E402
from eventclip_b200.models import FTCLIPClassifier          # noqa: E402
from eventclip_b200.synth import SENSORS, synth_batch, synth_text_feats       # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    ds, arch, B = "n_caltech101", sys.argv[1] if len(sys.argv) > 1 else "ViT-B/16", 4
    cfg = SENSORS[ds]
    q = dict(max_imgs=2, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True,
             count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])
    ev, off = synth_batch(ds, B, 900, kind="clustered", E=30000)
    m = clip.init_weights_(clip.CLIP(arch), seed=41).to(dev).eval()
    cd = dict(clip_model=m, prompt="a {}", class_names=None, agg_func="mean", lora="qkvo-16",
              only_conv1=False, only_bias=False, only_ln=False, text_feats=synth_text_feats(cfg["n_cls"], clip.ARCHS[arch][4], 8))
    ft = FTCLIPClassifier(adapter_dict=dict(adapter_type="text-identity", residual=True), clip_dict=cd,
                          loss_dict=dict(use_logits_loss=True, use_probs_loss=False)).to(dev)
    ft.attach_event_frontend(q, cfg["shape"], cfg["max_n"])
    gen = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for n, p in ft.named_parameters():
            if "lora_up" in n:
                p.copy_((0.02 * torch.randn(p.shape, generator=gen)).to(dev))
    labels = torch.tensor([5, 17, 99, 0])
    tuner = train.FineTuner(ft.train(), lr=5e-4)
    sel = np.tile(np.arange(2, dtype=np.int32), (B, 1))
    evd = torch.from_numpy(ev).to(dev)
    names = {id(p): n for n, p in ft.named_parameters()}
    ref = None
    n_it = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    n_bad = 0
    for it in range(n_it):
        if it % 8 == 0:
            poison = torch.full((1 << 28,), float("nan"), device=dev)       # 1 GiB of NaNs goes back to the allocator
            del poison
        loss = tuner.forward_backward(evd, off, labels, sel=sel)
        torch.cuda.synchronize()
        g = tuner.flat_g.clone()
        if ref is None:
            ref = g
            print("loss", loss.item(), "nan in grads:", bool(torch.isnan(g).any()))
            continue
        if not torch.equal(g, ref):
            bad = []
            for p in tuner.group0 + tuner.group1:
                a, b = tuner.flat.grad_view(p), None
                o, k, _ = tuner.flat._where[id(p)]
                if not torch.equal(g[o:o + k], ref[o:o + k]):
                    d = (g[o:o + k] - ref[o:o + k]).abs().max().item()
                    bad.append((names[id(p)].replace("model.visual.transformer.", ""), d))
            n_bad += 1
            print(f"iter {it}: loss {loss.item()} DIFFERS in {len(bad)} tensors; first: {bad[:6]}")
    print(f"{n_bad} of {n_it - 1} repeats differ")


if __name__ == "__main__":
    main()
