"""Throughput of the other BASELINE.json configurations on one B200 (not the bench line; numbers go to profiles/).

    python scripts/bench_configs.py            # C1 (ZS B/32 N-Caltech), C3 (FS adapter B/16 N-ImageNet), C4 (ZS L/14 N-ImageNet)
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import clip
from eventclip_b200.graph import GraphedClassifier
from eventclip_b200.models import FSCLIPClassifier, FTCLIPClassifier, ZSCLIPClassifier
from eventclip_b200.synth import SENSORS, synth_batch, synth_text_feats

dev = torch.device("cuda", 0)


def qargs(cfg):
    return dict(max_imgs=10, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True,
                count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])


def run(name, ds, arch, B, few_shot, steps=8):
    cfg = SENSORS[ds]
    C = clip.ARCHS[arch][4]
    model = clip.init_weights_(clip.CLIP(arch), seed=0).to(dev).eval()
    text = synth_text_feats(cfg["n_cls"], C, 1)
    cd = dict(clip_model=model, prompt="a point cloud image of a {}", class_names=None, agg_func="mean", text_feats=text)
    if few_shot:
        ad = dict(adapter_type="text-trans", in_dim=C, d_model=256, num_heads=4, ffn_dim=1024, norm_first=True, num_layers=2,
                  residual=0.95)
        m = FSCLIPClassifier(adapter_dict=ad, clip_dict=cd, loss_dict=dict(use_logits_loss=True, use_probs_loss=False))
    else:
        m = ZSCLIPClassifier(clip_dict=cd)
    m = m.to(dev).eval()
    m.attach_event_frontend(qargs(cfg), cfg["shape"], cfg["max_n"])
    ev1, off1 = synth_batch(ds, 8, 77)
    ev = np.concatenate([ev1] * (B // 8))
    off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off1), B // 8))]).astype(np.int64)
    evd, offt = torch.from_numpy(ev).to(dev), torch.from_numpy(off)
    torch.manual_seed(0)
    sel = m.event_frontend.draw_selection(off)
    g = GraphedClassifier(m, max_events=ev.shape[0])
    d = dict(events=evd, event_offsets=offt, sel_idx=sel)
    with torch.no_grad():
        for _ in range(3):
            out = g(d)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    with torch.no_grad():
        for _ in range(steps):
            out = g(d)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    nv = int(out["valid_masks"].sum())
    r = dict(config=name, dataset=ds, arch=arch, batch=B, valid_views=nv, ms_per_step=ms, samples_per_s=B / ms * 1e3,
             views_per_s=nv / ms * 1e3, encoder_tflops=clip.flops_per_image(arch) * nv / ms / 1e9)
    print(json.dumps(r))
    del g, m, model
    torch.cuda.empty_cache()
    return r


def run_finetune(name, ds, arch, B, lora="qkvo-16", steps=8, eager=False):
    """BASELINE config 5: one fine-tune step = events -> frames -> forward -> loss -> backward -> (all-reduce) -> Adam."""
    from eventclip_b200 import train
    from eventclip_b200 import _lib
    cfg = SENSORS[ds]
    C = clip.ARCHS[arch][4]
    model = clip.init_weights_(clip.CLIP(arch), seed=0).to(dev).eval()
    cd = dict(clip_model=model, prompt="a point cloud image of a {}", class_names=None, agg_func="mean", lora=lora,
              only_conv1=False, only_bias=False, only_ln=False, text_feats=synth_text_feats(cfg["n_cls"], C, 1))
    m = FTCLIPClassifier(adapter_dict=dict(adapter_type="text-identity", residual=True), clip_dict=cd,
                         loss_dict=dict(use_logits_loss=True, use_probs_loss=False)).to(dev).train()
    q = qargs(cfg)
    q["max_imgs"] = 2                                     # configs/ftclip/*: 2 views per sample in training
    m.attach_event_frontend(q, cfg["shape"], cfg["max_n"])
    ev1, off1 = synth_batch(ds, 8, 77)
    ev = np.concatenate([ev1] * (B // 8))
    off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off1), B // 8))]).astype(np.int64)
    evd = torch.from_numpy(ev).to(dev)
    labels = torch.randint(0, cfg["n_cls"], (B,), generator=torch.Generator().manual_seed(1)).to(dev)
    torch.manual_seed(0)
    sel = m.event_frontend.draw_selection(off)
    from eventclip_b200.graph import GraphedFineTuner
    tuner = train.FineTuner(m, lr=2e-5)
    stepper = tuner if eager else GraphedFineTuner(tuner, max_events=ev.shape[0])
    for _ in range(3):
        loss = stepper.step(evd, off, labels, sel=sel)
    torch.cuda.synchronize()
    l0 = _lib.LAUNCHES
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        loss = stepper.step(evd, off, labels, sel=sel)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / steps
    nv = int(tuner.last["out"]["valid_masks"].sum())
    r = dict(config=name, mode="eager" if eager else "cuda graph", dataset=ds, arch=arch, batch=B, valid_views=nv, ms_per_step=ms,
             samples_per_s=B / ms * 1e3, views_per_s=nv / ms * 1e3, trainable=tuner.flat.numel, loss=float(loss), launches_per_step=(_lib.LAUNCHES - l0) / steps,
             encoder_tflops_fwd_bwd=3 * clip.flops_per_image(arch) * nv / ms / 1e9,
             peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30)
    print(json.dumps(r))
    del stepper, tuner, m, model
    torch.cuda.empty_cache()
    return r


if __name__ == "__main__":
    if "--finetune-one" in sys.argv:      # for ncu launch lists: one configuration, eager launches
        run_finetune("C5 (eager, profiling run)", "n_caltech101", "ViT-B/16", 32, steps=2, eager=True)
        sys.exit(0)
    if "--finetune" in sys.argv:
        res = [run_finetune("C5 LoRA qkvo-16 fine-tune step ViT-B/16 N-Caltech101 (32 samples x 2 views)", "n_caltech101", "ViT-B/16", 32),
               run_finetune("full fine-tune step ViT-B/16 N-Caltech101 (lora=-1, configs/ftclip/*ncaltech*vitb16.py, 32 x 2 views)",
                            "n_caltech101", "ViT-B/16", 32, lora=-1),
               run_finetune("C5 at batch 128 (the 4-GPU global batch of configs/ftclip on one GPU)", "n_caltech101", "ViT-B/16", 128),
               run_finetune("LoRA qkvo-16 fine-tune step ViT-L/14 N-ImageNet (32 samples x 2 views, configs/ftclip/*lora16.py)", "n_imagenet", "ViT-L/14", 32)]
        json.dump(res, open("gpurun_out/bench_finetune.json", "w"), indent=1)
        sys.exit(0)
    res = [run("C1 zero-shot ViT-B/32 N-Caltech101 (5 views/sample)", "n_caltech101", "ViT-B/32", 64, False),
           run("C3 few-shot joint adapter ViT-B/16 N-ImageNet (2 views/sample, 1000 classes)", "n_imagenet", "ViT-B/16", 64, True),
           run("C4 zero-shot ViT-L/14 N-ImageNet (2 views/sample)", "n_imagenet", "ViT-L/14", 64, False)]
    json.dump(res, open("gpurun_out/bench_configs.json", "w"), indent=1)
