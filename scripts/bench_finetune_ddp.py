"""Fine-tune step (BASELINE config 5) on N GPUs of one node: per-rank batch 32 samples x 2 views, one NCCL all-reduce of the
flat gradient buffer per step (the reference's DDP does the same averaging).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        scripts/bench_finetune_ddp.py [--steps 10]
Prints one JSON line from rank 0: step time (CUDA events, max over ranks), whole-job samples/s, and whether every rank holds
bitwise identical parameters after the timed steps.
"""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eventclip_b200 import clip, train
from eventclip_b200.graph import GraphedFineTuner
from eventclip_b200.models import FTCLIPClassifier
from eventclip_b200.synth import SENSORS, synth_batch, synth_text_feats


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ds, arch, B = "n_caltech101", "ViT-B/16", args.batch
    cfg = SENSORS[ds]
    torch.manual_seed(0)                                    # identical initial parameters on every rank
    model = clip.init_weights_(clip.CLIP(arch), seed=0).to(dev).eval()
    cd = dict(clip_model=model, prompt="a point cloud image of a {}", class_names=None, agg_func="mean", lora="qkvo-16",
              only_conv1=False, only_bias=False, only_ln=False, text_feats=synth_text_feats(cfg["n_cls"], 512, 1))
    m = FTCLIPClassifier(adapter_dict=dict(adapter_type="text-identity", residual=True), clip_dict=cd,
                         loss_dict=dict(use_logits_loss=True, use_probs_loss=False)).to(dev).train()
    q = dict(max_imgs=2, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True,
             count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])
    m.attach_event_frontend(q, cfg["shape"], cfg["max_n"])
    ev1, off1 = synth_batch(ds, 8, 77 + 100 * rank)          # every rank trains on its own shard
    ev = np.concatenate([ev1] * (B // 8))
    off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off1), B // 8))]).astype(np.int64)
    evd = torch.from_numpy(ev).to(dev)
    labels = torch.randint(0, cfg["n_cls"], (B,), generator=torch.Generator().manual_seed(rank)).to(dev)
    sel = np.tile(np.arange(2, dtype=np.int32), (B, 1))
    tuner = train.FineTuner(m, lr=2e-5)
    stepper = GraphedFineTuner(tuner, max_events=ev.shape[0])
    for _ in range(args.warmup):
        stepper.step(evd, off, labels, sel=sel)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        loss = stepper.step(evd, off, labels, sel=sel)
    b.record()
    torch.cuda.synchronize()
    ms = torch.tensor([a.elapsed_time(b) / args.steps], device=dev)
    chk = torch.stack([tuner.flat_p.double().sum(), tuner.flat_p.double().abs().sum()])
    same = True
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(torch.equal(c, allc[0]) for c in allc)
    if rank == 0:
        print(json.dumps(dict(config="C5 LoRA qkvo-16 fine-tune step ViT-B/16 N-Caltech101", n_gpus=world, per_gpu_batch=B,
                              ms_per_step=ms.item(), samples_per_s=world * B / ms.item() * 1e3, loss_rank0=float(loss),
                              grad_bytes=tuner.flat.numel * 4, params_identical_across_ranks=same)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
