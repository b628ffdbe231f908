#!/usr/bin/env python
"""bench.py -- events -> logits throughput of the EventCLIP hot path on B200 (and the reference CPU arm).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Headline workload (BASELINE.json configs[1], the configuration the metric is quoted on): zero-shot EventCLIP ViT-B/16 on
synthetic N-Cars-shaped streams (120x100 sensor, 4000 events/sample, 2 classes), batch 256 per GPU, bf16 tensor-core
math with fp32 accumulation, random-init CLIP weights (no checkpoints offline).  One step = one batch through
ec_event2img -> patch GEMM -> 12 ViT blocks -> proj -> head.  Rank r works on its own batch (samples shard by rank,
weak scaling); NCCL carries only the final prediction-counter all-reduce.

The streams are LABELLED synthetic samples with per-sample layouts (eventclip_b200.synth.synth_labeled_batch) and the text
features are calibrated on a separate batch so that the predictions split evenly (synth.calibrate_text_feats): a random-init
tower with Gaussian text features predicts one class for every input, which would make "top-1 identical" vacuous.

One JSON line on stdout (rank 0):
  value      samples/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public classifier API with HOST (pinned) buffers: H2D of the packed events and
             D2H of the predictions inside the timed region
  parity     the graph-replayed, timed path against the fp32 CPU oracle on a whole timed batch: patch rows bit-exact,
             logits rel-L2 (plain and with the batch mean removed), top-1 agreement, class split, smallest top-2 margin
  roofline   dominant kernel = the tcgen05 GEMM: algorithmic FLOPs of its launches / their summed in-graph durations,
             against the measured cuBLAS bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port (C event2img + fp32 PyTorch CLIP + head) on this host's cores, bounded sample
  event2img  secondary metric of BASELINE.json: Gevents/s of the fused kernel alone vs the HBM roofline, per sensor and per
             stream kind (uniform / clustered / hot pixel)
  other_configs  BASELINE.json configs[0], [2], [3], [4] (C1, C3, C4, C5) measured the same way, a few steps each, with
             their own parity check; under torchrun C4 and the C5 fine-tune step (gradient all-reduce) run at every N
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "event samples/s (events->logits)"
FALLBACK_PEAKS = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0)   # B200_PROFILING.md fallback

# BASELINE.json configs: C2 is the headline (configs[1]); the others are reported under "other_configs"
CONFIGS = {
    "C1": dict(ds="n_caltech101", arch="ViT-B/32", B=32, kind="zs",
               name="zero-shot ViT-B/32, N-Caltech101-shaped streams (240x180, 100k events/sample, 101 classes), batch 32"),
    "C2": dict(ds="n_cars", arch="ViT-B/16", B=256, kind="zs",
               name="zero-shot ViT-B/16, N-Cars-shaped streams (120x100, 4000 events/sample, 2 classes), batch 256"),
    "C3": dict(ds="n_imagenet", arch="ViT-B/16", B=64, kind="fs",
               name="few-shot joint adapter (text-trans, residual 0.95) ViT-B/16, N-ImageNet-shaped streams (640x480, 1M "
                    "events/sample, 2 of 14 chunks, 1000 classes), batch 64 per GPU"),
    "C4": dict(ds="n_imagenet", arch="ViT-L/14", B=64, kind="zs",
               name="zero-shot ViT-L/14, N-ImageNet-shaped streams (640x480, 1M events/sample, 2 of 14 chunks, 1000 classes), "
                    "batch 64 per GPU"),
    "C5": dict(ds="n_caltech101", arch="ViT-B/16", B=32, kind="ft",
               name="LoRA qkvo-16 fine-tune step ViT-B/16 (+ prompt-tuned text features), N-Caltech101-shaped streams, 32 samples "
                    "x 2 views per GPU, gradient all-reduce (4.9 MB) over NCCL"),
}
HEAD = "C2"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        d["_source"] = "measured"
        return d
    d = dict(FALLBACK_PEAKS)
    d["_source"] = "fallback"
    return d


def qargs(cfg, max_imgs=10):
    return dict(max_imgs=max_imgs, N=cfg["N"], split_method="event_count", convert_method="event_histogram",
                grayscale=True, count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed work (B200_PROFILING.md clocks line): started before the
    device-timed region and stopped after the end-to-end loop -- the same steps under the same load -- because the timed
    region alone (K x 9 ms) is shorter than nvidia-smi's first few sampling periods."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                                         text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.t.join(timeout=2)

    def summary(self):
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except (ValueError, IndexError):
                pass
        if not sm:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        return dict(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm))


# ------------------------------------------------------------------------------------------------ synthetic workload
def make_batch(cname, seed0, uniq=None):
    """One labelled packed batch of a config: `uniq` distinct samples tiled up to the batch size (the 1M-event N-ImageNet
    streams take ~0.15 s each to draw on the host).  Returns (events float32 [sum E,4], offsets int64 [B+1], labels)."""
    from eventclip_b200.synth import synth_labeled_batch
    c = CONFIGS[cname]
    B = c["B"]
    uniq = B if uniq is None else min(uniq, B)
    ev1, off1, lab1 = synth_labeled_batch(c["ds"], uniq, seed0)
    if uniq == B:
        return ev1, off1, lab1
    rep = (B + uniq - 1) // uniq
    ev = np.concatenate([ev1] * rep)
    lens = np.tile(np.diff(off1), rep)[:B]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    return ev[:off[-1]], off, np.tile(lab1, rep)[:B]


def draw_sel(e2i, off, seed):
    """The reference's torch.randperm(K)[:T] draw (event2img.py:83-86), seeded so that both arms use the same chunks."""
    g = torch.Generator().manual_seed(seed)
    return e2i.draw_selection(off, generator=g)


# ------------------------------------------------------------------------------------------------ CPU oracle legs
def oracle_frames(ev, off, cfg, T, sel=None):
    from oracle import event2img as orc
    imgs, valids = [], []
    for b in range(len(off) - 1):
        kw = {} if sel is None else dict(sel=sel[b], only_selected=True)
        im, va, _ = orc.event2img_sample(ev[off[b]:off[b + 1]], cfg["shape"], cfg["N"], T, cfg["count_non_zero"],
                                         cfg["background_mask"], **kw)
        imgs.append(im)
        valids.append(va)
    return torch.from_numpy(np.stack(imgs)), torch.from_numpy(np.stack(valids))


def oracle_forward(cname, oracle_clip, ev, off, T, sel, text, adapter_state=None, want_feats=False):
    """The oracle port of the whole path for one packed batch: C event2img + fp32 PyTorch CLIP + head
    (models/clip_cls.py:131-162 zero-shot, :308-350 few-shot)."""
    from eventclip_b200.synth import SENSORS
    from oracle import heads_oracle
    c = CONFIGS[cname]
    cfg = SENSORS[c["ds"]]
    imgs, valid = oracle_frames(ev, off, cfg, T, sel)
    with torch.no_grad():
        feats = oracle_clip.encode_image(imgs[valid])
        if c["kind"] == "zs":
            out = heads_oracle.zs_head(feats, valid, text, 100.0, "mean")
        else:
            adapter = None
            if adapter_state is not None:
                adapter = lambda f, v: heads_oracle.adapter_forward(adapter_state, f, v, num_heads=4, residual=0.95)
            out = heads_oracle.fs_head(feats, valid, text, 100.0, "mean", adapter)
    out["imgs"], out["valid"] = imgs, valid
    if want_feats:
        out["feats"] = feats
    return out


def parity_stats(got_logits, ref_logits, k=1):
    """GPU logits vs oracle logits of the same samples: plain and centred (batch mean removed per class) relative L2,
    top-1 agreement (all samples, and those whose oracle top-2 margin exceeds twice the largest logit error), class split."""
    g, r = got_logits.double().cpu(), ref_logits.double().cpu()
    err = (g - r)
    gc, rc = g - g.mean(0, keepdim=True), r - r.mean(0, keepdim=True)
    top2 = r.topk(2, dim=-1).values
    margin = (top2[:, 0] - top2[:, 1])
    tol = 2.0 * float(err.abs().max())
    agree = (g.argmax(-1) == r.argmax(-1))
    clear = margin > tol
    pred = r.argmax(-1)
    counts = torch.bincount(pred, minlength=r.shape[1])
    st = dict(samples=int(r.shape[0]), logits_rel_l2=float(err.norm() / r.norm()),
              logits_centered_rel_l2=float((gc - rc).norm() / rc.norm().clamp_min(1e-30)),
              max_abs_logit_err=float(err.abs().max()), logit_std_over_batch=float(rc.std()),
              top1_agree=float(agree.float().mean()), top1_disagreements=int((~agree).sum()),
              top1_agree_clear_margin=float(agree[clear].float().mean()) if bool(clear.any()) else None,
              clear_margin_samples=int(clear.sum()), margin_tol=tol,
              min_margin=float(margin.min()), median_margin=float(margin.median()),
              oracle_class_split=sorted(counts.tolist(), reverse=True)[:4], oracle_classes_predicted=int((counts > 0).sum()))
    if k > 1:
        gk = g.topk(k, -1).indices
        st["oracle_top1_in_gpu_top%d" % k] = float((gk == pred[:, None]).any(-1).float().mean())
    return st


def patches_bit_exact(patches, imgs, valid, P):
    """16-bit im2col rows written by the fused event kernel vs the oracle's frames, bitwise.  Three-channel formats: the
    round-to-nearest-even of the oracle's float32 tensor (bf16 or fp16).  Gray format (one plane, conv1 folded): the oracle's
    resampled byte / 128, recovered from its float32 tensor and checked to reproduce all three channels exactly."""
    x = imgs[valid]                                                  # [nv, 3, 224, 224]
    n, G = x.shape[0], 224 // P
    if patches.shape[1] < 3 * P * P:
        mean = torch.tensor([0.48145466, 0.4578275, 0.40821073], dtype=torch.float32).view(1, 3, 1, 1)
        std = torch.tensor([0.26862954, 0.26130258, 0.27577711], dtype=torch.float32).view(1, 3, 1, 1)
        g = torch.round((x[:, :1].double() * 0.26862954 + 0.48145466) * 255.0).abs().to(torch.float32)      # the byte behind channel 0 (abs: no -0.0)
        if not torch.equal((g / 255.0 - mean) / std, x):                # ToTensor + Normalize in float32, as the oracle does
            return False
        ref = (g / 128.0).reshape(n, G, P, G, P).permute(0, 1, 3, 2, 4).reshape(n * G * G, P * P).to(patches.dtype)
        got = patches[: n * G * G, : P * P].cpu()
        return bool(torch.equal(got.view(torch.int16), ref.view(torch.int16)))
    ref = x.reshape(n, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(n * G * G, 3 * P * P).to(patches.dtype)
    got = patches[: n * G * G, : 3 * P * P].cpu()
    return bool(torch.equal(got.view(torch.int16), ref.view(torch.int16)))


def run_cpu(samples_per_step, steps, warmup):
    """CPU arm: the oracle port of the headline workload on all host cores (bounded sample)."""
    from eventclip_b200.datasets import Event2Image
    from eventclip_b200.synth import SENSORS, calibrate_text_feats, synth_labeled_batch
    from oracle import clip_oracle
    c = CONFIGS[HEAD]
    cfg = SENSORS[c["ds"]]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle_clip = clip_oracle.build_clip(c["arch"], seed=0)
    T = Event2Image(qargs(cfg), cfg["shape"], cfg["max_n"]).max_imgs
    evc, offc, _ = synth_labeled_batch(c["ds"], 16, 555)              # calibration of the synthetic text features
    text = calibrate_text_feats(oracle_forward(HEAD, oracle_clip, evc, offc, T, None, torch.zeros(2, 512), want_feats=True)["feats"],
                                cfg["n_cls"])
    ev, off, _ = synth_labeled_batch(c["ds"], samples_per_step, 9000)
    for _ in range(warmup):
        oracle_forward(HEAD, oracle_clip, ev, off, T, None, text)
    t0 = time.perf_counter()
    for _ in range(steps):
        oracle_forward(HEAD, oracle_clip, ev, off, T, None, text)
    dt = time.perf_counter() - t0
    return samples_per_step * steps / dt, dt / steps * 1e3, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps = 32
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    value, ms, cores = run_cpu(sps, steps, warmup)
    arch = CONFIGS[HEAD]["arch"]
    cb = dict(value=value, unit="samples/s", cores=cores, kind="port",
              sample=f"{sps} samples/step x {steps} steps of the bench workload (oracle: C event2img + fp32 PyTorch CLIP {arch} + head)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"zero-shot {arch}, N-Cars-shaped labelled streams 120x100, 4000 events/sample, 2 classes, "
                               f"{sps} samples per CPU step (bounded sample of the batch-256 workload)"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ event2img metric
def event2img_cpu_mev(ds, n_samples):
    """SURVEY 8(d) CPU baseline of the second metric: the oracle's C port of events -> float32 frames on the host cores
    (one sample per worker thread, ctypes releases the GIL), Mevents/s of events histogrammed.  Bounded sample."""
    from concurrent.futures import ThreadPoolExecutor
    from eventclip_b200.datasets import Event2Image
    from eventclip_b200.synth import SENSORS, synth_batch
    from oracle import event2img as orc
    cfg = SENSORS[ds]
    T = Event2Image(qargs(cfg), cfg["shape"], cfg["max_n"]).max_imgs
    ev, off = synth_batch(ds, n_samples, 100)
    cores = os.cpu_count() or 1
    work = lambda b: orc.event2img_sample(ev[off[b]:off[b + 1]], cfg["shape"], cfg["N"], T, cfg["count_non_zero"],
                                          cfg["background_mask"], sel=np.arange(T), only_selected=True)[2]
    work(0)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        ks = list(ex.map(work, range(n_samples)))
    dt = time.perf_counter() - t0
    used = sum(min(int(k), T) * min(cfg["N"], int(off[b + 1] - off[b])) for b, k in enumerate(ks))
    return dict(mevents_per_s=used / dt / 1e6, cores=cores, kind="port", sample=f"{n_samples} samples, {used} events histogrammed")


def event2img_metric(dev, pk, kinds=("uniform", "clustered", "hotpixel"), cpu=True):
    """BASELINE.json's second metric: Gevents/s of the fused kernel alone (bf16 patch rows out), HBM roofline.  The headline
    figure of each sensor is the uniform stream; `by_stream_kind` times the same launch on clustered streams (50 Gaussian
    blobs: neighbouring bins, bank pressure) and hot-pixel streams (2 % of the events on one pixel: same-address atomics)."""
    from eventclip_b200 import ops
    from eventclip_b200.datasets import Event2Image
    from eventclip_b200.synth import SENSORS, synth_batch
    out = {}
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(run, reps):
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        s.record()
        for _ in range(reps):
            run()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    # batch sizes give whole waves of clusters on 148 SMs (1480 / 2072 / 296 frames) and inputs far beyond L2
    for ds, B, reps in (("n_caltech101", 296, 5), ("n_cars", 2072, 5), ("n_imagenet", 144, 3)):
        cfg = SENSORS[ds]
        e2i = Event2Image(qargs(cfg), cfg["shape"], cfg["max_n"])
        T = e2i.max_imgs
        sel = np.tile(np.arange(T, dtype=np.int32), (B, 1))
        by_kind = {}
        for kind in kinds:
            ev1, off1 = synth_batch(ds, 8, 100, kind=kind)
            evs = np.concatenate([ev1] * (B // 8))
            off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off1), B // 8))]).astype(np.int64)
            evd = torch.from_numpy(evs).to(dev)
            frames, valid, chunks, nv = ops.plan_frames(off, e2i.N, T, sel=sel, compact=True)
            fd = frames.to(dev)
            rec = np.frombuffer(frames.numpy().tobytes(), dtype=[("s", "<i8"), ("n", "<i4"), ("o", "<i4")])
            ev_read = int(rec["n"].sum())
            outbuf = torch.zeros((nv * 196, 768), dtype=torch.bfloat16, device=dev)
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            run = lambda: ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="patch",
                                        patch=16, ldk=768, out_tensor=outbuf, status=status)
            ms = timed(run, reps)
            byts = 16 * ev_read + nv * 3 * 224 * 224 * 2
            by_kind[kind] = dict(ms=ms, gevents_per_s=ev_read / ms / 1e6, achieved_gbs=byts / ms / 1e6,
                                 frac=byts / ms / 1e6 / pk["hbm_gbs"], status=int(status.item()))
            if kind != "uniform":
                del evd, outbuf
                continue
            # row F2: the same frames from the compact wire format (4 bytes per event); the metric's byte count changes with
            # it, so both conventions are reported: against the reference's 16-byte events and against the bytes really read
            words = ops.pack_events(evd, cfg["shape"])
            runc = lambda: ops.event2img(words, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="patch",
                                         patch=16, ldk=768, out_tensor=outbuf, status=status)
            msc = timed(runc, reps)
            bytc = 4 * ev_read + nv * 3 * 224 * 224 * 2
            compact = dict(ms=msc, gevents_per_s=ev_read / msc / 1e6, frac_reference_bytes=byts / msc / 1e6 / pk["hbm_gbs"],
                           frac_compact_bytes=bytc / msc / 1e6 / pk["hbm_gbs"], input_mb=words.numel() * 4 / 1e6)
            del words
            # the format the classifiers really consume since round 2: ONE exact gray plane per patch row, conv1 folded onto it
            # (clip.packed_gray): a third of the output bytes.  Reported against the reference's byte count (three normalised
            # float channels -> 16-bit) and against the bytes this format really writes.
            outg = torch.zeros((nv * 196, 256), dtype=torch.float16, device=dev)
            rung = lambda: ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="gray_f16",
                                         patch=16, ldk=256, out_tensor=outg, status=status)
            msg = timed(rung, reps)
            bytg = 16 * ev_read + nv * 224 * 224 * 2
            gray = dict(ms=msg, gevents_per_s=ev_read / msg / 1e6, frac_reference_bytes=byts / msg / 1e6 / pk["hbm_gbs"],
                        frac_gray_bytes=bytg / msg / 1e6 / pk["hbm_gbs"], output_mb=nv * 224 * 224 * 2 / 1e6)
            del outg
            out[ds] = dict(compact_wire_format=compact, gray_patch_format=gray, frames=int(nv), events_histogrammed=ev_read, events_in_streams=int(off[-1]), ms=ms,
                           gevents_per_s=ev_read / ms / 1e6, gevents_per_s_stream=int(off[-1]) / ms / 1e6,
                           algorithmic_bytes=byts, achieved_gbs=byts / ms / 1e6, frac=byts / ms / 1e6 / pk["hbm_gbs"],
                           input_mb=evs.nbytes / 1e6, geometry=ops.event2img_geometry(cfg["shape"]))
            del evd, outbuf
        out[ds]["by_stream_kind"] = by_kind
        u = by_kind.get("uniform", {}).get("gevents_per_s")
        if u:
            out[ds]["slowdown_vs_uniform"] = {k: 1.0 - v["gevents_per_s"] / u for k, v in by_kind.items() if k != "uniform"}
        if cpu:
            try:
                out[ds]["cpu_baseline"] = event2img_cpu_mev(ds, {"n_caltech101": 64, "n_cars": 512, "n_imagenet": 16}[ds])
            except Exception as ex:      # the oracle is test infrastructure: never let it take the bench line down
                out[ds]["cpu_baseline"] = dict(error=str(ex)[:200])
    return out


# ------------------------------------------------------------------------------------------------ B200 arm
class Workload:
    """Model + classifier + batches of one BASELINE config on one GPU, text features calibrated through the library's own
    event kernel and encoder (the calibration batch is disjoint from the timed ones)."""

    def __init__(self, cname, dev, rank, n_batches, uniq=None):
        from eventclip_b200 import clip, ops
        from eventclip_b200.datasets import Event2Image
        from eventclip_b200.models import FSCLIPClassifier, FTCLIPClassifier, ZSCLIPClassifier
        from eventclip_b200.synth import SENSORS, calibrate_text_feats
        c = CONFIGS[cname]
        self.cname, self.c, self.dev = cname, c, dev
        self.cfg = cfg = SENSORS[c["ds"]]
        self.B = c["B"]
        self.C = clip.ARCHS[c["arch"]][4]
        q = qargs(cfg, max_imgs=2 if c["kind"] == "ft" else 10)      # configs/ftclip/*: 2 views per sample in training
        self.e2i = Event2Image(q, cfg["shape"], cfg["max_n"])
        self.T = self.e2i.max_imgs
        self.model = clip.init_weights_(clip.CLIP(c["arch"]), seed=0).to(dev).eval()
        vis = self.model.visual
        # calibration: features of a labelled batch (disjoint seeds, same classes) through ec_event2img + the encoder
        from eventclip_b200.synth import synth_labeled_batch
        ncal = {"n_cars": 64, "n_caltech101": 101, "n_imagenet": 16}[c["ds"]]
        evc, offc, labc = synth_labeled_batch(c["ds"], ncal, 3000000)
        nc = len(offc) - 1
        selc = draw_sel(self.e2i, offc, 99)
        frames, valid, _, nv = ops.plan_frames(offc, self.e2i.N, self.T, sel=selc, compact=True)
        with torch.no_grad():
            patches, _, _ = ops.event2img(torch.from_numpy(evc).to(dev), frames.to(dev), cfg["shape"], nv, cfg["count_non_zero"],
                                          cfg["background_mask"], out=vis.patch_fmt, patch=vis.patch_size, ldk=vis.patch_ldk)
            feats = vis.forward_patches(patches, nv).float().cpu()
        first = torch.from_numpy(np.concatenate([[0], np.cumsum(valid.sum(1).numpy())[:-1]]))     # first view of every sample
        self.text = calibrate_text_feats(feats[first], cfg["n_cls"])
        del patches, feats
        cd = dict(clip_model=self.model, prompt="a point cloud image of a {}", class_names=None, agg_func="mean",
                  text_feats=self.text)
        with contextlib.redirect_stdout(sys.stderr):                 # the classifiers print like the reference's do; stdout carries
            if c["kind"] == "zs":                                    # the one JSON line only
                m = ZSCLIPClassifier(clip_dict=cd)
            elif c["kind"] == "fs":
                ad = dict(adapter_type="text-trans", in_dim=self.C, d_model=256, num_heads=4, ffn_dim=1024, norm_first=True,
                          num_layers=2, residual=0.95)
                torch.manual_seed(3)
                m = FSCLIPClassifier(adapter_dict=ad, clip_dict=cd, loss_dict=dict(use_logits_loss=True, use_probs_loss=False))
            else:
                cd.update(lora="qkvo-16", only_conv1=False, only_bias=False, only_ln=False)
                torch.manual_seed(0)                                 # LoRA lora_down draws from the global RNG
                m = FTCLIPClassifier(adapter_dict=dict(adapter_type="text-identity", residual=True), clip_dict=cd,
                                     loss_dict=dict(use_logits_loss=True, use_probs_loss=False))
        self.cls = m.to(dev)
        self.cls = self.cls.train() if c["kind"] == "ft" else self.cls.eval()
        self.cls.attach_event_frontend(q, cfg["shape"], cfg["max_n"])
        self.host, self.devb, self.sels, self.labels = [], [], [], []
        for i in range(n_batches):
            ev, off, lab = make_batch(cname, 10000 * (rank + 1) + 1000 * i, uniq=uniq)
            he = torch.from_numpy(ev).pin_memory()
            self.host.append((he, torch.from_numpy(off)))
            self.devb.append((he.to(dev), torch.from_numpy(off)))
            self.sels.append(torch.from_numpy(draw_sel(self.e2i, off, 7 + i)))
            self.labels.append(lab)
        self.max_events = max(h[0].shape[0] for h in self.host)

    def data(self, i, resident=True):
        ev, off = (self.devb if resident else self.host)[i % len(self.host)]
        return dict(events=ev, event_offsets=off, sel_idx=self.sels[i % len(self.host)])

    def oracle(self, i, n):
        """Oracle port on the first n samples of batch i with this workload's text features / adapter weights."""
        from oracle import clip_oracle
        ev, off = self.host[i]
        ev, off = ev.numpy(), off.numpy()
        oc = clip_oracle.build_clip(self.c["arch"], seed=0)         # same seeded init as clip.init_weights_
        text = self.text
        ap = None
        if self.c["kind"] == "fs":
            sd = self.cls.state_dict()
            ap = {k[len("adapter."):]: v.detach().cpu() for k, v in sd.items() if k.startswith("adapter.")}
            text = sd["text_feats"].detach().cpu()
        sel = self.sels[i].numpy()[:n]
        return oracle_forward(self.cname, oc, ev[:off[n]], off[:n + 1], self.T, sel, text, ap)


def timed_steps(fn, K, world, dev, barrier):
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(K):
        fn(i)
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def other_inference_config(cname, dev, rank, world, K, barrier, pk, parity_n):
    """C1 / C3 / C4: a few graph-replayed steps (resident inputs, CUDA events, max over ranks), the e2e serving loop from
    pinned host memory, and the parity of the replayed path against the oracle on `parity_n` samples (rank 0)."""
    from eventclip_b200 import clip
    from eventclip_b200.graph import GraphedClassifier
    w = Workload(cname, dev, rank, n_batches=2, uniq=8 if CONFIGS[cname]["ds"] == "n_imagenet" else None)
    g = GraphedClassifier(w.cls, max_events=w.max_events)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(i):
        flush.zero_()
        with torch.no_grad():
            return g(w.data(i))

    for i in range(3):
        out = step(i)
    nv = int(out["valid_masks"].sum())
    ms = timed_steps(step, K, world, dev, barrier) / K

    def host_batches(n):
        for i in range(n):
            yield w.data(i, resident=False)

    with torch.no_grad():
        list(g.stream(host_batches(2), pre=flush.zero_))
        barrier()
        t0 = time.perf_counter()
        list(g.stream(host_batches(K), pre=flush.zero_))
        torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # the same serving loop fed with the compact wire format (4 bytes per event across PCIe instead of 16).
    # The host batches are packed before the timed loop (datasets.formats.pack_events_host: the role of the reference's
    # DataLoader workers, which parse the event files); logits equal those of the float route bit for bit (tests/test_formats.py).
    e2e_c = None
    try:
        if cname == "C4":
            raise StopIteration
        from eventclip_b200.datasets.formats import pack_events_host
        gc = GraphedClassifier(w.cls, max_events=w.max_events, compact=True)
        packed = []
        for i in range(2):
            d = w.data(i, resident=False)
            words = torch.from_numpy(pack_events_host(d["events"].numpy(), w.cfg["shape"]).view(np.int32)).pin_memory()
            packed.append(dict(d, events=words))

        def packed_batches(n):
            for i in range(n):
                yield packed[i % 2]

        with torch.no_grad():
            list(gc.stream(packed_batches(2), pre=flush.zero_))
            barrier()
            t0 = time.perf_counter()
            list(gc.stream(packed_batches(K), pre=flush.zero_))
            torch.cuda.synchronize()
        dtc = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(dtc, op=dist.ReduceOp.MAX)
        e2e_c = dict(value=world * w.B * K / float(dtc.item()), unit="samples/s", h2d_bytes_per_step=int(gc.h2d_bytes // K + nv * 16),
                     note="events cross PCIe as 4-byte words (flat pixel index | polarity), packed on the host before the timed loop")
        del gc
    except StopIteration:
        e2e_c = dict(skipped="encoder-bound: the float32 uploads already hide behind the 21 ms step")
    except Exception as ex:
        e2e_c = dict(error=repr(ex)[:200])
    fl = clip.flops_per_image(w.c["arch"]) * nv
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    r = dict(config=cname, workload=w.c["name"], n_gpus=world, per_gpu_batch=w.B, valid_views_per_step=nv, steps=K,
             ms_per_step=ms, samples_per_s=world * w.B / ms * 1e3, views_per_s=world * nv / ms * 1e3,
             encoder_tflops=fl / ms / 1e9, encoder_frac_of_peak=fl / ms / 1e9 / peak,
             e2e=dict(value=world * w.B * K / float(dt.item()), unit="samples/s",
                      h2d_bytes_per_step=int(g.h2d_bytes // K + nv * 16), d2h_bytes_per_step=w.B * 4 + 4,
                      events_in_host_batch_bytes=int(w.host[0][0].numel() * 4),
                      note="only the event ranges the frame plan reads are uploaded (2 of 14 chunks per N-ImageNet sample)"),
             e2e_compact_wire=e2e_c)
    if rank == 0 and parity_n > 0:
        try:
            with torch.no_grad():
                got = g(w.data(0))
                logits = got["logits"][:parity_n].float().cpu()
                patches = w.cls._last_patches
                ref = w.oracle(0, parity_n)
            st = parity_stats(logits, ref["logits"], k=5 if w.cfg["n_cls"] > 5 else 1)
            nvp = int(ref["valid"].sum())
            st["patch_rows_bit_exact"] = patches_bit_exact(patches, ref["imgs"], ref["valid"], w.model.visual.patch_size) \
                if nvp * (224 // w.model.visual.patch_size) ** 2 <= patches.shape[0] else None
            r["parity"] = st
        except Exception as ex:
            r["parity"] = dict(error=repr(ex)[:300])
    del g, w, flush
    torch.cuda.empty_cache()
    return r


def finetune_config(dev, rank, world, K, barrier, pk):
    """C5: the LoRA fine-tune step (events -> frames -> forward -> loss -> backward -> gradient all-reduce -> Adam -> weight
    refresh), CUDA-graph replayed; the communication exposed on the step is measured by timing the same steps with the
    all-reduce left out."""
    from eventclip_b200 import clip, train
    from eventclip_b200.graph import GraphedFineTuner
    import torch.distributed as dist
    w = Workload("C5", dev, rank, n_batches=1, uniq=8)
    tuner = train.FineTuner(w.cls, lr=2e-5)
    stepper = GraphedFineTuner(tuner, max_events=w.max_events)
    evd, off = w.devb[0]
    labels = torch.from_numpy(w.labels[0]).to(dev)
    sel = w.sels[0].numpy()
    first_loss = None
    for i in range(3):
        loss = stepper.step(evd, off, labels, sel=sel)
        if i == 0:
            first_loss = float(loss)
    ms = timed_steps(lambda i: stepper.step(evd, off, labels, sel=sel), K, world, dev, barrier) / K
    exposed = None
    if world > 1:
        real = tuner.allreduce
        tuner.allreduce = lambda: None
        ms_nocomm = timed_steps(lambda i: stepper.step(evd, off, labels, sel=sel), K, world, dev, barrier) / K
        tuner.allreduce = real
        tuner.flat.broadcast(tuner.pg)                           # the un-reduced steps let the ranks drift: re-sync
        exposed = ms - ms_nocomm
    nv = w.B * w.T
    chk = torch.stack([tuner.flat_p.double().sum(), tuner.flat_p.double().abs().sum()])
    same = True
    if world > 1:
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(torch.equal(c, allc[0]) for c in allc)
    peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
    fl = 3 * clip.flops_per_image("ViT-B/16") * nv
    r = dict(config="C5", workload=w.c["name"], n_gpus=world, per_gpu_batch=w.B, views_per_sample=w.T, steps=K, ms_per_step=ms,
             samples_per_s=world * w.B / ms * 1e3, fwd_bwd_tflops=fl / ms / 1e9, fwd_bwd_frac_of_peak=fl / ms / 1e9 / peak,
             trainable_params=tuner.flat.numel, allreduce_bytes=tuner.flat.numel * 4,
             exposed_comm_ms=exposed, params_identical_across_ranks=same, loss_first_step=first_loss)
    if rank == 0:
        try:       # parity of the step's loss: the oracle's fp32 forward on the same batch (LoRA up = 0 at the first step)
            from oracle import heads_oracle
            ref = w.oracle(0, w.B)
            ce = torch.nn.functional.cross_entropy(ref["logits"], torch.from_numpy(w.labels[0]))
            r["parity"] = dict(loss_first_step_oracle=float(ce), loss_abs_err=abs(float(ce) - first_loss))
        except Exception as ex:
            r["parity"] = dict(error=repr(ex)[:300])
    del stepper, tuner, w
    torch.cuda.empty_cache()
    return r


def b200_arm(args):
    import torch.distributed as dist
    from eventclip_b200 import clip, ops, _lib
    from eventclip_b200.graph import GraphedClassifier

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    K, Wm = args.steps, max(args.warmup, 3)
    NB = 3
    w = Workload(HEAD, dev, rank, n_batches=NB)
    cfg, B, T, zs, model = w.cfg, w.B, w.T, w.cls, w.model
    arch = w.c["arch"]
    host, devb = w.host, w.devb
    # labels: the ORACLE's predictions (fp32 CPU restatement of the reference path) for batch 0 of the timed region, computed
    # after the timing; -1 = no label (steps on the other batches do not count)
    labels = [torch.full((B,), -1, dtype=torch.int32, device=dev) for _ in range(NB)]
    counters = torch.zeros(2, dtype=torch.int64, device=dev)     # {n labelled, top-1 hits}: the AverageMeter state of test.py:67
    preds0 = torch.zeros(B, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    runner = zs if args.no_graph else GraphedClassifier(zs, max_events=w.max_events)
    # In-kernel time stamps of every GEMM launch (ec_gemm_timing): the stamp slots are baked into the captured graph, so
    # the launches are timed INSIDE the replayed, timed steps.  The shapes are logged in launch order while the warm-up /
    # capture run goes through Python.
    import ctypes
    import eventclip_b200.clip as clipmod
    STAMP_CAP = 4096
    stamps = torch.zeros(2 * STAMP_CAP, dtype=torch.int64, device=dev)
    gemm_log = []
    _orig_gemm = ops.gemm_bf16

    def logged_gemm(A, Wt, bias=None, epi="bf16", out=None, res=None, row_map=0, M=None):
        m = A.shape[0] if M is None else M
        gemm_log.append((m, Wt.shape[0], Wt.shape[1], epi))
        return _orig_gemm(A, Wt, bias, epi, out, res, row_map, M)

    _orig_ln, _orig_stats = ops.gemm_ln, ops.gemm_bf16_stats

    def logged_gemm_ln(x, Wg, colsum, cbias, st, n_parts, epi="bf16", out=None):
        gemm_log.append((x.shape[0], Wg.shape[0], Wg.shape[1], "ln_" + epi))
        return _orig_ln(x, Wg, colsum, cbias, st, n_parts, epi, out)

    def logged_gemm_stats(A, W, bias, x, st):
        gemm_log.append((A.shape[0], W.shape[0], W.shape[1], "f16_resadd_stats"))
        return _orig_stats(A, W, bias, x, st)

    _orig_stats2_log = ops.gemm_bf16_stats2

    def logged_gemm_stats2(A, W, bias, x2, st):
        gemm_log.append((A.shape[0], W.shape[0], W.shape[1], "f16x2_resadd_stats"))
        return _orig_stats2_log(A, W, bias, x2, st)

    if not args.no_graph:
        _lib.load().ec_gemm_timing(ctypes.c_void_p(stamps.data_ptr()), STAMP_CAP)
        clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = logged_gemm, logged_gemm_ln, logged_gemm_stats
        clipmod.ops.gemm_bf16_stats2 = logged_gemm_stats2

    def step(i, resident=True):
        flush.zero_()                                    # L2 flush between iterations (inside the timed region)
        with torch.no_grad():
            out = runner(w.data(i, resident))
        pred = out["top5_logits"][:, 0]
        lab = labels[i % NB]
        counters[0] += (lab >= 0).sum()
        counters[1] += (pred == lab).sum()
        return pred

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        step(i)
        if i == 0 and not args.no_graph:     # the graph exists now (2 eager warm-up passes + 1 capture went through Python)
            clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = _orig_gemm, _orig_ln, _orig_stats
            clipmod.ops.gemm_bf16_stats2 = _orig_stats2_log
            _lib.load().ec_gemm_timing(None, 0)          # later launches are not stamped; the graph keeps its slots
    # ---- parity of the path that is about to be timed (rank 0): graph replay of batch 0 vs the oracle on the whole batch.
    #      Its predictions become the labels of batch 0, so the accuracy counters of the timed region mean something. ----
    parity = None
    if world > 1:
        args.parity_samples = min(args.parity_samples, 64)       # the other ranks wait at the barrier meanwhile
    # torchrun exports OMP_NUM_THREADS=1: the oracle (rank 0 only, the other ranks idle at a barrier) takes the host's cores back
    torch.set_num_threads(max(1, (os.cpu_count() or 1) // (1 if world == 1 else 2)))
    if rank == 0 and not args.quick:
        try:
            with torch.no_grad():
                got = runner(w.data(0))
                logits0 = got["logits"].float().cpu()
                patches = zs._last_patches
            t0 = time.perf_counter()
            ref = w.oracle(0, args.parity_samples)
            parity = parity_stats(logits0[:args.parity_samples], ref["logits"])
            parity["patch_rows_bit_exact"] = patches_bit_exact(patches, ref["imgs"], ref["valid"], model.visual.patch_size)
            parity["oracle_seconds"] = time.perf_counter() - t0
            parity["path"] = ("eager launches" if args.no_graph else "CUDA-graph replay (the timed path)") + \
                ", batch 0 of the timed region, first %d samples; oracle = C event2img + fp32 PyTorch CLIP + head" % args.parity_samples
            lab0 = torch.full((B,), -1, dtype=torch.int32)
            lab0[:args.parity_samples] = ref["logits"].argmax(-1).to(torch.int32)
            labels[0].copy_(lab0)
            # the same batch through the other numeric settings of the inference forward (eager launches, not timed): what the
            # 16-bit operand format and the residual stream's dtype each cost in error
            vis = model.visual
            keep = (vis.operand_dtype, vis.residual_dtype, vis.residual_split)
            variants = {}
            for name, op, res, split in (("bf16_operands_fp16_residual", torch.bfloat16, torch.float16, False),
                                         ("fp16_operands_fp16_residual", torch.float16, torch.float16, False),
                                         ("fp16_operands_fp16x2_residual", torch.float16, torch.float16, True),
                                         ("fp16_operands_fp32_residual", torch.float16, torch.float32, False)):
                vis.operand_dtype, vis.residual_dtype, vis.residual_split = op, res, split
                with torch.no_grad():
                    lg = zs(w.data(0))["logits"].float().cpu()
                st = parity_stats(lg[:args.parity_samples], ref["logits"])
                variants[name] = {k: st[k] for k in ("logits_centered_rel_l2", "max_abs_logit_err", "top1_agree", "top1_disagreements")}
            vis.operand_dtype, vis.residual_dtype, vis.residual_split = keep
            parity["numeric_variants"] = variants
        except Exception as ex:
            parity = dict(error=repr(ex)[:300])
    counters.zero_()
    # ---- device-timed region: K steps, inputs resident ----
    barrier()
    l0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    clk = ClockSampler(local)
    clk.__enter__()
    for i in range(2):                                   # nvidia-smi is up and sampling before the timed steps start
        step(i)
    barrier()
    counters.zero_()
    l0 = _lib.LAUNCHES
    e0.record()
    for i in range(K):
        step(i)
    if world > 1:
        dist.all_reduce(counters)                        # the only collective of the inference path
    e1.record()
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    launches = _lib.LAUNCHES - l0
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * K / (ms_total / 1e3)
    if hasattr(runner, "check_status"):
        runner.check_status()                            # bad event coordinates would raise here (ValueError, as numpy does)
    in_step = None
    if not args.no_graph and gemm_log and len(gemm_log) % 3 == 0:
        n_g = len(gemm_log) // 3                         # launches per step; the captured pass is the last third
        st = stamps[:2 * len(gemm_log)].cpu().numpy().reshape(-1, 2)[2 * n_g:]
        dur_us = (st[:, 1] - st[:, 0]) / 1e3             # last replay of the timed region
        if (dur_us > 0).all():
            in_step = dict(keys=gemm_log[2 * n_g:], dur_us=dur_us)

    if args.quick:
        clk.__exit__()
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "ms_per_step": ms_total / K, "gpu_launches": launches, "quick": True}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end-to-end: pinned host buffers; every step's H2D copy and the D2H read of its predictions are inside ----
    def host_batches(n):
        for i in range(n):
            yield w.data(i, resident=False)

    def e2e_loop(n):
        hits = 0
        if args.no_graph:
            for i in range(n):
                hits += int((step(i, resident=False).cpu() == 0).sum())     # D2H of the step's predictions (synchronises)
        else:
            # the library's serving loop (GraphedClassifier.stream): batch i+1 uploads while batch i computes
            for pred in runner.stream(host_batches(n), pre=flush.zero_):
                hits += int((pred == 0).sum())
        return hits

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    e2e_loop(K)
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], device=dev)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    e2e = world * B * K / float(dt.item())
    clk.__exit__()
    h2d = host[0][0].numel() * 4 + B * T * 16
    d2h = B * 4 + 4

    line = None
    roofline = e2i = cb = None
    if rank == 0:
        # ---- roofline of the dominant kernel: the tcgen05 GEMM, CUDA events around each of its launches ----
        rec = []
        orig = ops.gemm_bf16

        def timed_gemm(A, Wt, bias=None, epi="bf16", out=None, res=None, row_map=0, M=None):
            m = A.shape[0] if M is None else M
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = orig(A, Wt, bias, epi, out, res, row_map, M)
            b.record()
            rec.append((a, b, 2.0 * m * Wt.shape[0] * Wt.shape[1], (m, Wt.shape[0], Wt.shape[1], epi)))
            return r

        def timed_ln(x, Wg, colsum, cbias, st, n_parts, epi="bf16", out=None):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = _orig_ln(x, Wg, colsum, cbias, st, n_parts, epi, out)
            b.record()
            rec.append((a, b, 2.0 * x.shape[0] * Wg.shape[0] * Wg.shape[1], (x.shape[0], Wg.shape[0], Wg.shape[1], "ln_" + epi)))
            return r

        def timed_stats(A, W, bias, x, st):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = _orig_stats(A, W, bias, x, st)
            b.record()
            rec.append((a, b, 2.0 * A.shape[0] * W.shape[0] * W.shape[1], (A.shape[0], W.shape[0], W.shape[1], "f16_resadd_stats")))
            return r

        _orig_stats2 = ops.gemm_bf16_stats2

        def timed_stats2(A, W, bias, x2, st):       # EC_RESIDUAL=fp16x2: the (hi, lo) residual update
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            r = _orig_stats2(A, W, bias, x2, st)
            b.record()
            rec.append((a, b, 2.0 * A.shape[0] * W.shape[0] * W.shape[1], (A.shape[0], W.shape[0], W.shape[1], "f16x2_resadd_stats")))
            return r

        clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = timed_gemm, timed_ln, timed_stats
        clipmod.ops.gemm_bf16_stats2 = timed_stats2
        nrep = 3
        for i in range(nrep):       # eager launches here: events cannot be recorded around nodes of a replayed graph
            flush.zero_()
            with torch.no_grad():
                zs(w.data(i))
        torch.cuda.synchronize()
        clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = orig, _orig_ln, _orig_stats
        clipmod.ops.gemm_bf16_stats2 = _orig_stats2
        gemm_ms = sum(a.elapsed_time(b) for a, b, _, _ in rec)
        gemm_flops = sum(f for _, _, f, _ in rec)
        shapes = {}
        for a, b, f, key in rec:
            t = shapes.setdefault("M%d_N%d_K%d_%s" % key, [0, 0.0, 0.0])
            t[0] += 1
            t[1] += a.elapsed_time(b)
            t[2] += f
        by_shape = {k: dict(launches=v[0] // nrep, us_per_launch=1e3 * v[1] / v[0], tflops=v[2] / (v[1] / 1e3) / 1e12)
                    for k, v in shapes.items()}
        n_gemm = len(rec) // nrep
        achieved = gemm_flops / (gemm_ms / 1e3) / 1e12
        isolated = dict(achieved=achieved, gemm_ms_per_step=gemm_ms / nrep, by_shape=by_shape,
                        note="eager launches with CUDA events around each GEMM: gaps between launches let the clocks recover")
        timing = "CUDA events around eager launches (no graph)"
        if in_step is not None:
            # the headline figure: the same launches timed inside the replayed steps of the timed region (%globaltimer stamps)
            fl = np.array([2.0 * m * n * k for m, n, k, _ in in_step["keys"]])
            achieved = float(fl.sum() / (in_step["dur_us"].sum() * 1e-6) / 1e12)
            gemm_ms, gemm_flops, nrep, n_gemm = float(in_step["dur_us"].sum()) / 1e3, float(fl.sum()), 1, len(fl)
            shapes = {}
            for key, f, du in zip(in_step["keys"], fl, in_step["dur_us"]):
                t = shapes.setdefault("M%d_N%d_K%d_%s" % key, [0, 0.0, 0.0])
                t[0] += 1
                t[1] += du / 1e3
                t[2] += f
            by_shape = {k: dict(launches=v[0], us_per_launch=1e3 * v[1] / v[0], tflops=v[2] / (v[1] / 1e3) / 1e12)
                        for k, v in shapes.items()}
            timing = "in-kernel %globaltimer stamps of the GEMM nodes of the replayed graph, last step of the timed region"
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        traffic, traffic_src = None, None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            traffic_src = "profiles/gemm_traffic.json (one ncu --set full capture of this kernel; not measured in this run)"
        roofline = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                        traffic_source=traffic_src,
                        kernel="gemm_kernel<BN> (tcgen05.mma kind::f16, TMA, TMEM)", launches_per_step=n_gemm,
                        flops_per_step=gemm_flops / nrep, gemm_ms_per_step=gemm_ms / nrep, peak_source=pk["_source"],
                        peak_kind="sustained cuBLAS bf16", by_shape=by_shape, timing=timing,
                        share_of_step=(gemm_ms / nrep) / (ms_total / K), isolated_launches=isolated)
        e2i = event2img_metric(dev, pk)
        # ---- CPU baseline: oracle port on this host, bounded sample ----
        cpu_value, cpu_ms, cores = run_cpu(32, 2, 1) if world == 1 else (None, None, os.cpu_count())
        cb = dict(value=cpu_value, unit="samples/s", cores=cores, kind="port",
                  sample="2 steps x 32 samples of the bench workload through the oracle (C event2img + fp32 PyTorch CLIP + head)")
    enc_flops = clip.flops_per_image(arch) * B * T
    residual = ("fp16 kept as a (hi, lo) pair of fp16 planes: hi = fp16(x) feeds the GEMMs, lo carries the rounding residue through every update"
                if model.visual.residual_split and model.visual.residual_dtype == torch.float16 else
                "fp16 (the reference's CUDA precision)" if model.visual.residual_dtype == torch.float16 else "fp32")
    operands = "fp16" if model.visual.operand_dtype == torch.float16 else "bf16"
    acc = counters.tolist()
    del runner, w, zs, model, host, devb
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs (collective: every rank takes part; C4 and C5 are the ones BASELINE.json scales) ----
    others = {}
    if not args.no_others:
        Ko = max(2, min(K, args.other_steps))
        names = ["C1", "C3", "C4", "C5"] if world == 1 else ["C4", "C5"]
        for cname in names:
            try:
                if cname == "C5":
                    others[cname] = finetune_config(dev, rank, world, Ko, barrier, pk)
                else:
                    others[cname] = other_inference_config(cname, dev, rank, world, Ko, barrier, pk,
                                                           parity_n=(8 if cname != "C4" else 4) if world == 1 else 0)
            except Exception as ex:
                if world > 1:
                    raise                                    # a rank that drops out would hang the others at the next barrier
                others[cname] = dict(error=repr(ex)[:300])
                torch.cuda.empty_cache()
    if rank == 0:
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": operands,
            "data": "synthetic",
            "config": {"workload": f"zero-shot {arch} on synthetic N-Cars-shaped streams (120x100, 4000 events/sample, "
                                   f"2 classes), batch {B} per GPU, random-init CLIP (BASELINE.json configs[1])",
                       "streams": "labelled synthetic samples with per-sample blob layouts (synth.synth_labeled_batch); text "
                                  "features calibrated on a disjoint batch so the predictions split evenly",
                       "per_gpu_batch": B, "views_per_sample": T,
                       "numerics": operands + " tensor-core operands (tcgen05 kind::f16; EC_OPERANDS selects bf16 / fp16, same rate), fp32 "
                                   "accumulation, residual stream in " + residual,
                       "l2": "256 MiB buffer rewritten before every step (inside the timed region)",
                       "sharding": "samples by rank; one all-reduce of 2 int64 counters at the end",
                       "launch": "eager" if args.no_graph else "CUDA graph replay of the device part (event2img..head)"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "eager classifier call per step" if args.no_graph else
                           "GraphedClassifier.stream(): pinned host events, copy stream one batch ahead, predictions + status word read back every step"},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "parity": parity,
            "roofline": roofline,
            "encoder": {"algorithmic_tflops_per_step": enc_flops / 1e12,
                        "whole_step_tflops": enc_flops / (ms_total / K / 1e3) / 1e12,
                        "whole_step_frac_of_peak": enc_flops / (ms_total / K / 1e3) / 1e12 / peak},
            "event2img": e2i,
            "cpu_baseline": cb,
            "accuracy_counters": {"n_labelled": acc[0], "top1_hits": acc[1],
                                  "labels": "oracle predictions for batch 0 of the timed region (steps on the other batches carry no label)"},
            "other_configs": others,
        }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--quick", action="store_true", help="timed region only (for ncu launch lists)")
    ap.add_argument("--e2i-only", action="store_true", help="only the event2img Gevents/s section")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of CUDA-graph replay")
    ap.add_argument("--no-others", action="store_true", help="skip the other_configs block (C1, C3, C4, C5)")
    ap.add_argument("--other-steps", type=int, default=5, help="timed steps per other config")
    ap.add_argument("--parity-samples", type=int, default=256, help="samples of timed batch 0 the oracle checks")
    args = ap.parse_args()
    if args.e2i_only:
        torch.cuda.set_device(0)
        print(json.dumps(event2img_metric(torch.device("cuda", 0), peaks(), cpu=False)))
    elif args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
