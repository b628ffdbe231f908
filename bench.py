#!/usr/bin/env python
"""bench.py -- events -> logits throughput of the EventCLIP hot path on B200 (and the reference CPU arm).

    python bench.py [--gpus N --steps K --warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1], the configuration the metric is quoted on): zero-shot EventCLIP ViT-B/16 on
synthetic N-Cars-shaped streams (120x100 sensor, 4000 events/sample, 2 classes), batch 256 per GPU, bf16 tensor-core
math with fp32 accumulation, random-init CLIP weights (no checkpoints offline).  One step = one batch through
ec_event2img -> patch GEMM -> 12 ViT blocks -> proj -> head.  Rank r works on its own batch (samples shard by rank,
weak scaling); NCCL carries only the final prediction-counter all-reduce.

One JSON line on stdout (rank 0):
  value      samples/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public classifier API with HOST (pinned) buffers: H2D of the packed events and
             D2H of the predictions inside the timed region
  roofline   dominant kernel = the tcgen05 GEMM: algorithmic FLOPs of its launches / their summed CUDA-event time,
             against the measured cuBLAS bf16 peak of MEASURED_PEAKS.json
  cpu_baseline  the oracle port (C event2img + fp32 PyTorch CLIP + head) on this host's cores, bounded sample
  event2img  secondary metric of BASELINE.json: Gevents/s of the fused kernel alone vs the HBM roofline
"""
import argparse
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

ARCH = "ViT-B/16"
DATASET = "n_cars"
BATCH = 256
METRIC = "event samples/s (events->logits)"
FALLBACK_PEAKS = dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0)   # B200_PROFILING.md fallback


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
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
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


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_events_to_logits(ev, off, cfg, oracle_clip, text, T):
    """Oracle port of the whole path for one packed batch: C event2img + fp32 PyTorch CLIP + head."""
    from oracle import event2img as orc
    from oracle import heads_oracle
    B = len(off) - 1
    imgs, valids = [], []
    for b in range(B):
        im, va, _ = orc.event2img_sample(ev[off[b]:off[b + 1]], cfg["shape"], cfg["N"], T, cfg["count_non_zero"],
                                         cfg["background_mask"])
        imgs.append(im)
        valids.append(va)
    imgs, valid = torch.from_numpy(np.stack(imgs)), torch.from_numpy(np.stack(valids))
    with torch.no_grad():
        feats = oracle_clip.encode_image(imgs[valid])
    return heads_oracle.zs_head(feats, valid, text, 100.0, "mean")


def run_cpu(samples_per_step, steps, warmup):
    from eventclip_b200.synth import SENSORS, synth_batch
    from oracle import clip_oracle
    cfg = SENSORS[DATASET]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    oracle_clip = clip_oracle.build_clip(ARCH, seed=0)
    text = clip_oracle.synth_text_feats(cfg["n_cls"], 512, 1)
    ev, off = synth_batch(DATASET, samples_per_step, 9000)
    for _ in range(warmup):
        cpu_events_to_logits(ev, off, cfg, oracle_clip, text, 1)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_events_to_logits(ev, off, cfg, oracle_clip, text, 1)
    dt = time.perf_counter() - t0
    return samples_per_step * steps / dt, dt / steps * 1e3, cores


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sps = 32
    steps, warmup = max(1, min(args.steps, 8)), max(1, min(args.warmup, 2))
    value, ms, cores = run_cpu(sps, steps, warmup)
    cb = dict(value=value, unit="samples/s", cores=cores, kind="port",
              sample=f"{sps} samples/step x {steps} steps of the bench workload (oracle: C event2img + fp32 PyTorch CLIP {ARCH} + head)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"zero-shot {ARCH}, N-Cars-shaped streams 120x100, 4000 events/sample, 2 classes, "
                               f"{sps} samples per CPU step (bounded sample of the batch-256 workload)"},
        "cpu_baseline": cb,
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ------------------------------------------------------------------------------------------------ B200 arm
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


def event2img_metric(dev, pk):
    """BASELINE.json's second metric: Gevents/s of the fused kernel alone (bf16 patch rows out), HBM roofline."""
    from eventclip_b200 import ops
    from eventclip_b200.datasets import Event2Image
    from eventclip_b200.synth import SENSORS, synth_batch
    out = {}
    # batch sizes give whole waves of clusters on 148 SMs (1480 / 2072 / 296 frames) and inputs far beyond L2
    for ds, B, reps in (("n_caltech101", 296, 5), ("n_cars", 2072, 5), ("n_imagenet", 144, 3)):
        cfg = SENSORS[ds]
        e2i = Event2Image(qargs(cfg), cfg["shape"], cfg["max_n"])
        ev1, off1 = synth_batch(ds, 8, 100)
        evs = np.concatenate([ev1] * (B // 8))
        off = np.concatenate([[0], np.cumsum(np.tile(np.diff(off1), B // 8))]).astype(np.int64)
        evd = torch.from_numpy(evs).to(dev)
        T = e2i.max_imgs
        sel = np.tile(np.arange(T, dtype=np.int32), (B, 1))
        frames, valid, chunks, nv = ops.plan_frames(off, e2i.N, T, sel=sel, compact=True)
        fd = frames.to(dev)
        rec = np.frombuffer(frames.numpy().tobytes(), dtype=[("s", "<i8"), ("n", "<i4"), ("o", "<i4")])
        ev_read = int(rec["n"].sum())
        outbuf = torch.zeros((nv * 196, 768), dtype=torch.bfloat16, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        run = lambda: ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="patch",
                                    patch=16, ldk=768, out_tensor=outbuf, status=status)
        for _ in range(3):
            run()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            run()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / reps
        byts = 16 * ev_read + nv * 3 * 224 * 224 * 2
        # row F2: the same frames from the compact wire format (4 bytes per event); the metric's byte count changes with
        # it, so both conventions are reported: against the reference's 16-byte events and against the bytes really read
        words = ops.pack_events(evd, cfg["shape"])
        runc = lambda: ops.event2img(words, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="patch",
                                     patch=16, ldk=768, out_tensor=outbuf, status=status)
        for _ in range(3):
            runc()
        torch.cuda.synchronize()
        s.record()
        for _ in range(reps):
            runc()
        e.record()
        torch.cuda.synchronize()
        msc = s.elapsed_time(e) / reps
        bytc = 4 * ev_read + nv * 3 * 224 * 224 * 2
        compact = dict(ms=msc, gevents_per_s=ev_read / msc / 1e6, frac_reference_bytes=byts / msc / 1e6 / pk["hbm_gbs"],
                       frac_compact_bytes=bytc / msc / 1e6 / pk["hbm_gbs"], input_mb=words.numel() * 4 / 1e6)
        del words
        out[ds] = dict(compact_wire_format=compact, frames=int(nv), events_histogrammed=ev_read, events_in_streams=int(off[-1]), ms=ms,
                       gevents_per_s=ev_read / ms / 1e6, gevents_per_s_stream=int(off[-1]) / ms / 1e6,
                       algorithmic_bytes=byts, achieved_gbs=byts / ms / 1e6, frac=byts / ms / 1e6 / pk["hbm_gbs"],
                       input_mb=evs.nbytes / 1e6, geometry=ops.event2img_geometry(cfg["shape"]))
        del evd, outbuf
        try:
            out[ds]["cpu_baseline"] = event2img_cpu_mev(ds, {"n_caltech101": 64, "n_cars": 512, "n_imagenet": 16}[ds])
        except Exception as e:      # the oracle is test infrastructure: never let it take the bench line down
            out[ds]["cpu_baseline"] = dict(error=str(e)[:200])
    return out


def b200_arm(args):
    import torch.distributed as dist
    from eventclip_b200 import clip, ops, _lib
    from eventclip_b200.models import ZSCLIPClassifier
    from eventclip_b200.synth import SENSORS, synth_batch, synth_text_feats

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    cfg = SENSORS[DATASET]
    B, K, Wm = BATCH, args.steps, max(args.warmup, 3)

    model = clip.init_weights_(clip.CLIP(ARCH), seed=0).to(dev).eval()
    text = synth_text_feats(cfg["n_cls"], 512, 1)
    zs = ZSCLIPClassifier(clip_dict=dict(clip_model=model, prompt="a point cloud image of a {}", class_names=None,
                                         agg_func="mean", text_feats=text)).to(dev).eval()
    zs.attach_event_frontend(qargs(cfg), cfg["shape"], cfg["max_n"])
    T = zs.event_frontend.max_imgs
    NB = 3
    host, devb = [], []
    for i in range(NB):
        ev, off = synth_batch(DATASET, B, 10000 * (rank + 1) + 1000 * i)
        he = torch.from_numpy(ev).pin_memory()
        host.append((he, torch.from_numpy(off)))
        devb.append((he.to(dev), torch.from_numpy(off)))
    sel = torch.from_numpy(np.tile(np.arange(T, dtype=np.int32), (B, 1)))
    labels = torch.zeros(B, dtype=torch.int32, device=dev)
    counters = torch.zeros(2, dtype=torch.int64, device=dev)     # {n, top-1 hits}: the AverageMeter state of test.py:67
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > 126 MB L2

    from eventclip_b200.graph import GraphedClassifier
    runner = zs if args.no_graph else GraphedClassifier(zs, max_events=host[0][0].shape[0])
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

    if not args.no_graph:
        _lib.load().ec_gemm_timing(ctypes.c_void_p(stamps.data_ptr()), STAMP_CAP)
        clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = logged_gemm, logged_gemm_ln, logged_gemm_stats

    def step(i, resident=True):
        ev, off = (devb if resident else host)[i % NB]
        flush.zero_()                                    # L2 flush between iterations (inside the timed region)
        with torch.no_grad():
            out = runner(dict(events=ev, event_offsets=off, sel_idx=sel))
        pred = out["top5_logits"][:, 0]
        counters[0] += B
        counters[1] += (pred == labels).sum()
        return pred

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(Wm):
        step(i)
        if i == 0 and not args.no_graph:     # the graph exists now (2 eager warm-up passes + 1 capture went through Python)
            clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = _orig_gemm, _orig_ln, _orig_stats
            _lib.load().ec_gemm_timing(None, 0)          # later launches are not stamped; the graph keeps its slots
    # ---- device-timed region: K steps, inputs resident ----
    barrier()
    l0 = _lib.LAUNCHES
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        e0.record()
        for i in range(K):
            step(i)
        if world > 1:
            dist.all_reduce(counters)                    # the only collective of the inference path
        e1.record()
        barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    launches = _lib.LAUNCHES - l0
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * K / (ms_total / 1e3)
    in_step = None
    if not args.no_graph and gemm_log and len(gemm_log) % 3 == 0:
        n_g = len(gemm_log) // 3                         # launches per step; the captured pass is the last third
        st = stamps[:2 * len(gemm_log)].cpu().numpy().reshape(-1, 2)[2 * n_g:]
        dur_us = (st[:, 1] - st[:, 0]) / 1e3             # last replay of the timed region
        if (dur_us > 0).all():
            in_step = dict(keys=gemm_log[2 * n_g:], dur_us=dur_us)

    if args.quick:
        if rank == 0:
            print(json.dumps({"metric": METRIC, "value": value, "ms_per_step": ms_total / K, "gpu_launches": launches, "quick": True}))
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- end-to-end: pinned host buffers; every step's H2D copy and the D2H read of its predictions are inside ----
    def host_batches(n):
        for i in range(n):
            he, off = host[i % NB]
            yield dict(events=he, event_offsets=off, sel_idx=sel)

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
    h2d = host[0][0].numel() * 4 + B * T * 16
    d2h = B * 4

    line = None
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

        import eventclip_b200.clip as clipmod
        clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = timed_gemm, timed_ln, timed_stats
        nrep = 3
        for i in range(nrep):       # eager launches here: events cannot be recorded around nodes of a replayed graph
            flush.zero_()
            with torch.no_grad():
                zs(dict(events=devb[i % NB][0], event_offsets=devb[i % NB][1], sel_idx=sel))
        torch.cuda.synchronize()
        clipmod.ops.gemm_bf16, clipmod.ops.gemm_ln, clipmod.ops.gemm_bf16_stats = orig, _orig_ln, _orig_stats
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
        traffic = None
        tp = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_launch")
        roofline = dict(bound="tensor", achieved=achieved, peak=peak, unit="TFLOP/s", frac=achieved / peak, traffic=traffic,
                        kernel="gemm_kernel<BN> (tcgen05.mma kind::f16, TMA, TMEM)", launches_per_step=n_gemm,
                        flops_per_step=gemm_flops / nrep, gemm_ms_per_step=gemm_ms / nrep, peak_source=pk["_source"],
                        peak_kind="sustained cuBLAS bf16", by_shape=by_shape, timing=timing,
                        share_of_step=(gemm_ms / nrep) / (ms_total / K), isolated_launches=isolated)
        enc_flops = clip.flops_per_image(ARCH) * B * T
        e2i = event2img_metric(dev, pk)
        # ---- CPU baseline: oracle port on this host, bounded sample ----
        cpu_value, cpu_ms, cores = run_cpu(32, 2, 1) if world == 1 else (None, None, os.cpu_count())
        cb = dict(value=cpu_value, unit="samples/s", cores=cores, kind="port",
                  sample="2 steps x 32 samples of the bench workload through the oracle (C event2img + fp32 PyTorch CLIP + head)")
        line = {
            "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": K, "warmup": Wm,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"zero-shot {ARCH} on synthetic N-Cars-shaped streams (120x100, 4000 events/sample, "
                                   f"2 classes), batch {B} per GPU, random-init CLIP (BASELINE.json configs[1])",
                       "per_gpu_batch": B, "views_per_sample": T,
                       "numerics": "bf16 tensor-core operands, fp32 accumulation, residual stream in "
                                   + ("fp16 (the reference's CUDA precision)" if model.visual.residual_dtype == torch.float16 else "fp32"), "l2": "256 MiB buffer rewritten before every step (inside the timed region)",
                       "sharding": "samples by rank; one all-reduce of 2 int64 counters at the end",
                       "launch": "eager" if args.no_graph else "CUDA graph replay of the device part (event2img..head)"},
            "e2e": {"value": e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "eager classifier call per step" if args.no_graph else
                           "GraphedClassifier.stream(): pinned host events, copy stream one batch ahead, predictions read back every step"},
            "gpu_launches": launches,
            "clocks": clk.summary(),
            "roofline": roofline,
            "encoder": {"algorithmic_tflops_per_step": enc_flops / 1e12,
                        "whole_step_tflops": enc_flops / (ms_total / K / 1e3) / 1e12,
                        "whole_step_frac_of_peak": enc_flops / (ms_total / K / 1e3) / 1e12 / peak},
            "event2img": e2i,
            "cpu_baseline": cb,
            "accuracy_counters": counters.tolist(),
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
    args = ap.parse_args()
    if args.e2i_only:
        torch.cuda.set_device(0)
        print(json.dumps(event2img_metric(torch.device("cuda", 0), peaks())))
    elif args.impl == "reference":
        reference_arm(args)
    else:
        b200_arm(args)


if __name__ == "__main__":
    main()
