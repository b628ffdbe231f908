"""CUDA-graph replay of the fused events -> logits forward for fixed batch geometry.

The device part of the classifier forward (`device_forward`: ec_event2img -> patch GEMM -> ViT blocks -> head, ~100
kernel launches) is captured once and replayed, so a step costs one graph launch instead of ~100 ctypes calls.
A graph is specific to the COUNTS of a plan (frames, valid views, B, T); the frame table, the valid mask and the slot
maps are static device tensors refreshed before every replay, so batches with different chunk offsets or padding share
a graph (`GraphedClassifier` keeps a small cache keyed by the counts).  `GraphedFineTuner` does the same for the
fine-tune step: one graph for forward + loss + backward, the gradient all-reduce outside it, one graph for Adam + the
LoRA re-merge.
"""
import torch

from . import _lib as L


class _Entry:
    pass


def _static_plan(model, plan, dev):
    """Device copy of a host plan whose tensors are STATIC graph inputs: the slot map is always materialised so that two
    batches with the same counts (frames, valid views, B, T) share a graph whatever their chunk offsets / padding."""
    B, T, nv = plan["B"], plan["T"], plan["n_valid"]
    d = model.plan_to_device(plan, dev)
    if d["row_of_slot"] is None:
        d["row_of_slot"] = torch.arange(B * T, dtype=torch.int32, device=dev)
    d["slot_of_row"] = torch.zeros(nv, dtype=torch.int32, device=dev)
    _refresh_plan(d, plan)
    return d


def _refresh_plan(static, plan):
    """Copies a new host plan of the same counts into the static device tensors (async from pinned memory); nothing is
    copied when the plan equals the one already on the device (fixed-geometry streams)."""
    sig = (plan["frames"].numpy().tobytes(), plan["valid_u8"].numpy().tobytes())
    if static.get("_sig") == sig:
        return
    static["_sig"] = sig
    pin = lambda t: t if (t.is_pinned() or not torch.cuda.is_available()) else t.pin_memory()
    B, T = plan["B"], plan["T"]
    static["frames"].copy_(pin(plan["frames"]), non_blocking=True)
    static["valid_u8"].copy_(pin(plan["valid_u8"]), non_blocking=True)
    static["valid_dev"].copy_(pin(plan["valid"]), non_blocking=True)
    ros = plan["row_of_slot"]
    if ros is None:
        ros = torch.arange(B * T, dtype=torch.int32)
    static["row_of_slot"].copy_(pin(ros), non_blocking=True)
    static["slot_of_row"].copy_(pin((ros >= 0).nonzero().squeeze(1).to(torch.int32)), non_blocking=True)
    static["valid"] = plan["valid"]


def _plan_key(plan):
    return (plan["frames"].shape[0], plan["n_valid"], plan["B"], plan["T"])


_FRAME_DT = [("s", "<i8"), ("n", "<i4"), ("o", "<i4")]


def _used_ranges(plan, n_events, min_saving=0.25):
    """When the plan reads only part of a HOST event stream (N-ImageNet: 2 of 14 chunks per sample), the rest need not cross
    PCIe.  Returns (ranges, plan') with ranges = [(src_start, count, dst_start)] of the events the frames histogram, packed
    back to back, and plan' = the plan with its frame table re-based onto that packing; (None, plan) when nearly everything
    is used anyway.  Overlapping chunks (a tail chunk starting inside its predecessor, vis.py:55-72) are copied once each."""
    import numpy as np
    rec = np.frombuffer(plan["frames"].numpy().tobytes(), dtype=_FRAME_DT).copy()
    used = int(rec["n"].clip(min=0).sum())
    if used >= (1.0 - min_saving) * n_events:
        return None, plan
    ranges, cur = [], 0
    for i in range(rec.shape[0]):
        c = int(rec["n"][i])
        if c <= 0:
            continue
        ranges.append((int(rec["s"][i]), c, cur))
        rec["s"][i] = cur
        cur += c
    frames = torch.from_numpy(np.frombuffer(rec.tobytes(), dtype=np.uint8).reshape(-1, 16).copy())
    p2 = dict(plan)
    p2["frames"] = frames.pin_memory() if torch.cuda.is_available() else frames
    p2["_n_events"] = cur
    return ranges, p2


class GraphedClassifier:
    def __init__(self, model, max_events, max_graphs=8, compact=False):
        """compact=True: the event batches arrive in the compact wire format (int32 [sum E] words of
        datasets.formats.pack_events_host / ops.pack_events: 4 bytes per event across PCIe instead of 16)."""
        self.model = model
        self.dev = model.device
        if self.dev.type != "cuda":
            raise L.ECError("GraphedClassifier needs the model on a CUDA device")
        self.compact = bool(compact)
        self.ev_bytes = 4 if self.compact else 16
        # static input buffer
        self.events = torch.zeros((max_events,), dtype=torch.int32, device=self.dev) if self.compact else \
            torch.zeros((max_events, 4), dtype=torch.float32, device=self.dev)
        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.cache = {}
        self.max_graphs = max_graphs
        self._params = list(model.parameters())
        self._weights_sig = self._signature()
        self._nonempty = False                # the static event buffer holds a real batch (set by the first call)

    def _signature(self):
        """Changes whenever a parameter is updated in place or replaced (optimizer step, load_state_dict): the captured
        graphs read packed copies of the weights whose buffers are re-created then, so they must be captured again."""
        vis = getattr(getattr(self.model, "model", None), "visual", None)
        return (sum(p._version for p in self._params) + sum(p.data_ptr() for p in self._params[:1]),
                getattr(vis, "_epoch", 0))      # _epoch: in-place updates the version counters cannot see (train.FineTuner)

    def _build(self, plan):
        e = _Entry()
        e.plan = _static_plan(self.model, plan, self.dev)
        torch.cuda.synchronize(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side), torch.no_grad():      # warm-up outside capture: tables, packed weights, attributes
            for _ in range(2):
                self.model.device_forward(self.events, e.plan, status=self.status)
            vis = getattr(getattr(self.model, "model", None), "visual", None)
            if vis is not None and hasattr(vis, "validate_ln_fold") and not getattr(vis, "_fold_checked", False) \
                    and not getattr(self.model, "training", False) and self._nonempty:
                # first real batch: check the LayerNorm folding against the LayerNorm kernels on this tower's own activations
                vis.validate_ln_fold(self.model._last_patches, e.plan["n_valid"])
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        e.graph = torch.cuda.CUDAGraph()
        n0 = L.LAUNCHES
        with torch.cuda.graph(e.graph), torch.no_grad():
            e.out = self.model.device_forward(self.events, e.plan, status=self.status)
        e.n_launch = L.LAUNCHES - n0          # kernel nodes of the graph
        L.LAUNCHES = n0                       # capture enqueued nothing
        return e

    def __call__(self, data_dict):
        """data_dict: {'events' float32 [sum E,4] -- or int32 [sum E] packed words with compact=True -- (pinned host or CUDA),
        'event_offsets', optional 'sel_idx'}.
        Returns the classifier's out_dict; its tensors are the graph's static outputs (overwritten by the next call).
        Host events are copied range by range when the plan uses only part of the stream (see _used_ranges)."""
        plan = self.model.plan_events(data_dict["event_offsets"], data_dict.get("sel_idx", None))
        ev = data_dict["events"]
        ranges = None
        if not ev.is_cuda:
            ranges, plan = _used_ranges(plan, ev.shape[0])
        return self._run(plan, ev, ranges)

    def _copy_in(self, dst, ev, ranges):
        """events -> dst (device), whole or as the packed ranges of _used_ranges; returns the bytes moved."""
        if ev.dtype != dst.dtype:
            raise L.ECError(f"GraphedClassifier(compact={self.compact}) expects {dst.dtype} events, got {ev.dtype}")
        if ranges is None:
            dst[:ev.shape[0]].copy_(ev, non_blocking=True)
            return ev.shape[0] * self.ev_bytes
        for s0, c, d0 in ranges:
            dst[d0:d0 + c].copy_(ev[s0:s0 + c], non_blocking=True)
        return sum(c for _, c, _ in ranges) * self.ev_bytes

    def _run(self, plan, ev, ranges, staged=False):
        """Replay (or capture) the graph of this plan's geometry on `ev`: a host / device event tensor, or -- staged=True --
        a device tensor that already holds the packed ranges."""
        n = plan.get("_n_events", ev.shape[0])
        if n > self.events.shape[0]:
            raise L.ECError(f"batch has {n} events but the graph buffer holds {self.events.shape[0]}")
        sig = self._signature()
        if sig != self._weights_sig:          # weights changed since the graphs were captured
            torch.cuda.synchronize(self.dev)
            self.cache.clear()
            self._weights_sig = sig
        key = _plan_key(plan)
        ent = self.cache.get(key)
        fill = (lambda: self.events[:n].copy_(ev[:n], non_blocking=True)) if staged or ranges is None else \
            (lambda: self._copy_in(self.events, ev, ranges))
        if ent is None:
            if len(self.cache) >= self.max_graphs:
                self.cache.pop(next(iter(self.cache)))
            fill()                                # the warm-up passes of _build run on this batch
            self._nonempty = True
            ent = self.cache[key] = self._build(plan)
        else:
            _refresh_plan(ent.plan, plan)
        fill()
        ent.graph.replay()
        L.LAUNCHES += ent.n_launch
        return ent.out

    def check_status(self):
        """Synchronising read of the event kernel's status word (sticky across replays): raises ValueError for event
        coordinates outside the sensor, as the reference's numpy path does (datasets/vis.py:9-14)."""
        from . import ops
        ops.raise_on_status(self.status)

    def stream(self, batches, result=None, pre=None):
        """Pipelined serving loop over an iterable of data_dicts with PINNED HOST events -- the role the reference's
        DataLoader prefetch + `non_blocking` copies play around `model(data_dict)` (test.py:55-81).  While batch i is
        replayed, a copy stream uploads batch i+1 into one of two staging buffers; `result(out_dict)` (default: top-1
        of the aggregated logits) is read back through pinned memory and yielded as a host tensor, in order, once its
        batch has finished.  `pre()` runs on the compute stream before every replay (bench.py flushes L2 there)."""
        dev = self.dev
        if result is None:
            result = lambda out: out["top5_logits"][:, 0]
        if not hasattr(self, "_stage"):
            self._stage = [torch.empty_like(self.events) for _ in range(2)]
            self._copy = torch.cuda.Stream(device=dev)
            self._uploaded = [torch.cuda.Event() for _ in range(2)]
            self._consumed = [torch.cuda.Event() for _ in range(2)]
        cur = torch.cuda.current_stream(dev)
        done = [torch.cuda.Event() for _ in range(2)]
        host_res = [None, None]
        host_st = [torch.zeros(1, dtype=torch.int32).pin_memory() for _ in range(2)]   # status word rides with the result

        def finish(j):
            done[j].synchronize()
            if int(host_st[j][0]) != 0:
                from . import ops
                ops.raise_on_status(host_st[j])
            return host_res[j].clone()

        plans = [None, None]
        self.h2d_bytes = 0                                            # event bytes uploaded by the last stream() (all batches)

        def upload(i, d):
            ev = d["events"]
            plan = self.model.plan_events(d["event_offsets"], d.get("sel_idx", None))
            ranges = None
            if not ev.is_cuda:
                if not ev.is_pinned():
                    raise L.ECError("GraphedClassifier.stream: host events must be pinned (torch.Tensor.pin_memory())")
                ranges, plan = _used_ranges(plan, ev.shape[0])
                if ranges is None:
                    plan = dict(plan, _n_events=ev.shape[0])
                n = plan["_n_events"]
                if n > self.events.shape[0]:
                    raise L.ECError(f"batch has {n} events but the graph buffer holds {self.events.shape[0]}")
                with torch.cuda.stream(self._copy):
                    if i >= 2:
                        self._copy.wait_event(self._consumed[i % 2])      # batch i-2 has left this staging buffer
                    self.h2d_bytes += self._copy_in(self._stage[i % 2], ev, ranges)
                    self._uploaded[i % 2].record(self._copy)
            plans[i % 2] = plan

        def launch(i, d):
            ev = d["events"]
            plan = plans[i % 2]
            if not ev.is_cuda:
                cur.wait_event(self._uploaded[i % 2])
                ev = self._stage[i % 2]
            if pre is not None:
                pre()
            out = self._run(plan, ev, None, staged=True)              # device-to-device copy into the graph's input + replay
            self._consumed[i % 2].record(cur)
            r = result(out)
            if host_res[i % 2] is None or host_res[i % 2].shape != r.shape or host_res[i % 2].dtype != r.dtype:
                host_res[i % 2] = torch.empty(r.shape, dtype=r.dtype).pin_memory()
            host_res[i % 2].copy_(r, non_blocking=True)
            host_st[i % 2].copy_(self.status, non_blocking=True)
            done[i % 2].record(cur)

        it = iter(batches)
        nxt = next(it, None)
        if nxt is None:
            return
        upload(0, nxt)
        i = 0
        while nxt is not None:
            d, nxt = nxt, next(it, None)
            if nxt is not None:
                upload(i + 1, nxt)                                    # in flight while batch i computes
            launch(i, d)
            if i >= 1:
                yield finish((i - 1) % 2)
            i += 1
        yield finish((i - 1) % 2)


class GraphedFineTuner:
    """train.FineTuner with the device work replayed from CUDA graphs (a step is ~470 kernel launches on ViT-B/16).

        gt = GraphedFineTuner(train.FineTuner(model, lr=2e-5), max_events=...)
        loss = gt.step(events, offsets, labels)      # loss: static device scalar, valid until the next step
    """

    def __init__(self, tuner, max_events, max_graphs=4):
        self.tuner, self.model = tuner, tuner.model
        self.dev = tuner.flat_p.device
        self.events = torch.zeros((max_events, 4), dtype=torch.float32, device=self.dev)
        self.cache, self.max_graphs = {}, max_graphs
        # Adam's bias correction is computed on the host from the step count, so the two ec_adam launches stay eager; the
        # LoRA re-merge that follows (6 launches per block, pointer-stable) is replayed from its own graph.
        self.refresh_graph, self.refresh_launches = None, 0

    def _refresh_weights(self):
        if self.refresh_graph is None:
            self.tuner.refresh_weights()                 # eager once (also the warm-up)
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            n0 = L.LAUNCHES
            with torch.cuda.graph(g):
                self.tuner.refresh_weights()
            self.refresh_launches = L.LAUNCHES - n0
            L.LAUNCHES = n0
            self.refresh_graph = g
            return
        self.refresh_graph.replay()
        L.LAUNCHES += self.refresh_launches

    def _build(self, plan, labels):
        e = _Entry()
        t = self.tuner
        e.plan = _static_plan(self.model, plan, self.dev)
        e.labels = torch.zeros(plan["B"], dtype=torch.int32, device=self.dev)
        e.labels.copy_(labels)
        torch.cuda.synchronize(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side):
            t.device_forward_backward(self.events, e.plan, e.labels)
        torch.cuda.current_stream(self.dev).wait_stream(side)
        torch.cuda.synchronize(self.dev)
        e.graph = torch.cuda.CUDAGraph()
        n0 = L.LAUNCHES
        with torch.cuda.graph(e.graph):
            e.loss = t.device_forward_backward(self.events, e.plan, e.labels)
            e.out = t.last["out"]
        e.n_launch = L.LAUNCHES - n0
        L.LAUNCHES = n0
        return e

    def forward_backward(self, events, offsets, labels, sel=None):
        plan = self.model.plan_events(offsets, sel)
        n = events.shape[0]
        if n > self.events.shape[0]:
            raise L.ECError(f"batch has {n} events but the graph buffer holds {self.events.shape[0]}")
        labels = labels.to(dtype=torch.int32)
        key = _plan_key(plan)
        ent = self.cache.get(key)
        if ent is None:
            if len(self.cache) >= self.max_graphs:
                self.cache.pop(next(iter(self.cache)))
            self.events[:n].copy_(events, non_blocking=True)
            ent = self.cache[key] = self._build(plan, labels)
        else:
            _refresh_plan(ent.plan, plan)
            ent.labels.copy_(labels, non_blocking=True)
            self.events[:n].copy_(events, non_blocking=True)
        ent.graph.replay()
        L.LAUNCHES += ent.n_launch
        self.tuner.last = {"out": ent.out}
        return ent.loss

    def step(self, events, offsets, labels, sel=None, lr=None, clip_lr=None):
        loss = self.forward_backward(events, offsets, labels, sel)
        self.tuner.allreduce()
        self.tuner.optimizer_step(lr, clip_lr, refresh=False)
        self._refresh_weights()
        return loss
