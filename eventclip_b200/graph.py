"""CUDA-graph replay of the fused events -> logits forward for fixed batch geometry.

The device part of the classifier forward (`device_forward`: ec_event2img -> patch GEMM -> ViT blocks -> head, ~100
kernel launches) is captured once and replayed, so a step costs one graph launch instead of ~100 ctypes calls.
A graph is specific to (number of frames, number of valid views, slot map); batches whose plan differs get their
own graph (`GraphedClassifier` keeps a small cache keyed by the plan).
"""
import torch

from . import _lib as L


class _Entry:
    pass


class GraphedClassifier:
    def __init__(self, model, max_events, max_graphs=8):
        self.model = model
        self.dev = model.device
        if self.dev.type != "cuda":
            raise L.ECError("GraphedClassifier needs the model on a CUDA device")
        self.events = torch.zeros((max_events, 4), dtype=torch.float32, device=self.dev)   # static input buffer
        self.status = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self.cache = {}
        self.max_graphs = max_graphs

    def _key(self, plan):
        ros = plan["row_of_slot"]
        return (plan["n_valid"], plan["B"], plan["T"], plan["frames"].numpy().tobytes(),
                None if ros is None else ros.numpy().tobytes())

    def _build(self, plan):
        e = _Entry()
        e.plan = self.model.plan_to_device(plan, self.dev)
        torch.cuda.synchronize(self.dev)
        side = torch.cuda.Stream(device=self.dev)
        side.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(side), torch.no_grad():      # warm-up outside capture: tables, packed weights, attributes
            for _ in range(2):
                self.model.device_forward(self.events, e.plan, status=self.status)
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
        """data_dict: {'events' float32 [sum E,4] (pinned host or CUDA), 'event_offsets', optional 'sel_idx'}.
        Returns the classifier's out_dict; its tensors are the graph's static outputs (overwritten by the next call)."""
        plan = self.model.plan_events(data_dict["event_offsets"], data_dict.get("sel_idx", None))
        ev = data_dict["events"]
        n = ev.shape[0]
        if n > self.events.shape[0]:
            raise L.ECError(f"batch has {n} events but the graph buffer holds {self.events.shape[0]}")
        key = self._key(plan)
        ent = self.cache.get(key)
        if ent is None:
            if len(self.cache) >= self.max_graphs:
                self.cache.pop(next(iter(self.cache)))
            ent = self.cache[key] = self._build(plan)
        self.events[:n].copy_(ev, non_blocking=True)
        ent.graph.replay()
        L.LAUNCHES += ent.n_launch
        return ent.out
