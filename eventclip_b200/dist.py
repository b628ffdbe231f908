"""Sample sharding and the final metric reduction of the multi-GPU inference path.

The path shards by sample (every events -> logits unit is independent): rank r takes a
contiguous block of the sample index range, weights and text features are replicated, and the only collective is
one all-reduce of six int64 counters at the end -- the state of the reference's AverageMeters (test.py:55-81):
    [n_samples, top1_probs, top1_logits, top5_probs, top5_logits, n_valid_views]
Backend: NCCL over NVLink on B200 (one process per GPU), gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(n, rank=None, world_size=None):
    """Contiguous block [lo, hi) of n samples for `rank`; sizes differ by at most one."""
    if rank is None:
        rank, world_size = world()
    base, extra = divmod(n, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class AccuracyMeter:
    """Counts what test.py:66-81 averages, as exact integers so that the cross-rank reduction is order-independent."""

    def __init__(self, device="cpu"):
        self.counters = torch.zeros(6, dtype=torch.int64, device=device)
        self._status = []

    def update(self, out_dict, labels):
        st = out_dict.get("status")
        if st is not None and not any(st is s for s in self._status):
            self._status.append(st)          # device status words of the event kernel: read (once) when the meter reports
        labels = labels.to(out_dict["logits"].device).long()
        B = labels.shape[0]
        if "top5_logits" in out_dict:
            t5l, t5p = out_dict["top5_logits"].long(), out_dict["top5_probs"].long()
        else:
            k = min(5, out_dict["logits"].shape[-1])
            t5l = out_dict["logits"].topk(k, dim=-1).indices
            t5p = out_dict["probs"].topk(k, dim=-1).indices
        upd = torch.stack([
            torch.tensor(B, device=labels.device),
            (t5p[:, 0] == labels).sum(), (t5l[:, 0] == labels).sum(),
            (t5p == labels[:, None]).any(-1).sum(), (t5l == labels[:, None]).any(-1).sum(),
            out_dict["valid_masks"].sum().to(labels.device),
        ]).to(self.counters.device)
        self.counters += upd

    def all_reduce(self):
        """The one collective of the inference path.  Before it, the event kernel's status words are checked: bad event
        coordinates raise ValueError here, as numpy does inside the reference's data loader (datasets/vis.py:9-14)."""
        self.check_status()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.counters, op=dist.ReduceOp.SUM)
        return self

    def check_status(self):
        from . import ops
        for st in self._status:
            ops.raise_on_status(st)

    def result(self):
        self.check_status()
        c = self.counters.tolist()
        n = max(c[0], 1)
        return dict(n=c[0], probs_acc=c[1] / n, logits_acc=c[2] / n, probs_acc5=c[3] / n, logits_acc5=c[4] / n,
                    valid_views=c[5])


def gather_predictions(pred):
    """All ranks' per-sample predictions in rank order (parity checks); pred: int tensor [b_local]."""
    rank, ws = world()
    if ws == 1:
        return pred
    sizes = [torch.zeros(1, dtype=torch.int64, device=pred.device) for _ in range(ws)]
    dist.all_gather(sizes, torch.tensor([pred.shape[0]], dtype=torch.int64, device=pred.device))
    m = int(max(s.item() for s in sizes))
    pad = torch.full((m,), -1, dtype=pred.dtype, device=pred.device)
    pad[:pred.shape[0]] = pred
    bufs = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:int(s.item())] for b, s in zip(bufs, sizes)])


class FlatParams:
    """Trainable parameters re-homed as views of ONE flat fp32 buffer with a same-layout flat gradient buffer, so that
    the fine-tune step needs one collective (the reference's DDP buckets the same gradients, SURVEY.md section 8(e):
    LoRA qkvo-16 + 101x512 text features = 1 231 360 floats = 4.9 MB) and one optimizer launch per parameter group.

    groups: list of lists of nn.Parameter (e.g. [outside model.visual, inside model.visual] for the reference's two
    learning rates, method.py:165-182).  Each group occupies one contiguous span."""

    def __init__(self, groups):
        params = [p for g in groups for p in g]
        if not params:
            raise ValueError("FlatParams needs at least one parameter")
        dev = params[0].device
        for p in params:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError("FlatParams keeps fp32 master parameters on one device")
        n = sum(p.numel() for p in params)
        self.flat_p = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.spans, self._where = [], {}
        off = 0
        for g in groups:
            lo = off
            for p in g:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                self._where[id(p)] = (off, k, p.shape)
                off += k
            self.spans.append((lo, off))
        self.numel = n

    def broadcast(self, group=None, src=0):
        """Every rank takes rank `src`'s parameters (DDP's construction-time broadcast); no-op without a process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            gsrc = dist.get_global_rank(group, src) if group is not None else src
            dist.broadcast(self.flat_p, src=gsrc, group=group)
        return self.flat_p

    def grad_view(self, p):
        off, k, shape = self._where[id(p)]
        return self.flat_g[off:off + k].view(shape)

    def average_gradients(self, group=None):
        """The one collective of the fine-tune step: mean of the flat gradient over the ranks, in place."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            if self.flat_g.is_cuda:      # NCCL averages inside the collective
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.AVG, group=group)
            else:                        # gloo (CPU tests) has no AVG
                dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM, group=group)
                self.flat_g.div_(dist.get_world_size(group))
        return self.flat_g
