"""Row F4 of SURVEY.md section 8(f): the 4-view flip test-time augmentation of the pseudo-label generator and its
confidence filtering (reference datasets/event2img.py:94-112, gen_data.py:132-164, 201-226).

The four variants (identity, h-flip, t-flip, h+t-flip) of a packed batch are produced on the device by ec_flip_events and
pushed through the classifier as ONE batch of 4B samples; the selection rules are bookkeeping on [B,4,n_cls] tensors.
The file-system side of gen_data.py (symlinked pseudo-label dataset) stays with the caller.
"""
import numpy as np
import torch

from . import ops
from . import _lib as L

VARIANTS = ((False, False), (True, False), (False, True), (True, True))     # (hflip, tflip) in the reference's order


def tta_events(events, offsets, W):
    """events CUDA float32 [sum E,4], offsets int64 [B+1] (host) -> (events4 [4*sum E,4], offsets4 [4B+1]) in VARIANT-major
    order: sample b of variant v is sample v*B + b (event2img.py:98-102: [events, h, t, h_t])."""
    off = np.asarray(offsets.cpu().numpy() if isinstance(offsets, torch.Tensor) else offsets, dtype=np.int64)
    off_dev = torch.from_numpy(off).to(events.device)
    parts = [events if not (h or t) else ops.flip_events(events, off_dev, W, hflip=h, tflip=t) for h, t in VARIANTS]
    n = int(off[-1])
    off4 = np.concatenate([off[:-1] + v * n for v in range(4)] + [[4 * n]]).astype(np.int64)
    return torch.cat(parts, dim=0), off4


@torch.no_grad()
def tta_forward(model, events, offsets, sel=None):
    """Classifier outputs for the 4 variants of every sample: dict with 'probs' / 'logits' [B,4,n_cls] in the reference's
    layout (gen_data.py:142 `probs.unflatten(0, (-1, 4))`).  `sel`: optional int32 [4B,T] chunk selection."""
    if model.event_frontend is None:
        raise L.ECError("call attach_event_frontend(...) first")
    W = model.event_frontend.resolution[1]
    ev4, off4 = tta_events(events, offsets, W)
    out = model(dict(events=ev4, event_offsets=torch.from_numpy(off4), **({} if sel is None else {"sel_idx": sel})))
    B = (len(off4) - 1) // 4
    lay = lambda t: t.view(4, B, *t.shape[1:]).transpose(0, 1).contiguous()
    return {"probs": lay(out["probs"]), "logits": lay(out["logits"]), "valid_masks": lay(out["valid_masks"])}


def tta_select(probs4, conf_thresh=-1.0, tta_consistent=False, tta_min_prob=False):
    """gen_data.py:141-164 for TTA predictions probs4 [B,4,n_cls] -> dict(probs [B,n_cls], max_probs, pred_labels, sel_mask)."""
    tta_mask = torch.ones(probs4.shape[0], dtype=torch.bool, device=probs4.device)
    if tta_consistent:                       # predictions over the 4 views should agree
        pred = probs4.argmax(dim=-1)
        tta_mask &= (pred[:, 0] == pred[:, 1]) & (pred[:, 0] == pred[:, 2]) & (pred[:, 0] == pred[:, 3])
    if tta_min_prob:                         # the least confident view must clear the threshold too
        tta_mask &= probs4.max(-1).values.min(-1).values > conf_thresh
    probs = probs4.mean(dim=1)
    max_probs, pred_labels = probs.max(dim=-1)
    return {"probs": probs, "max_probs": max_probs, "pred_labels": pred_labels,
            "sel_mask": (max_probs > conf_thresh) & tta_mask}


def select(probs, conf_thresh=-1.0):
    """gen_data.py:155, 160-162 without TTA."""
    max_probs, pred_labels = probs.max(dim=-1)
    return {"probs": probs, "max_probs": max_probs, "pred_labels": pred_labels, "sel_mask": max_probs > conf_thresh}


def topk_per_class(pred_labels, max_probs, sel_mask, n_cls, topk):
    """gen_data.py:201-226: of the selected samples predicted as each class keep the `topk` most confident.
    Returns a bool mask over the samples."""
    keep = torch.zeros_like(sel_mask)
    idx_all = torch.arange(sel_mask.numel(), device=sel_mask.device)
    for c in range(n_cls):
        idx = idx_all[sel_mask & (pred_labels == c)]
        if idx.numel():
            k = min(topk, idx.numel())
            keep[idx[max_probs[idx].topk(k).indices]] = True
    return keep
