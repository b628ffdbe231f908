"""oracle/tta_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference's flip-TTA selection (gen_data.py:141-164) and per-class top-k (gen_data.py:201-226) on
numpy arrays, and of the 4-variant event list (datasets/event2img.py:94-112).  Pinned by tests/golden/tta_golden.npz, whose
generator executes the reference's own source lines (tests/golden/make_golden.py::make_tta).
"""
import numpy as np

from . import event2img as orc


def tta_variants(events, shape):
    """[events, h, t, h_t] (event2img.py:98-102)."""
    W = shape[1]
    h = orc.flip_events(events, W, hflip=True)
    t = orc.flip_events(events, W, tflip=True)
    ht = orc.flip_events(h, W, tflip=True)
    return [np.asarray(events, np.float32), h, t, ht]


def tta_select(probs4, conf_thresh, tta_consistent, tta_min_prob):
    probs4 = np.asarray(probs4, np.float32)
    mask = np.ones(probs4.shape[0], bool)
    if tta_consistent:
        pred = probs4.argmax(-1)
        mask &= (pred[:, 0] == pred[:, 1]) & (pred[:, 0] == pred[:, 2]) & (pred[:, 0] == pred[:, 3])
    if tta_min_prob:
        mask &= probs4.max(-1).min(-1) > np.float32(conf_thresh)
    probs = probs4.mean(1, dtype=np.float32)
    return dict(probs=probs, max_probs=probs.max(-1), pred_labels=probs.argmax(-1),
                sel_mask=(probs.max(-1) > np.float32(conf_thresh)) & mask)


def topk_per_class(pred_labels, max_probs, sel_mask, n_cls, topk):
    keep = np.zeros_like(sel_mask)
    for c in range(n_cls):
        idx = np.nonzero(sel_mask & (pred_labels == c))[0]
        if len(idx):
            order = np.argsort(-max_probs[idx], kind="stable")[:min(topk, len(idx))]
            keep[idx[order]] = True
    return keep
