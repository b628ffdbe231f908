"""Seeded synthetic event streams of each dataset's sensor shape.

There are no datasets offline, so tests and bench.py feed the hot path with these.  Events are float32
[E, 4] rows (x, y, t, p) exactly as the reference loads them (datasets/caltech.py:149-151): integer-valued
coordinates, ascending t in seconds, p in {-1, +1}.
"""
import numpy as np

# sensor (H, W), events per frame N, max_n, count_non_zero, background_mask, typical E, max_t
# (datasets/caltech.py:52-58, cars.py:30-32, imagenet.py:48-50 and configs/zsclip/*_params.py:18-26)
SENSORS = {
    "n_caltech101": dict(shape=(180, 240), N=20000, max_n=225000, count_non_zero=False, background_mask=True,
                         E=100000, max_t=0.325, n_cls=101),
    "n_cars": dict(shape=(100, 120), N=30000, max_n=12500, count_non_zero=True, background_mask=False,
                   E=4000, max_t=0.1, n_cls=2),
    "n_imagenet": dict(shape=(480, 640), N=70000, max_n=135000, count_non_zero=False, background_mask=True,
                       E=1000000, max_t=0.055, n_cls=1000),
}


def synth_events(shape, E, seed, kind="uniform", max_t=0.1):
    """One sample's stream.  kind: 'uniform' | 'clustered' (50 Gaussian blobs, sigma 8 px) |
    'hotpixel' (one pixel receives 2 % of the events)."""
    H, W = shape
    rng = np.random.default_rng(seed)
    if kind == "clustered":
        cx = rng.uniform(0, W, 50)
        cy = rng.uniform(0, H, 50)
        b = rng.integers(0, 50, E)
        x = np.clip(np.rint(cx[b] + rng.normal(0, 8.0, E)), 0, W - 1)
        y = np.clip(np.rint(cy[b] + rng.normal(0, 8.0, E)), 0, H - 1)
    else:
        x = rng.integers(0, W, E).astype(np.float64)
        y = rng.integers(0, H, E).astype(np.float64)
        if kind == "hotpixel":
            hot = rng.random(E) < 0.02
            x[hot] = W // 3
            y[hot] = H // 2
    t = np.sort(rng.uniform(0, max_t, E))
    p = np.where(rng.random(E) < 0.5, -1.0, 1.0)
    return np.stack([x, y, t, p], axis=1).astype(np.float32)


def synth_class_events(shape, E, cls, seed, max_t=0.1):
    """One sample of a LABELLED synthetic stream: class `cls` fixes a layout of four Gaussian blobs (drawn from a generator
    seeded by the class id alone); the sample's own seed jitters the blob centres (+-3 % of the sensor), draws the blob
    weights and spreads 10 % of the events uniformly.  Gives bench.py / the tests a task with real labels and per-sample
    differences (the reference's datasets are not available offline)."""
    H, W = shape
    lay = np.random.default_rng(777 + 7919 * int(cls))
    cx, cy = lay.uniform(0.15 * W, 0.85 * W, 4), lay.uniform(0.15 * H, 0.85 * H, 4)
    sig = lay.uniform(0.03, 0.08, 4) * W
    rng = np.random.default_rng(seed)
    cx = cx + rng.normal(0, 0.03 * W, 4)
    cy = cy + rng.normal(0, 0.03 * H, 4)
    wts = rng.dirichlet(np.full(4, 4.0))
    nb = int(0.9 * E)
    b = rng.choice(4, nb, p=wts)
    x = np.clip(np.rint(cx[b] + rng.normal(0, 1, nb) * sig[b]), 0, W - 1)
    y = np.clip(np.rint(cy[b] + rng.normal(0, 1, nb) * sig[b]), 0, H - 1)
    x = np.concatenate([x, rng.integers(0, W, E - nb).astype(np.float64)])
    y = np.concatenate([y, rng.integers(0, H, E - nb).astype(np.float64)])
    perm = rng.permutation(E)
    t = np.sort(rng.uniform(0, max_t, E))
    p = np.where(rng.random(E) < 0.5, -1.0, 1.0)
    return np.stack([x[perm], y[perm], t, p], axis=1).astype(np.float32)


def synth_labeled_batch(dataset, B, seed0, n_cls=None, E=None):
    """B labelled streams (class = (seed0 + i) mod n_cls): (events [sum E, 4], offsets int64 [B+1], labels int64 [B])."""
    cfg = SENSORS[dataset]
    n_cls = cfg["n_cls"] if n_cls is None else n_cls
    E = cfg["E"] if E is None else E
    labels = (seed0 + np.arange(B)) % n_cls
    evs = [synth_class_events(cfg["shape"], E, labels[i], seed0 + i, cfg["max_t"]) for i in range(B)]
    offsets = np.zeros(B + 1, np.int64)
    offsets[1:] = np.cumsum([len(e) for e in evs])
    return np.concatenate(evs, axis=0), offsets, labels.astype(np.int64)


def calibrate_text_feats(feats, n_cls, labels=None):
    """Synthetic text features for random-init towers (no text checkpoints offline).  A random-init ViT's image features are
    one large common vector plus a small input-dependent part (ViT-B/16: norm 21.7, spread 0.17), so Gaussian text features
    make every sample predict the same class.  These are built from the image features of a calibration batch instead, all of
    them orthogonal to the common direction m = mean(feats), so that a sample scores  f . t_c = (P f) . t_c  with P the
    projector that removes m: the common-mode part cancels and the predictions split.
      labels given : t_c = normalise(P (mean of class c - mean of the class means))
      n_cls == 2   : t_0 = -t_1 = first principal direction of P (f - m), tilted along m so that the median calibration
                     sample sits on the decision boundary (an even split)
      otherwise    : t_2k = +v_k, t_2k+1 = -v_k for the principal directions v_k of P (f - m), strongest first (as many as
                     the calibration batch supports): the logits then spread like the features themselves do
    Classes without data get a seeded Gaussian direction projected the same way.  feats: float [n, C] on any device.
    Returns float32 [n_cls, C] on the CPU, rows L2-normalised (what clip_cls.py:84-85 caches)."""
    import torch
    f = feats.detach().double().cpu()
    n, C = f.shape
    g = torch.Generator().manual_seed(4321)
    mid = f.mean(0)
    proto = torch.zeros(n_cls, C, dtype=torch.float64)
    have = torch.zeros(n_cls, dtype=torch.bool)
    if labels is not None:
        lab = torch.as_tensor(labels).long().cpu()
        for c in range(n_cls):
            m = lab == c
            if m.any():
                proto[c], have[c] = f[m].mean(0), True
        mid = proto[have].mean(0)
    mh = mid / mid.norm()
    proj = lambda v: v - (v @ mh)[..., None] * mh
    d = proj(f - mid)
    if labels is None and n_cls == 2:
        v = torch.linalg.svd(d, full_matrices=False)[2][0]
        beta = ((f @ v) / (f @ mh)).median()          # f . (v - beta mh) changes sign at the median sample
        t1 = v - beta * mh
        t1 = t1 / t1.norm()
        return torch.stack([-t1, t1]).float()
    if labels is None:
        U, S, V = torch.linalg.svd(d, full_matrices=False)
        k = min(n_cls // 2, int((S > 1e-9 * S[0]).sum()))
        for j in range(k):
            proto[2 * j], proto[2 * j + 1] = mid + V[j], mid - V[j]
        have[:2 * k] = True
    spread = d.norm(dim=1).mean()
    for c in range(n_cls):
        if not have[c]:
            proto[c] = mid + torch.randn(C, generator=g, dtype=torch.float64) * spread / C ** 0.5
    t = proj(proto - mid)
    return (t / t.norm(dim=1, keepdim=True).clamp_min(1e-30)).float()


def synth_batch(dataset, B, seed0, kind="uniform", E=None):
    """B streams packed the way the fused path takes them: (events float32 [sum E, 4], offsets int64 [B+1])."""
    cfg = SENSORS[dataset]
    E = cfg["E"] if E is None else E
    evs = [synth_events(cfg["shape"], E, seed0 + i, kind, cfg["max_t"]) for i in range(B)]
    offsets = np.zeros(B + 1, np.int64)
    offsets[1:] = np.cumsum([len(e) for e in evs])
    return np.concatenate(evs, axis=0), offsets


def synth_text_feats(n_cls, C, seed):
    """L2-normalised seeded Gaussian [n_cls, C]: stand-in for the cached encode_text output (models/clip_cls.py:84-85)
    when no tokenizer vocabulary / checkpoint is available (bench and scripts; same draw as the oracle's helper)."""
    import torch
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(n_cls, C, generator=g)
    return t / t.norm(dim=-1, keepdim=True)
