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
