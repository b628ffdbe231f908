"""oracle/formats_oracle.py -- TEST INFRASTRUCTURE ONLY.

Restatement of the reference's event-file loaders (datasets/imagenet.py:8-27, datasets/caltech.py:149-151) and the
definition of the compact wire format (row F2) in plain Python loops.  The loaders are pinned against the unmodified
reference functions run on synthetic files (tests/golden/make_golden.py::make_formats -> formats_sha.json).
"""
import numpy as np


def load_npz(path):
    """imagenet.py:8-27."""
    rec = np.load(path)["event_data"]
    out = np.zeros((len(rec), 4), np.float64)
    out[:, 0], out[:, 1], out[:, 2] = rec["x"], rec["y"], rec["t"]
    out[:, 3] = rec["p"].astype(np.uint8)
    out[:, 2] = out[:, 2] / 1e6
    if out[:, 3].min() >= -0.5:
        neg = out[:, 3] <= 0.5
        out[neg, 3] = -1
    return out


def load_npy(path):
    """caltech.py:149-151."""
    return np.load(path).astype(np.float32)


def pack_word(x, y, p, H, W):
    """One event -> compact word (scalar definition; floats truncated toward zero like ndarray.astype(int))."""
    xi, yi, pi = int(np.float32(x)), int(np.float32(y)), int(np.float32(p))
    if pi == 0:
        return 0
    flat = xi + yi * W
    if flat < 0 or flat >= H * W:
        return 3 << 30
    return flat | ((1 if pi > 0 else 2) << 30)


def pack_events(events, shape):
    H, W = shape
    return np.array([pack_word(e[0], e[1], e[3], H, W) for e in np.asarray(events, np.float32)], dtype=np.uint32)
