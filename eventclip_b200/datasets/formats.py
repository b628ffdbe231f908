"""Row F2 of SURVEY.md section 8(f): on-disk event formats -> packed events, and the compact wire format.

Host-side mirrors of the reference's two loaders (same return values):
  load_events_npy   datasets/caltech.py:149-151   `.npy` float [E,4] rows (x, y, t, p)         (N-Caltech101, N-Cars)
  load_events_npz   datasets/imagenet.py:8-27     `.npz` structured `event_data{x,y,t,p}`, integer microsecond timestamps,
                                                  0/1 polarity                                   (N-ImageNet)
and the packer of the compact format consumed by ec_event2img_compact (include/eventclip_b200.h): one 32-bit word per
event = flat pixel index (x + y*W as np.bincount sees it at datasets/vis.py:9-14) | polarity code << 30.  Packing on the
host cuts the H2D copy from 16 to 4 bytes per event; `ops.pack_events` does the same on the device.
"""
import numpy as np


def load_events_npy(path):
    """caltech.py:149-151."""
    return np.load(path).astype(np.float32)


def load_events_npz(path):
    """imagenet.py:8-27: float64 [E,4]; t in seconds; polarity {0,1} -> {-1,+1} when no negative polarity is present."""
    event = np.load(path)["event_data"]
    event = np.stack([event["x"], event["y"], event["t"], event["p"].astype(np.uint8)], 1)
    event = event.astype(float)
    event[:, 2] /= 1e6
    if event[:, 3].min() >= -0.5:
        event[:, 3][event[:, 3] <= 0.5] = -1
    return event


def pack_events_host(events, shape):
    """float [E,4] -> uint32 [E] compact words.  Coordinates and polarity are truncated toward zero like the reference's
    `.astype(int)` (vis.py:46-50); code 3 marks an index outside [0, H*W) (the reference raises when it meets one)."""
    ev = np.asarray(events, dtype=np.float32)
    H, W = shape
    if H * W >= 1 << 30:
        raise ValueError("sensor too large for the 30-bit index of the compact format")
    x, y, p = ev[:, 0].astype(np.int64), ev[:, 1].astype(np.int64), ev[:, 3].astype(np.int64)
    flat = x + y * W
    ok = (flat >= 0) & (flat < H * W)
    code = np.where(p > 0, 1, 2).astype(np.uint32)
    word = np.where(ok, flat.astype(np.uint32) | (code << np.uint32(30)), np.uint32(3) << np.uint32(30))
    return np.where(p != 0, word, np.uint32(0)).astype(np.uint32)
