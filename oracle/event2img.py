"""ctypes front-end of oracle/event2img_oracle.c (TEST INFRASTRUCTURE ONLY).

Each function names the reference lines it restates (relative to /root/reference).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle_event2img.so")
_lib = None

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)   # method.py:17
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)   # method.py:18


def build(force=False):
    src = os.path.join(_HERE, "event2img_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle_event2img.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        i64p, f32p, u8p = C.POINTER(C.c_int64), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
        L.orc_split_event_count.argtypes = [C.c_int64, C.c_int64, i64p, i64p, C.c_int]
        L.orc_histogram.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, i64p]
        L.orc_frame_from_counts.argtypes = [i64p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, u8p, u8p, i64p]
        L.orc_resized_shape.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_resize_crop_224.argtypes = [u8p, C.c_int, C.c_int, u8p, u8p, u8p]
        L.orc_normalize.argtypes = [u8p, f32p]
        L.orc_events2frames.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, u8p, C.c_int]
        L.orc_event2img_sample.argtypes = [f32p, C.c_int64, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int,
                                           i64p, C.c_int, f32p, u8p]
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _ev(events):
    ev = np.ascontiguousarray(events, dtype=np.float32)
    assert ev.ndim == 2 and ev.shape[1] == 4
    return ev


def _check(rc):
    if rc == -1:
        raise ValueError("event coordinate outside the sensor (numpy bincount/reshape would raise)")
    if rc < 0:
        raise RuntimeError(f"oracle error {rc}")
    return rc


def split_event_count(E, N):
    """vis.py:55-72 -> (idx0, idx1) lists."""
    cap = int(E // max(N, 1)) + 2
    i0 = np.zeros(cap, np.int64)
    i1 = np.zeros(cap, np.int64)
    K = _check(lib().orc_split_event_count(int(E), int(N), _p(i0, C.c_int64), _p(i1, C.c_int64), cap))
    return i0[:K].tolist(), i1[:K].tolist()


def histogram(events, shape):
    """vis.py:44-52 + 9-14 -> int64 [H,W,2] (pos,neg)."""
    ev = _ev(events)
    H, W = shape
    counts = np.zeros((H, W, 2), np.int64)
    _check(lib().orc_histogram(_p(ev, C.c_float), ev.shape[0], H, W, _p(counts, C.c_int64)))
    return counts


def frame_from_counts(counts, count_non_zero=False, background_mask=True, thresh=10):
    """vis.py:16-41 -> (gray uint8 [H,W], zeroed bool [H,W,2], stats dict)."""
    counts = np.ascontiguousarray(counts, np.int64)
    H, W, _ = counts.shape
    gray = np.zeros((H, W), np.uint8)
    zeroed = np.zeros((H, W, 2), np.uint8)
    st = np.zeros(5, np.int64)
    _check(lib().orc_frame_from_counts(_p(counts, C.c_int64), H, W, int(count_non_zero), int(background_mask),
                                       int(thresh), _p(gray, C.c_uint8), _p(zeroed, C.c_uint8), _p(st, C.c_int64)))
    return gray, zeroed.astype(bool), dict(n=int(st[0]), S1=int(st[1]), S2=int(st[2]), max=int(st[3]),
                                           n_zeroed=int(st[4]))


def resized_shape(H, W):
    ho, wo = C.c_int(), C.c_int()
    lib().orc_resized_shape(H, W, C.byref(ho), C.byref(wo))
    return ho.value, wo.value


def resize_crop_224(gray, stages=False):
    """CLIP preprocess Resize(224,BICUBIC)+CenterCrop(224) on one gray plane -> uint8 [224,224]."""
    gray = np.ascontiguousarray(gray, np.uint8)
    H, W = gray.shape
    Ho, Wo = resized_shape(H, W)
    out = np.zeros((224, 224), np.uint8)
    hp = np.zeros((H, Wo), np.uint8)
    full = np.zeros((Ho, Wo), np.uint8)
    _check(lib().orc_resize_crop_224(_p(gray, C.c_uint8), H, W, _p(out, C.c_uint8), _p(hp, C.c_uint8),
                                     _p(full, C.c_uint8)))
    return (out, hp, full) if stages else out


def normalize(u8):
    """ToTensor + Normalize -> float32 [3,224,224]."""
    u8 = np.ascontiguousarray(u8, np.uint8)
    out = np.zeros((3, 224, 224), np.float32)
    lib().orc_normalize(_p(u8, C.c_uint8), _p(out, C.c_float))
    return out


def events2frames(events, shape, N, count_non_zero=False, background_mask=True):
    """vis.py:75-117 with grayscale=True -> uint8 [K,H,W,3]."""
    ev = _ev(events)
    H, W = shape
    cap = int(ev.shape[0] // N) + 2
    fr = np.zeros((cap, H, W), np.uint8)
    K = _check(lib().orc_events2frames(_p(ev, C.c_float), ev.shape[0], H, W, int(N), int(count_non_zero),
                                       int(background_mask), _p(fr, C.c_uint8), cap))
    return np.repeat(fr[:K, :, :, None], 3, axis=3)


def max_imgs(max_n, N, hard_limit):
    """event2img.py:70-72."""
    return max(min(round(max_n / N), hard_limit), 1)


def event2img_sample(events, shape, N, T, count_non_zero=False, background_mask=True, sel=None,
                     only_selected=False):
    """event2img.py:114-128 (+ _subsample_imgs 80-92) -> (img f32 [T,3,224,224], valid bool [T], K)."""
    ev = _ev(events)
    H, W = shape
    img = np.zeros((T, 3, 224, 224), np.float32)
    valid = np.zeros(T, np.uint8)
    selp = None
    if sel is not None:
        sel = np.ascontiguousarray(sel, np.int64)
        selp = _p(sel, C.c_int64)
    K = _check(lib().orc_event2img_sample(_p(ev, C.c_float), ev.shape[0], H, W, int(N), int(T), int(count_non_zero),
                                          int(background_mask), selp, int(only_selected), _p(img, C.c_float),
                                          _p(valid, C.c_uint8)))
    return img, valid.astype(bool), K


def center_events(events, shape):
    """datasets/utils.py:38-57 (float32 arithmetic on a float32 [E,4] array, returns a new array)."""
    ev = np.array(events, dtype=np.float32, copy=True)
    H, W = shape
    ev[:, 2] -= ev[:, 2].min()
    x_min, x_max = ev[:, 0].min(), ev[:, 0].max()
    y_min, y_max = ev[:, 1].min(), ev[:, 1].max()
    x_shift = ((x_max + x_min + np.float32(1.)) - np.float32(W)) // np.float32(2.)
    y_shift = ((y_max + y_min + np.float32(1.)) - np.float32(H)) // np.float32(2.)
    ev[:, 0] -= x_shift
    ev[:, 1] -= y_shift
    return ev


def flip_events(events, W, hflip=False, tflip=False):
    """datasets/utils.py:18-23 (h-flip) and 26-35 (t-flip) with p = 1, in the order event2img.py:100-103 applies them."""
    ev = np.array(events, dtype=np.float32, copy=True)
    if hflip:
        ev[:, 0] = np.float32(W - 1) - ev[:, 0]
    if tflip:
        ev = np.ascontiguousarray(np.flip(ev, axis=0))
        ev[:, 2] = ev[0, 2] - ev[:, 2]
        ev[:, 3] = -ev[:, 3]
    return ev
