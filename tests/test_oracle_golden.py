"""CPU: the oracle (oracle/) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  This is what pins the oracle; the GPU tests then compare CUDA against the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import event2img as orc
from eventclip_b200.synth import SENSORS, synth_events


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_split_event_count_cases(golden_dir):
    cases = json.load(open(os.path.join(golden_dir, "split_cases.json")))
    assert len(cases) >= 20
    for c in cases:
        i0, i1 = orc.split_event_count(c["E"], c["N"])
        assert i0 == c["idx0"] and i1 == c["idx1"], c


def test_gray_lut_exhaustive(golden_dir):
    """uint8 gray value as a function of (pos, neg, max): vis.py:27-39, both background_mask settings."""
    z = np.load(os.path.join(golden_dir, "gray_lut.npz"))
    for key, mask in (("mask", True), ("nomask", False)):
        tab = z[key]
        for mx in np.unique(tab[:, 0]):
            rows = tab[tab[:, 0] == mx]
            n = len(rows)
            counts = np.zeros((1, n + 1, 2), np.int64)
            counts[0, :n, 0] = rows[:, 1]
            counts[0, :n, 1] = rows[:, 2]
            counts[0, n, 0] = mx                       # pins hist.max()
            gray, _, st = orc.frame_from_counts(counts, False, mask, thresh=0)
            assert st["max"] == mx
            assert (gray[0, :n] == rows[:, 3]).all(), (key, mx)


def test_small_sensor_all_stages(golden_dir):
    z = np.load(os.path.join(golden_dir, "event2img_small.npz"))
    for name in "abc":
        H, W, N, cnz, bg = [int(v) for v in z[f"{name}_cfg"]]
        ev = z[f"{name}_events"]
        i0, i1 = orc.split_event_count(len(ev), N)
        ref_counts, ref_frames, ref_u8 = z[f"{name}_counts"], z[f"{name}_frames"], z[f"{name}_u8"]
        assert len(i0) == ref_counts.shape[0]
        imgs = []
        for k, (a, b) in enumerate(zip(i0, i1)):
            counts = orc.histogram(ev[a:b], (H, W))
            assert (counts == ref_counts[k]).all()
            gray, _, _ = orc.frame_from_counts(counts, bool(cnz), bool(bg))
            assert (gray == ref_frames[k]).all()
            u8 = orc.resize_crop_224(gray)
            assert (u8 == ref_u8[k]).all()
            imgs.append(orc.normalize(u8))
        assert sha(np.stack(imgs)) == bytes(z[f"{name}_img_sha"]).hex()
        fr = orc.events2frames(ev, (H, W), N, bool(cnz), bool(bg))
        assert (fr[..., 0] == ref_frames).all() and (fr[..., 1] == fr[..., 2]).all()


@pytest.mark.parametrize("idx", range(9))
def test_real_sensor_checksums(golden_dir, idx):
    c = json.load(open(os.path.join(golden_dir, "event2img_sha.json")))[idx]
    cfg = SENSORS[c["dataset"]]
    ev = synth_events(cfg["shape"], c["E"], c["seed"], c["kind"], cfg["max_t"])
    assert sha(ev) == c["events"], "synthetic generator drifted from the one the goldens were made with"
    i0, i1 = orc.split_event_count(len(ev), cfg["N"])
    assert len(i0) == c["K"]
    counts = np.stack([orc.histogram(ev[a:b], cfg["shape"]) for a, b in zip(i0, i1)])
    assert sha(counts.astype(np.int32)) == c["counts"]
    grays = np.stack([orc.frame_from_counts(k, cfg["count_non_zero"], cfg["background_mask"])[0] for k in counts])
    assert sha(grays) == c["frames"]
    u8 = np.stack([orc.resize_crop_224(g) for g in grays])
    assert sha(u8) == c["u8"]
    img = np.stack([orc.normalize(u) for u in u8])
    assert sha(img) == c["img"]


@pytest.mark.parametrize("idx", range(30))
def test_other_sensor_shapes_checksums(golden_dir, idx):
    """The oracle against the UNMODIFIED reference on five more sensor shapes (N-MNIST 34x34, DVS128, portrait 240x180, wide
    64x200, DAVIS346 260x346): other aspect ratios, the crop on the other axis, W % 4 != 0; both flag settings."""
    c = json.load(open(os.path.join(golden_dir, "event2img_shapes_sha.json")))[idx]
    shape = tuple(c["shape"])
    ev = synth_events(shape, c["E"], c["seed"], c["kind"])
    assert sha(ev) == c["events"], "synthetic generator drifted from the one the goldens were made with"
    i0, i1 = orc.split_event_count(len(ev), c["N"])
    assert len(i0) == c["K"]
    counts = np.stack([orc.histogram(ev[a:b], shape) for a, b in zip(i0, i1)])
    assert sha(counts.astype(np.int32)) == c["counts"]
    grays = np.stack([orc.frame_from_counts(k, c["count_non_zero"], c["background_mask"])[0] for k in counts])
    assert sha(grays) == c["frames"]
    u8 = np.stack([orc.resize_crop_224(g) for g in grays])
    assert sha(u8) == c["u8"]
    assert sha(np.stack([orc.normalize(u) for u in u8])) == c["img"]


def test_event2img_sample_padding_and_selection():
    cfg = SENSORS["n_caltech101"]
    ev = synth_events(cfg["shape"], 50001, 3, "uniform")
    img, valid, K = orc.event2img_sample(ev, cfg["shape"], cfg["N"], 10, False, True)
    assert K == 3 and valid.tolist() == [True] * 3 + [False] * 7
    assert (img[3:] == 0).all() and np.abs(img[:3]).sum() > 0
    img2, valid2, _ = orc.event2img_sample(ev, cfg["shape"], cfg["N"], 2, False, True, sel=[2, 0])
    assert valid2.all() and (img2[0] == img[2]).all() and (img2[1] == img[0]).all()
    img3, _, _ = orc.event2img_sample(ev, cfg["shape"], cfg["N"], 2, False, True, sel=[2, 0], only_selected=True)
    assert (img3 == img2).all()


def test_out_of_range_coordinates_raise():
    ev = synth_events((100, 120), 100, 1)
    ev[5, 1] = 100            # y == H -> flat index >= H*W for x > 0 ... force the last row overflow
    ev[5, 0] = 119
    with pytest.raises(ValueError):
        orc.histogram(ev, (100, 120))
    ev[5, 0], ev[5, 1] = -1, 0
    with pytest.raises(ValueError):
        orc.histogram(ev, (100, 120))


def test_hot_pixel_rule_ties_and_constant():
    # all bins equal: std = 0, threshold = mean = c, c > c is false -> nothing removed
    counts = np.full((4, 4, 2), 7, np.int64)
    gray, zeroed, st = orc.frame_from_counts(counts, False, True)
    assert not zeroed.any() and st["max"] == 7
    # one dominant bin is removed and the max is taken over the rest (vis.py:23-27)
    counts = np.zeros((20, 20, 2), np.int64)
    counts[3, 4, 0] = 500
    counts[5, 5, 1] = 3
    counts[6, 6, 0] = 2
    gray, zeroed, st = orc.frame_from_counts(counts, False, True)
    assert zeroed[3, 4, 0] and zeroed.sum() == 1 and st["max"] == 3
    assert gray[3, 4] == 255 and gray[0, 0] == 255 and gray[5, 5] == 127


def _transform_case(c):
    shape = tuple(c["shape"])
    ev = synth_events(shape, c["E"], c["seed"], "clustered")
    ev = ev[(ev[:, 0] < shape[1] * 0.6) & (ev[:, 1] < shape[0] * 0.7)].copy()
    ev[:, 2] += np.float32(0.37)
    assert len(ev) == c["n"] and sha(ev) == c["events"]
    return shape, ev


def test_center_and_flip_events_vs_reference_golden(golden_dir):
    """datasets/utils.py:18-57."""
    for c in json.load(open(os.path.join(golden_dir, "event_transforms_sha.json"))):
        shape, ev = _transform_case(c)
        cen = orc.center_events(ev, shape)
        assert sha(cen) == c["centered"] and cen[:, 2].min() == 0
        assert sha(orc.flip_events(ev, shape[1], True, False)) == c["hflip"]
        assert sha(orc.flip_events(ev, shape[1], False, True)) == c["tflip"]
        assert sha(orc.flip_events(ev, shape[1], True, True)) == c["htflip"]
