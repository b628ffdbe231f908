"""Row F2: on-disk event formats and the compact wire format.
CPU: the loaders (product and oracle restatement) reproduce what the reference's own loaders returned for the same
synthetic files (tests/golden/formats_sha.json); the vectorised host packer equals the oracle's scalar definition.
GPU: ec_pack_events equals the host packer bit for bit; ec_event2img_compact yields the frames of the float path
bit for bit (counts, gray bytes, resized bytes, output tensor), including aliasing / ignored / rejected events."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from eventclip_b200.datasets import formats
from eventclip_b200.synth import SENSORS, synth_batch
from oracle import formats_oracle


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _write_cases(tmp_path):
    """Re-creates the files make_golden.make_formats wrote (same seeds)."""
    cases = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "formats_sha.json")))
    for c in cases:
        rng = np.random.default_rng(c["seed"])
        E = c["E"]
        if c["kind"] == "npz":
            rec = np.zeros(E, dtype=[("x", "<u2"), ("y", "<u2"), ("t", "<i8"), ("p", "?" if c["pol01"] else "<i2")])
            rec["x"], rec["y"] = rng.integers(0, 640, E), rng.integers(0, 480, E)
            rec["t"] = np.sort(rng.integers(0, 55000, E))
            rec["p"] = rng.integers(0, 2, E) if c["pol01"] else rng.choice([-1, 1], E)
            path = str(tmp_path / f"{c['seed']}.npz")
            np.savez(path, event_data=rec)
        else:
            arr = np.stack([rng.integers(0, 240, E), rng.integers(0, 180, E), np.sort(rng.random(E)), rng.choice([-1, 1], E)], 1)
            path = str(tmp_path / f"{c['seed']}.npy")
            np.save(path, arr)
        yield c, path


def test_loaders_vs_reference_golden(tmp_path):
    n = 0
    for c, path in _write_cases(tmp_path):
        for load in ((formats.load_events_npz, formats_oracle.load_npz) if c["kind"] == "npz" else
                     (formats.load_events_npy, formats_oracle.load_npy)):
            out = load(path)
            assert str(out.dtype) == c["dtype"] and out.shape == (c["E"], 4)
            assert sha(out) == c["sha"], (c["kind"], c["seed"], load.__module__)
        n += 1
    assert n == 5


def _tricky_events(shape, E, seed):
    """Synthetic stream plus the cases the reference's flat-index histogram treats specially (vis.py:9-14)."""
    H, W = shape
    ev, _ = synth_batch("n_caltech101", 1, seed, E=E) if shape == (180, 240) else synth_batch("n_cars", 1, seed, E=E)
    ev = ev.copy()
    ev[5] = (W + 3, 2, ev[5, 2], 1)          # x >= W aliases into the next row
    ev[6] = (-1, 1, ev[6, 2], -1)            # negative x with a non-negative flat index: legal for np.bincount
    ev[7] = (10.9, 20.9, ev[7, 2], 1)        # fractional coordinates truncate
    ev[8] = (-50, 0, ev[8, 2], 0)            # p == 0: never histogrammed, never range-checked
    ev[9] = (3, 4, ev[9, 2], 0.7)            # polarity truncates to 0 -> ignored
    return ev


@pytest.mark.parametrize("shape", [(180, 240), (100, 120)])
def test_host_packer_vs_oracle_definition(shape):
    ev = _tricky_events(shape, 4000, 3)
    w = formats.pack_events_host(ev, shape)
    assert w.dtype == np.uint32 and np.array_equal(w, formats_oracle.pack_events(ev, shape))
    assert w[8] == 0 and w[9] == 0 and (w[5] & 0x3fffffff) == 2 * shape[1] + shape[1] + 3
    bad = ev.copy()
    bad[0] = (0, shape[0], 0, 1)             # flat index == H*W: rejected
    assert formats.pack_events_host(bad, shape)[0] == np.uint32(3) << np.uint32(30)


@pytest.mark.gpu
@pytest.mark.parametrize("ds", ["n_caltech101", "n_cars", "n_imagenet"])
def test_compact_path_bit_exact(cuda_dev, ds):
    from eventclip_b200 import ops
    cfg = SENSORS[ds]
    shape = cfg["shape"]
    B = 3
    ev, off = synth_batch(ds, B, 60, kind="clustered")
    ev[5] = (shape[1] + 3, 2, ev[5, 2], 1)
    ev[6] = (-1, 1, ev[6, 2], -1)
    ev[7] = (10.9, 20.9, ev[7, 2], 1)
    ev[8] = (-50, 0, ev[8, 2], 0)
    evd = torch.from_numpy(ev).to(cuda_dev)
    words = ops.pack_events(evd, shape)
    host = formats.pack_events_host(ev, shape)
    assert np.array_equal(words.cpu().numpy().view(np.uint32), host)
    T = 3
    frames, valid, chunks, nv = ops.plan_frames(off, cfg["N"], T)
    fd = frames.to(cuda_dev)
    for out, patch in (("f32", 0), ("patch", 16)):
        a, sa, da = ops.event2img(evd, fd, shape, B * T, cfg["count_non_zero"], cfg["background_mask"], out=out, patch=patch,
                                  debug=True)
        b, sb, db = ops.event2img(words, fd, shape, B * T, cfg["count_non_zero"], cfg["background_mask"], out=out, patch=patch,
                                  debug=True)
        torch.cuda.synchronize()
        assert sa.item() == 0 and sb.item() == 0
        assert torch.equal(a, b)
        for k in ("counts", "gray", "u8"):
            assert torch.equal(da[k], db[k]), k
    # a rejected index raises the same status bit as the float path's out-of-range coordinate
    ev2 = ev.copy()
    ev2[0] = (0, shape[0], ev2[0, 2], 1)
    w2 = torch.from_numpy(formats.pack_events_host(ev2, shape).view(np.int32)).to(cuda_dev)
    _, st_c, _ = ops.event2img(w2, fd, shape, B * T, cfg["count_non_zero"], cfg["background_mask"])
    _, st_f, _ = ops.event2img(torch.from_numpy(ev2).to(cuda_dev), fd, shape, B * T, cfg["count_non_zero"], cfg["background_mask"])
    assert st_c.item() == st_f.item() != 0
    with pytest.raises(ValueError):
        ops.raise_on_status(st_c)
