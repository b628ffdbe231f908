"""GPU parity: the fused event2img kernel (through the C ABI) against the oracle and the golden fixtures.
Integer stages are bit-exact; the float32 output is bit-exact (LUT of IEEE float32 values); bf16 = RNE of it."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from eventclip_b200 import ops, _lib
from eventclip_b200.datasets import Event2Image, events2frames
from eventclip_b200.synth import SENSORS, synth_batch, synth_events
from oracle import event2img as orc

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _qargs(cfg, max_imgs=10):
    return dict(max_imgs=max_imgs, N=cfg["N"], split_method="event_count", convert_method="event_histogram",
                grayscale=True, count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])


def _run_frames(ev, shape, N, cnz, bg, dev, out="f32"):
    E = len(ev)
    T = int(E // N) + 2
    frames, valid, chunks, K = ops.plan_frames([0, E], N, T, compact=True)
    img, status, dbg = ops.event2img(torch.from_numpy(ev).to(dev), frames.to(dev), shape, K, cnz, bg, out=out, debug=True)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    return img, dbg, K


def test_small_sensor_all_stages_vs_golden(cuda_dev, golden_dir):
    z = np.load(os.path.join(golden_dir, "event2img_small.npz"))
    for name in "abc":
        H, W, N, cnz, bg = [int(v) for v in z[f"{name}_cfg"]]
        img, dbg, K = _run_frames(z[f"{name}_events"], (H, W), N, bool(cnz), bool(bg), cuda_dev)
        assert K == z[f"{name}_counts"].shape[0]
        assert (dbg["counts"].cpu().numpy() == z[f"{name}_counts"]).all(), name
        assert (dbg["gray"].cpu().numpy() == z[f"{name}_frames"]).all(), name
        assert (dbg["u8"].cpu().numpy() == z[f"{name}_u8"]).all(), name
        assert sha(img.cpu().numpy()) == bytes(z[f"{name}_img_sha"]).hex(), name


@pytest.mark.parametrize("idx", range(9))
def test_real_sensors_vs_golden_checksums(cuda_dev, golden_dir, idx):
    c = json.load(open(os.path.join(golden_dir, "event2img_sha.json")))[idx]
    cfg = SENSORS[c["dataset"]]
    ev = synth_events(cfg["shape"], c["E"], c["seed"], c["kind"], cfg["max_t"])
    assert sha(ev) == c["events"]
    img, dbg, K = _run_frames(ev, cfg["shape"], cfg["N"], cfg["count_non_zero"], cfg["background_mask"], cuda_dev)
    assert K == c["K"]
    assert sha(dbg["counts"].cpu().numpy()) == c["counts"]
    assert sha(dbg["gray"].cpu().numpy()) == c["frames"]
    assert sha(dbg["u8"].cpu().numpy()) == c["u8"]
    assert sha(img.cpu().numpy()) == c["img"]


@pytest.mark.parametrize("ds", ["n_caltech101", "n_cars", "n_imagenet"])
@pytest.mark.parametrize("kind", ["uniform", "clustered", "hotpixel"])
def test_batch_vs_oracle_every_stage(cuda_dev, ds, kind):
    """Packed batch with ragged lengths: padding, overlapping tail chunks, host-drawn view selection."""
    cfg = SENSORS[ds]
    N = cfg["N"]
    lens = {"n_caltech101": [100000, 50001, 19999, 20000, 30000, 30001],
            "n_cars": [4000, 12500, 300, 45001],
            "n_imagenet": [250000, 70000, 69999, 105001]}[ds]
    evs = [synth_events(cfg["shape"], E, 100 + i, kind, cfg["max_t"]) for i, E in enumerate(lens)]
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    e2i = Event2Image(_qargs(cfg), cfg["shape"], cfg["max_n"])
    T = e2i.max_imgs
    torch.manual_seed(5)
    sel = e2i.draw_selection(off)
    r = e2i(torch.from_numpy(np.concatenate(evs)).to(cuda_dev), off, sel=sel, out="f32", debug=True, check=True)
    img = r["img"].cpu().numpy()
    valid = r["valid_mask"].numpy()
    fi = 0
    rec = np.frombuffer(r["frames"].numpy().tobytes(), dtype=[("s", "<i8"), ("n", "<i4"), ("o", "<i4")])
    for b, ev in enumerate(evs):
        oimg, ovalid, K = orc.event2img_sample(ev, cfg["shape"], N, T, cfg["count_non_zero"], cfg["background_mask"],
                                               sel=sel[b] if K_gt(ev, N, T) else None, only_selected=True)
        assert (valid[b] == ovalid).all() and int(r["chunks"][b]) == K
        assert (img[b] == oimg).all(), (ds, kind, b)
        i0, i1 = orc.split_event_count(len(ev), N)
        for t in range(T):
            if rec[fi]["n"] > 0:
                k = int(sel[b][t]) if K > T else t
                counts = orc.histogram(ev[i0[k]:i1[k]], cfg["shape"])
                assert (r["debug"]["counts"][fi].cpu().numpy() == counts).all()
                gray, _, _ = orc.frame_from_counts(counts, cfg["count_non_zero"], cfg["background_mask"])
                assert (r["debug"]["gray"][fi].cpu().numpy() == gray).all()
                assert (r["debug"]["u8"][fi].cpu().numpy() == orc.resize_crop_224(gray)).all()
            fi += 1


def K_gt(ev, N, T):
    return len(orc.split_event_count(len(ev), N)[0]) > T


def test_output_formats_agree(cuda_dev):
    """bf16 NCHW = round-to-nearest-even of the float32 tensor; patch layout = its im2col; compact = gather."""
    cfg = SENSORS["n_caltech101"]
    ev, off = synth_batch("n_caltech101", 3, 40, E=50001)
    evd = torch.from_numpy(ev).to(cuda_dev)
    e2i = Event2Image(_qargs(cfg), cfg["shape"], cfg["max_n"])
    f32 = e2i(evd, off, out="f32")
    b16 = e2i(evd, off, out="bf16")
    assert torch.equal(b16["img"], f32["img"].to(torch.bfloat16))
    valid = f32["valid_mask"]
    assert (f32["img"][~valid.to(cuda_dev)] == 0).all()
    for P, ldk in ((32, 3072), (16, 768), (14, 592)):
        pt = e2i(evd, off, out="patch", compact=True, patch=P, ldk=ldk)
        G = 224 // P
        ref = f32["img"][valid.to(cuda_dev)].to(torch.bfloat16)                 # [Nv,3,224,224]
        ref = ref.view(-1, 3, G, P, G, P).permute(0, 2, 4, 1, 3, 5).reshape(-1, 3 * P * P)
        assert pt["n_valid"] == int(valid.sum())
        assert torch.equal(pt["img"][:, :3 * P * P], ref)
        assert (pt["img"][:, 3 * P * P:] == 0).all()
        assert torch.equal(ops.im2col(f32["img"][valid.to(cuda_dev)].contiguous(), P, ldk), pt["img"])


@pytest.mark.parametrize("ds", ["n_caltech101", "n_cars", "n_imagenet"])
def test_gray_patch_format(cuda_dev, ds):
    """EC_OUT_GRAY_*_PATCH: one plane per patch row holding the resampled byte / 128 exactly (fp16 and bf16), for every kernel
    family (single-CTA tensor-core, band-exchange cluster) and patch size (32, 16, 14 = the narrow store path): equal to the
    debug output of the resampled bytes, which the other tests pin against the reference."""
    cfg = SENSORS[ds]
    ev, off = synth_batch(ds, 2, 91, kind="clustered")
    frames, valid, chunks, nv = ops.plan_frames(off, cfg["N"], 2, compact=True)
    evd, fd = torch.from_numpy(ev).to(cuda_dev), frames.to(cuda_dev)
    _, _, dbg = ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="f32", debug=True)
    u8 = dbg["u8"].cpu().float()                                                        # [nv, 224, 224]
    for P, ldk in ((32, 1024), (16, 256), (14, 200)):
        G = 224 // P
        want = (u8 / 128.0).view(nv, G, P, G, P).permute(0, 1, 3, 2, 4).reshape(nv * G * G, P * P)
        for fmt, dt in (("gray_f16", torch.float16), ("gray", torch.bfloat16)):
            pt, st, _ = ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out=fmt, patch=P, ldk=ldk)
            assert int(st.item()) == 0 and pt.dtype == dt and pt.shape == (nv * G * G, ldk)
            assert torch.equal(pt[:, :P * P].cpu().float(), want), (ds, P, fmt)
            assert (pt[:, P * P:] == 0).all()
    words = ops.pack_events(evd, cfg["shape"])
    ptc, _, _ = ops.event2img(words, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="gray_f16", patch=16, ldk=256)
    pt, _, _ = ops.event2img(evd, fd, cfg["shape"], nv, cfg["count_non_zero"], cfg["background_mask"], out="gray_f16", patch=16, ldk=256)
    assert torch.equal(ptc, pt)


def test_events2frames_drop_in(cuda_dev):
    """Same call as the reference's datasets.vis.events2frames; uint8 [K,H,W,3] identical to the oracle."""
    for ds in ("n_caltech101", "n_cars"):
        cfg = SENSORS[ds]
        ev = synth_events(cfg["shape"], 45001, 9, "clustered")
        fr = events2frames(ev, split_method="event_count", convert_method="event_histogram", shape=cfg["shape"],
                           N=cfg["N"], grayscale=True, count_non_zero=cfg["count_non_zero"],
                           background_mask=cfg["background_mask"])
        ref = orc.events2frames(ev, cfg["shape"], cfg["N"], cfg["count_non_zero"], cfg["background_mask"])
        assert fr.dtype == np.uint8 and fr.shape == ref.shape and (fr == ref).all()
    with pytest.raises(NotImplementedError):
        events2frames(ev, "event_count", "voxel", shape=(100, 120), N=100)
    with pytest.raises(AssertionError):
        events2frames(ev, "time", "event_histogram", shape=(100, 120), N=100)


def test_error_flags(cuda_dev):
    ev = synth_events((100, 120), 500, 1)
    ev[7, 0], ev[7, 1] = 119, 100          # flat index == H*W
    with pytest.raises(ValueError):
        events2frames(ev, "event_count", "event_histogram", shape=(100, 120), N=30000, count_non_zero=True,
                      background_mask=False)
    ev[7, 0], ev[7, 1] = -3, 0
    with pytest.raises(ValueError):
        events2frames(ev, "event_count", "event_histogram", shape=(100, 120), N=30000)
    ev[7, 3] = 0                           # p == 0 events are never histogrammed, so their coordinates are not checked
    events2frames(ev, "event_count", "event_histogram", shape=(100, 120), N=30000)
    # x >= W aliases into the next row exactly like np.bincount(x + y*W) (vis.py:12)
    ev = synth_events((100, 120), 500, 2)
    ev[3, 0], ev[3, 1] = 125, 4
    fr = events2frames(ev, "event_count", "event_histogram", shape=(100, 120), N=30000)
    assert (fr == orc.events2frames(ev, (100, 120), 30000)).all()


def _overflow_stream(n_hot_pos, n_hot_neg, n_rest, seed):
    """N-ImageNet-shaped frame in which one pixel receives n_hot_pos positive and n_hot_neg negative events."""
    rng = np.random.default_rng(seed)
    E = n_hot_pos + n_hot_neg + n_rest
    ev = np.zeros((E, 4), np.float32)
    ev[:, 0], ev[:, 1] = 333, 257
    ev[:n_hot_pos, 3] = 1
    ev[n_hot_pos:n_hot_pos + n_hot_neg, 3] = -1
    ev[n_hot_pos + n_hot_neg:, 0] = rng.integers(0, 640, n_rest)
    ev[n_hot_pos + n_hot_neg:, 1] = rng.integers(0, 480, n_rest)
    ev[n_hot_pos + n_hot_neg:, 3] = np.where(rng.random(n_rest) < 0.5, -1.0, 1.0)
    ev = ev[rng.permutation(E)]
    ev[:, 2] = np.linspace(0, 0.05, E)
    return ev


@pytest.mark.parametrize("hot_pos,hot_neg,rest", [(70000, 0, 0), (66000, 0, 4000), (0, 67001, 2999), (65536, 1200, 3000),
                                                  (65535, 0, 4465), (65536, 0, 1)])
def test_counts_above_65535_are_exact(cuda_dev, hot_pos, hot_neg, rest):
    """The reference's histogram is int64 (datasets/vis.py:9-14); N-ImageNet's N = 70000 lets one 16-bit field of the packed
    bins wrap.  The wrapped pixel is recounted in 32 bits and the statistics are redone from the bins: counts, gray frame,
    resampled bytes and the float32 tensor equal the oracle's, and no status bit is raised."""
    ev = _overflow_stream(hot_pos, hot_neg, rest, seed=hot_pos % 97 + rest)
    N = 70000
    frames, _, _, K = ops.plan_frames([0, len(ev)], N, 2, compact=True)
    img, status, dbg = ops.event2img(torch.from_numpy(ev).to(cuda_dev), frames.to(cuda_dev), (480, 640), K, False, True,
                                     out="f32", debug=True)
    assert int(status.item()) == 0
    counts = orc.histogram(ev[:N], (480, 640))
    assert counts.max() >= 65535
    assert (dbg["counts"][0].cpu().numpy() == counts).all()
    gray, _, _ = orc.frame_from_counts(counts, False, True)
    assert (dbg["gray"][0].cpu().numpy() == gray).all()
    assert (dbg["u8"][0].cpu().numpy() == orc.resize_crop_224(gray)).all()
    oimg, _, _ = orc.event2img_sample(ev, (480, 640), N, 1, False, True)
    assert (img[0].cpu().numpy() == oimg[0]).all()
    # the non-debug kernel variant and the compact wire format take the same path
    img2, status2, _ = ops.event2img(torch.from_numpy(ev).to(cuda_dev), frames.to(cuda_dev), (480, 640), K, False, True, out="f32")
    words = ops.pack_events(torch.from_numpy(ev).to(cuda_dev), (480, 640))
    img3, status3, _ = ops.event2img(words, frames.to(cuda_dev), (480, 640), K, False, True, out="f32")
    assert torch.equal(img2, img) and torch.equal(img3, img) and int(status2.item()) == 0 and int(status3.item()) == 0


def test_band_exchange_kernel_cluster_sizes_and_rounds(cuda_dev):
    """event2img_big_kernel outside its home shape: EC_E2I_BIG=force routes sensors that fit one SM through 2-CTA clusters,
    256x512 takes a 4-CTA cluster, and EC_E2I_BIG_CAP shrinks the exchange round so that a frame needs several rounds.
    Every stage equals the oracle's (the environment is read per call)."""
    old = {k: os.environ.get(k) for k in ("EC_E2I_BIG", "EC_E2I_BIG_CAP")}
    try:
        for shape, N, cnz, bg, force, cap, want_cs in (((180, 240), 20000, False, True, True, None, 2),
                                                       ((180, 240), 20000, True, False, True, "1024", 2),
                                                       ((128, 128), 9000, False, True, True, "96", 2),
                                                       ((256, 512), 50000, False, True, False, None, 4),
                                                       ((480, 640), 70000, False, True, False, "2048", 8)):
            os.environ.pop("EC_E2I_BIG", None)
            os.environ.pop("EC_E2I_BIG_CAP", None)
            if force:
                os.environ["EC_E2I_BIG"] = "force"
            if cap:
                os.environ["EC_E2I_BIG_CAP"] = cap
            assert ops.event2img_geometry(shape)["cluster"] == want_cs, shape
            for kind in ("uniform", "clustered", "hotpixel"):
                ev = synth_events(shape, int(2.6 * N) + 17, 31, kind)
                img, dbg, K = _run_frames(ev, shape, N, cnz, bg, cuda_dev)
                i0, i1 = orc.split_event_count(len(ev), N)
                assert K == len(i0) == 3
                for k in range(K):
                    counts = orc.histogram(ev[i0[k]:i1[k]], shape)
                    assert (dbg["counts"][k].cpu().numpy() == counts).all(), (shape, kind, k)
                    gray, _, _ = orc.frame_from_counts(counts, cnz, bg)
                    assert (dbg["gray"][k].cpu().numpy() == gray).all(), (shape, kind, k)
                    assert (dbg["u8"][k].cpu().numpy() == orc.resize_crop_224(gray)).all(), (shape, kind, k)
                oimg, _, _ = orc.event2img_sample(ev, shape, N, K, cnz, bg)
                assert (img.cpu().numpy() == oimg[:K]).all(), (shape, kind)
                # non-debug variant, bf16 patch rows
                frames, _, _, K2 = ops.plan_frames([0, len(ev)], N, K, compact=True)
                pt, st, _ = ops.event2img(torch.from_numpy(ev).to(cuda_dev), frames.to(cuda_dev), shape, K2, cnz, bg, out="patch",
                                          patch=16, ldk=768)
                ref = img.to(torch.bfloat16).view(-1, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(-1, 768)
                assert torch.equal(pt, ref) and int(st.item()) == 0, (shape, kind)
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def test_full_size_properties(cuda_dev):
    """BASELINE-size batch: size-independent checks (event conservation, determinism, padding, permutation
    invariance of the histogram within a chunk)."""
    cfg = SENSORS["n_caltech101"]
    B = 32
    ev, off = synth_batch("n_caltech101", B, 1000)
    e2i = Event2Image(_qargs(cfg), cfg["shape"], cfg["max_n"])
    evd = torch.from_numpy(ev).to(cuda_dev)
    a = e2i(evd, off, debug=True, check=True)
    b = e2i(evd, off, debug=True, check=True)
    assert torch.equal(a["img"], b["img"])                                   # shared-memory atomics are order-free
    counts = a["debug"]["counts"]
    n = torch.tensor([f for f in np.frombuffer(a["frames"].numpy().tobytes(),
                                               dtype=[("s", "<i8"), ("n", "<i4"), ("o", "<i4")])["n"]])
    assert torch.equal(counts.sum((1, 2, 3)).cpu(), n.to(torch.int64))        # every event lands in exactly one bin
    assert a["valid_mask"].sum().item() == B * 5 and (a["img"][:, 5:] == 0).all()
    # shuffling events inside one chunk must not change its frame
    ev2 = ev.copy()
    rng = np.random.default_rng(0)
    ev2[:20000] = ev2[rng.permutation(20000)]
    c = e2i(torch.from_numpy(ev2).to(cuda_dev), off)
    assert torch.equal(c["img"], a["img"])


def test_center_and_flip_events_on_device(cuda_dev, golden_dir):
    """Row F1: ec_center_events / ec_flip_events against the reference's golden checksums and the oracle, on a packed
    batch (one CTA per sample), then frames of the centred stream against the oracle."""
    cases = json.load(open(os.path.join(golden_dir, "event_transforms_sha.json")))
    from tests.test_oracle_golden import _transform_case
    for c in cases:
        shape, ev = _transform_case(c)
        evs = [ev, ev[: len(ev) // 2].copy(), ev[len(ev) // 3:].copy()]
        off = np.concatenate([[0], np.cumsum([len(e) for e in evs])]).astype(np.int64)
        packed = torch.from_numpy(np.concatenate(evs)).to(cuda_dev)
        offd = torch.from_numpy(off).to(cuda_dev)
        for hf, tf, key in ((True, False, "hflip"), (False, True, "tflip"), (True, True, "htflip")):
            out = ops.flip_events(packed, offd, shape[1], hf, tf).cpu().numpy()
            assert sha(out[: len(ev)]) == c[key]
            for i, e in enumerate(evs):
                assert (out[off[i]:off[i + 1]] == orc.flip_events(e, shape[1], hf, tf)).all()
        cen = ops.center_events(packed.clone(), offd, shape).cpu().numpy()
        assert sha(cen[: len(ev)]) == c["centered"]
        for i, e in enumerate(evs):
            assert (cen[off[i]:off[i + 1]] == orc.center_events(e, shape)).all()
    # centred stream -> frames
    shape, ev = _transform_case(cases[0])
    cen = ops.center_events(torch.from_numpy(ev).to(cuda_dev), torch.tensor([0, len(ev)], device=cuda_dev), shape)
    fr = events2frames(cen, "event_count", "event_histogram", shape=shape, N=20000, grayscale=True)
    assert (fr == orc.events2frames(orc.center_events(ev, shape), shape, 20000)).all()


@pytest.mark.parametrize("shape,N", [((34, 34), 1500),      # N-MNIST: W % 4 != 0 -> SIMT kernel, one CTA
                                     ((128, 128), 5000),    # DVS128: tensor-core kernel
                                     ((240, 180), 20000),   # portrait: the crop falls on the rows
                                     ((64, 200), 6000),     # wide: both axes up-sampled, 175 columns cropped away
                                     ((260, 346), 25000)])  # DAVIS346: two-CTA cluster, W % 4 != 0
@pytest.mark.parametrize("flags", [(False, True), (True, False)])
def test_other_sensor_shapes_vs_oracle(cuda_dev, golden_dir, shape, N, flags):
    """Sensors outside BASELINE.json's three: every kernel variant (tensor-core, SIMT, cluster) against the oracle stage by
    stage, and the float32 tensor against the checksum of the UNMODIFIED reference's output (event2img_shapes_sha.json)."""
    cnz, bg = flags
    gold = {(tuple(c["shape"]), c["count_non_zero"], c["seed"]): c
            for c in json.load(open(os.path.join(golden_dir, "event2img_shapes_sha.json")))}
    for seed, kind in ((1, "uniform"), (2, "clustered"), (3, "hotpixel")):
        ev = synth_events(shape, int(2.6 * N) + 7, seed, kind)
        img, dbg, K = _run_frames(ev, shape, N, cnz, bg, cuda_dev)
        c = gold[(tuple(shape), cnz, seed)]
        assert sha(ev) == c["events"] and K == c["K"]
        assert sha(img.cpu().numpy()) == c["img"] and sha(dbg["u8"].cpu().numpy()) == c["u8"], (shape, kind)
        assert sha(dbg["counts"].cpu().numpy().astype(np.int32)) == c["counts"] and sha(dbg["gray"].cpu().numpy()) == c["frames"]
        i0, i1 = orc.split_event_count(len(ev), N)
        assert K == len(i0)
        img = img.cpu().numpy()
        for k in range(K):
            counts = orc.histogram(ev[i0[k]:i1[k]], shape)
            assert (dbg["counts"][k].cpu().numpy() == counts).all(), (shape, kind, k)
            gray, _, _ = orc.frame_from_counts(counts, cnz, bg)
            assert (dbg["gray"][k].cpu().numpy() == gray).all(), (shape, kind, k)
            u8 = orc.resize_crop_224(gray)
            assert (dbg["u8"][k].cpu().numpy() == u8).all(), (shape, kind, k)
            assert (img[k] == orc.normalize(u8)).all(), (shape, kind, k)


def test_simt_and_tensor_core_kernels_agree(cuda_dev):
    """EC_E2I_TC=0 forces the SIMT kernel (read once per process, hence the subprocess): same bytes as the default."""
    import subprocess
    import sys
    code = ("import hashlib, numpy as np, torch\n"
            "from eventclip_b200 import ops\n"
            "from eventclip_b200.synth import SENSORS, synth_batch\n"
            "for ds in ('n_caltech101', 'n_cars'):\n"
            "    cfg = SENSORS[ds]\n"
            "    ev, off = synth_batch(ds, 4, 77, kind='clustered')\n"
            "    frames, valid, chunks, nv = ops.plan_frames(off, cfg['N'], 3, compact=True)\n"
            "    dev = torch.device('cuda', 0)\n"
            "    for out, patch in (('f32', 0), ('patch', 16)):\n"
            "        img, st, _ = ops.event2img(torch.from_numpy(ev).to(dev), frames.to(dev), cfg['shape'], nv, cfg['count_non_zero'],\n"
            "                                   cfg['background_mask'], out=out, patch=patch)\n"
            "        torch.cuda.synchronize()\n"
            "        a = img.view(torch.int16) if img.dtype == torch.bfloat16 else img\n"
            "        print(ds, out, int(st.item()), hashlib.sha256(a.cpu().numpy().tobytes()).hexdigest())\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for tc in ("1", "0"):
        env = dict(os.environ, EC_E2I_TC=tc, PYTHONPATH=root)
        r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, cwd=root, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        outs.append(r.stdout)
    assert outs[0] == outs[1] and len(outs[0].splitlines()) == 4, outs
