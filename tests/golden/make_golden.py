"""Generates the committed golden fixtures by running the UNMODIFIED reference in the authoring container.

    python tests/golden/make_golden.py

Needs /root/reference (read-only) plus PIL / torchvision; nothing here runs on the GPU box.  The reference has no
tests or golden vectors of its own, so these files are the parity pins:
  split_cases.json      datasets/vis.py:55-72 split_event_count on edge-case (E, N) pairs
  gray_lut.npz          datasets/vis.py:27-39 uint8 value for (pos, neg, max) grids, both background_mask settings
  event2img_small.npz   full stage-by-stage arrays for a small sensor (events, counts, frames, resized u8)
  event2img_sha.json    sha256 of every stage for the three real sensor shapes x {uniform, clustered, hotpixel}
  event2img_shapes_sha.json  the same for five other sensor shapes x both flag settings x the three stream kinds
  heads_golden.npz      outputs of the reference ZS / FS / FT classifiers (models/clip_cls.py, clip_cls_ft.py,
                        adapter.py, lora.py) driven by the oracle CLIP image tower on seeded inputs
"""
import hashlib
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))

from oracle import ref_import, clip_oracle  # noqa: E402
from eventclip_b200.synth import SENSORS, synth_events  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def clip_preprocess():
    import torchvision.transforms as T
    return T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224),
                      lambda im: im.convert("RGB"), T.ToTensor(),
                      T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


def ref_counts(vis, ev, shape, N):
    """int64 [K,H,W,2] via the reference's own parse/split + np.bincount (vis.py:9-14)."""
    x, y, t, p = vis.parse_events(ev)
    i0, i1, _, _ = vis.split_event_count(t, N)
    H, W = shape
    out = []
    for a, b in zip(i0, i1):
        xx, yy, pp = x[a:b], y[a:b], p[a:b]
        pos = np.bincount(xx[pp > 0] + yy[pp > 0] * W, minlength=H * W).reshape(H, W)
        neg = np.bincount(xx[pp < 0] + yy[pp < 0] * W, minlength=H * W).reshape(H, W)
        out.append(np.stack([pos, neg], -1))
    return np.stack(out).astype(np.int64)


def ref_pipeline(vis, prep, ev, shape, q):
    from PIL import Image
    frames = vis.events2frames(ev.copy(), "event_count", "event_histogram", shape=shape, **q)
    tens = torch.stack([prep(Image.fromarray(f)) for f in frames]).numpy()
    # resized+cropped uint8 = invert ToTensor/Normalize is lossy; recompute the uint8 stage with PIL directly
    import torchvision.transforms as T
    rc = T.Compose([T.Resize(224, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(224)])
    u8 = np.stack([np.asarray(rc(Image.fromarray(f)))[:, :, 0] for f in frames])
    return frames, u8, tens


def make_split_cases(vis):
    cases = []
    for E, N in [(1, 10), (9, 10), (10, 10), (11, 10), (14, 10), (15, 10), (16, 10), (19, 10), (20, 10), (21, 10),
                 (25, 10), (26, 10), (30, 10), (99999, 20000), (100000, 20000), (110000, 20000), (110001, 20000),
                 (4000, 30000), (45000, 30000), (45001, 30000), (1000000, 70000), (1015000, 70000), (1015001, 70000),
                 (225000, 20000), (7, 7), (3, 2), (5, 3)]:
        t = np.arange(E, dtype=np.float64)
        i0, i1, _, _ = vis.split_event_count(t, N)
        cases.append(dict(E=E, N=N, idx0=[int(v) for v in i0], idx1=[int(v) for v in i1]))
    json.dump(cases, open(os.path.join(HERE, "split_cases.json"), "w"))
    print("split_cases", len(cases))


def make_gray_lut(vis):
    red = np.full(3, 127, np.uint8)
    out = {}
    rng = np.random.default_rng(5)
    for mask in (True, False):
        grids, rand = [], []
        for mx in range(1, 41):
            pp, nn = np.meshgrid(np.arange(mx + 1), np.arange(mx + 1))
            grids.append((mx, pp.ravel(), nn.ravel()))
        for mx in (63, 64, 127, 254, 255, 256, 300, 508, 511, 512, 1000, 4095, 20000, 65535):
            pos = rng.integers(0, mx + 1, 400)
            neg = rng.integers(0, mx + 1, 400)
            grids.append((mx, pos, neg))
        rows = []
        for mx, pos, neg in grids:
            n = len(pos) + 1
            P = np.concatenate([pos, [mx]]).astype(np.int64)
            Ng = np.concatenate([neg, [0]]).astype(np.int64)
            xs = np.concatenate([np.repeat(np.arange(n), P), np.repeat(np.arange(n), Ng)]).astype(np.int32)
            ps = np.concatenate([np.ones(P.sum()), -np.ones(Ng.sum())]).astype(np.int32)
            img = vis.make_event_histogram(xs, np.zeros_like(xs), ps, red, red, (1, n), thresh=0.,
                                           background_mask=mask)
            g = img[0, :-1, 0]
            rows.append(np.stack([np.full(len(pos), mx), pos, neg, g], 1))
        out["mask" if mask else "nomask"] = np.concatenate(rows).astype(np.int32)
    np.savez_compressed(os.path.join(HERE, "gray_lut.npz"), **out)
    print("gray_lut", {k: v.shape for k, v in out.items()})


def make_event2img(vis):
    prep = clip_preprocess()
    # (1) small sensor, everything stored
    small = {}
    for name, shape, E, N, cnz, bg, kind in [("a", (40, 56), 3300, 1000, False, True, "clustered"),
                                             ("b", (40, 56), 2600, 1000, True, False, "uniform"),
                                             ("c", (36, 48), 700, 1000, True, True, "hotpixel")]:
        ev = synth_events(shape, E, 900 + ord(name), kind)
        q = dict(N=N, grayscale=True, count_non_zero=cnz, background_mask=bg)
        frames, u8, tens = ref_pipeline(vis, prep, ev, shape, q)
        small[f"{name}_events"] = ev
        small[f"{name}_cfg"] = np.array([shape[0], shape[1], N, int(cnz), int(bg)])
        small[f"{name}_counts"] = ref_counts(vis, ev, shape, N).astype(np.int32)
        small[f"{name}_frames"] = frames[..., 0]
        assert (frames[..., 0] == frames[..., 1]).all() and (frames[..., 0] == frames[..., 2]).all()
        small[f"{name}_u8"] = u8
        small[f"{name}_img_sha"] = np.frombuffer(bytes.fromhex(sha(tens)), np.uint8)
    np.savez_compressed(os.path.join(HERE, "event2img_small.npz"), **small)
    # (2) real sensors: checksums per stage
    shas = []
    for ds, cfg in SENSORS.items():
        for kind in ("uniform", "clustered", "hotpixel"):
            E = cfg["E"] if ds != "n_imagenet" else 300000
            seed = 4242
            ev = synth_events(cfg["shape"], E, seed, kind, cfg["max_t"])
            q = dict(N=cfg["N"], grayscale=True, count_non_zero=cfg["count_non_zero"],
                     background_mask=cfg["background_mask"])
            frames, u8, tens = ref_pipeline(vis, prep, ev, cfg["shape"], q)
            counts = ref_counts(vis, ev, cfg["shape"], cfg["N"])
            shas.append(dict(dataset=ds, kind=kind, E=E, seed=seed, K=int(frames.shape[0]), events=sha(ev),
                             counts=sha(counts.astype(np.int32)), frames=sha(frames[..., 0]), u8=sha(u8),
                             img=sha(tens)))
            print(ds, kind, frames.shape)
    json.dump(shas, open(os.path.join(HERE, "event2img_sha.json"), "w"), indent=1)


OTHER_SHAPES = [((34, 34), 1500), ((128, 128), 5000), ((240, 180), 20000), ((64, 200), 6000), ((260, 346), 25000)]


def make_event2img_shapes(vis):
    """Sensors outside BASELINE.json's three (N-MNIST, DVS128, a portrait and a wide sensor, DAVIS346): per-stage checksums
    of the reference pipeline, both flag settings, the three stream kinds -- pins the oracle (and through it every event
    kernel variant) on shapes with other aspect ratios, crops on the other axis and W % 4 != 0."""
    prep = clip_preprocess()
    out = []
    for shape, N in OTHER_SHAPES:
        for cnz, bg in ((False, True), (True, False)):
            for seed, kind in ((1, "uniform"), (2, "clustered"), (3, "hotpixel")):
                ev = synth_events(shape, int(2.6 * N) + 7, seed, kind)
                q = dict(N=N, grayscale=True, count_non_zero=cnz, background_mask=bg)
                frames, u8, tens = ref_pipeline(vis, prep, ev, shape, q)
                counts = ref_counts(vis, ev, shape, N)
                out.append(dict(shape=list(shape), N=N, count_non_zero=cnz, background_mask=bg, seed=seed, kind=kind,
                                E=int(len(ev)), K=int(frames.shape[0]), events=sha(ev), counts=sha(counts.astype(np.int32)),
                                frames=sha(frames[..., 0]), u8=sha(u8), img=sha(tens)))
    json.dump(out, open(os.path.join(HERE, "event2img_shapes_sha.json"), "w"), indent=1)
    print("event2img shapes", len(out))


def make_heads():
    """Reference classifiers (unmodified, via nerv/clip stubs) on top of the oracle CLIP tower."""
    rm = ref_import.load_models()
    out = {}
    arch = "ViT-tiny/32"
    B, T, n_cls = 6, 4, 11
    g = torch.Generator().manual_seed(77)
    imgs = torch.randn(B, T, 3, 224, 224, generator=g)
    valid = torch.tensor([[1, 1, 1, 1], [1, 1, 0, 0], [1, 0, 0, 0], [1, 1, 1, 0], [1, 1, 1, 1], [1, 1, 0, 0]]).bool()
    imgs = imgs * valid[:, :, None, None, None].float()     # padded views are zeros (event2img.py:89-91)
    out["valid"] = valid.numpy()
    C = clip_oracle.ARCHS[arch][4]
    text = clip_oracle.synth_text_feats(n_cls, C, 5)
    out["text"] = text.numpy()
    names = [f"class_{i}" for i in range(n_cls)]
    data = dict(img=imgs, valid_mask=valid)

    def fresh_clip():
        return clip_oracle.build_clip(arch, seed=3)

    with torch.no_grad():
        feats = fresh_clip().encode_image(imgs[valid])
    out["img_feats"] = feats.numpy()
    # zero-shot; agg_func='max' raises in the reference itself (clip_cls.py:117 subtracts a [B,T] mask from
    # [B,T,n_cls] logits without unsqueezing), so only the two working aggregations have golden outputs
    for agg in ("mean", "sum"):
        m = rm.ZSCLIPClassifier(clip_dict=dict(clip_model=fresh_clip(), prompt="a {}", class_names=names, agg_func=agg))
        m.text_feats = text.clone()
        m.eval()
        with torch.no_grad():
            o = m(data)
        for k in ("full_logits", "logits", "probs"):
            out[f"zs_{agg}_{k}"] = o[k].numpy()
    # few-shot: joint adapter (text-trans, residual 0.8) and text-identity
    for tag, ad in (("fs_trans", dict(adapter_type="text-trans", in_dim=C, d_model=32, num_heads=2, ffn_dim=64,
                                      norm_first=True, num_layers=2, residual=0.8)),
                    ("fs_ident", dict(adapter_type="text-identity", residual=True))):
        torch.manual_seed(11)
        clipm = fresh_clip()
        cd = dict(clip_model=clipm, prompt="a {}", class_names=names, agg_func="mean")
        # FS builds its prompt parameter from cached text feats: pre-seed through a ZS pass-through
        orig = rm.FSCLIPClassifier._build_prompts

        def seeded(self, adapter_type, _t=text):
            self.text_feats = torch.nn.Parameter(_t.clone().float(), requires_grad=True)
            return adapter_type[5:]

        rm.FSCLIPClassifier._build_prompts = seeded
        try:
            m = rm.FSCLIPClassifier(adapter_dict=ad, clip_dict=cd, loss_dict=dict(use_logits_loss=True, use_probs_loss=False))
        finally:
            rm.FSCLIPClassifier._build_prompts = orig
        with torch.no_grad():      # make the learned parts non-trivial
            m.text_feats.add_(0.05 * torch.randn(m.text_feats.shape, generator=g))
        m.eval()
        with torch.no_grad():
            o = m(data)
        for k in ("full_logits", "logits", "probs"):
            out[f"{tag}_{k}"] = o[k].numpy()
        for k, v in m.state_dict().items():
            out[f"{tag}_sd_{k}"] = v.numpy()
    # fine-tune with LoRA qkvo-4, non-zero up factors
    torch.manual_seed(12)
    clipm = fresh_clip()
    cd = dict(clip_model=clipm, prompt="a {}", class_names=names, agg_func="mean", lora="qkvo-4", only_conv1=False,
              only_bias=False, only_ln=False)
    origf = rm.FTCLIPClassifier._build_prompts

    def seededf(self, adapter_type, _t=text):
        self.text_feats = torch.nn.Parameter(_t.clone().float(), requires_grad=True)
        return adapter_type[5:]

    rm.FTCLIPClassifier._build_prompts = seededf
    try:
        m = rm.FTCLIPClassifier(adapter_dict=dict(adapter_type="text-identity", residual=True), clip_dict=cd,
                                loss_dict=dict(use_logits_loss=True, use_probs_loss=False))
    finally:
        rm.FTCLIPClassifier._build_prompts = origf
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "lora_up" in n:
                p.copy_(0.05 * torch.randn(p.shape, generator=g))
    m.eval()
    with torch.no_grad():
        o = m(data)
    for k in ("full_logits", "logits", "probs"):
        out[f"ft_lora_{k}"] = o[k].numpy()
    sd = m.state_dict()
    out["ft_lora_keys"] = np.array(sorted(sd.keys()))
    for k, v in sd.items():
        if "lora" in k or k == "text_feats":
            out[f"ft_lora_sd_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "heads_golden.npz"), **out)
    print("heads_golden", len(out), "arrays")


def make_event_transforms():
    """datasets/utils.py center_events / flips on off-centre synthetic streams."""
    from oracle import event2img as orc
    U = ref_import.load_utils()
    out = []
    for seed, (shape, E) in enumerate([((180, 240), 5000), ((100, 120), 777), ((480, 640), 20000)]):
        ev = synth_events(shape, E, 50 + seed, "clustered")
        ev = ev[(ev[:, 0] < shape[1] * 0.6) & (ev[:, 1] < shape[0] * 0.7)].copy()
        ev[:, 2] += np.float32(0.37)
        ref = U.center_events(ev.copy(), resolution=shape)
        h = U.random_flip_events_along_x(ev.copy(), resolution=shape, p=1.)
        t = U.random_time_flip_events(ev.copy(), p=1.)
        ht = U.random_time_flip_events(h.copy(), p=1.)
        assert (ref == orc.center_events(ev, shape)).all()
        out.append(dict(shape=list(shape), E=E, seed=50 + seed, n=int(len(ev)), events=sha(ev), centered=sha(ref),
                        hflip=sha(h), tflip=sha(t), htflip=sha(ht)))
    json.dump(out, open(os.path.join(HERE, "event_transforms_sha.json"), "w"), indent=1)
    print("event_transforms", len(out))


def make_ft_train():
    """One fine-tune step of the UNMODIFIED reference FTCLIPClassifier (LoRA qkvo-4 + prompt-tuned text features) on the
    oracle CLIP tower: train-mode forward, calc_train_loss, autograd backward, one torch.optim.Adam step with the two
    learning rates of method.py:150-191.  Starts from the state stored in heads_golden.npz."""
    rm = ref_import.load_models()
    G = np.load(os.path.join(HERE, "heads_golden.npz"))
    arch, n_cls = "ViT-tiny/32", 11
    g = torch.Generator().manual_seed(77)
    valid = torch.from_numpy(G["valid"])
    imgs = torch.randn(6, 4, 3, 224, 224, generator=g) * valid[:, :, None, None, None].float()
    names = [f"class_{i}" for i in range(n_cls)]
    labels = torch.tensor([3, 0, 10, 7, 7, 1])
    out = {"labels": labels.numpy()}
    cases = [("mean", "mean", dict(lora="qkvo-4", only_conv1=False, only_bias=False, only_ln=False), True),
             ("sum", "sum", dict(lora="qkvo-4", only_conv1=False, only_bias=False, only_ln=False), True),
             # clip_cls_ft.py:78-80: no LoRA and no only_* switch = the whole image tower trains (configs/ftclip/*vitb16.py)
             ("full", "mean", dict(lora=-1, only_conv1=False, only_bias=False, only_ln=False), False),
             # clip_cls_ft.py:58-77: the union of the only_* subsets
             ("subset", "mean", dict(lora=-1, only_conv1=True, only_bias=True, only_ln=True, only_cls_fc=True,
                                     only_cls_token=True), False)]
    for tag, agg, flags, with_step in cases:
        clipm = clip_oracle.build_clip(arch, seed=3)
        cd = dict(clip_model=clipm, prompt="a {}", class_names=names, agg_func=agg, **flags)
        origf = rm.FTCLIPClassifier._build_prompts

        def seededf(self, adapter_type, _t=torch.from_numpy(G["text"])):
            self.text_feats = torch.nn.Parameter(_t.clone().float(), requires_grad=True)
            return adapter_type[5:]

        rm.FTCLIPClassifier._build_prompts = seededf
        try:
            m = rm.FTCLIPClassifier(adapter_dict=dict(adapter_type="text-identity", residual=True), clip_dict=cd,
                                    loss_dict=dict(use_logits_loss=True, use_probs_loss=False))
        finally:
            rm.FTCLIPClassifier._build_prompts = origf
        sd = {k[len("ft_lora_sd_"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("ft_lora_sd_")}
        if flags["lora"] == -1:
            sd = {"text_feats": sd["text_feats"]}      # same prompt-tuned text features, plain (un-injected) tower
        m.load_state_dict(sd, strict=False)      # the reference's override returns nothing
        for k, v in sd.items():
            assert torch.equal(m.state_dict()[k], v), k
        m.train()
        data = dict(img=imgs, valid_mask=valid, label=labels)
        o = m(data)
        loss = m.calc_train_loss(data, o)["ce_loss"]
        loss.backward()
        out[f"{tag}_loss"] = np.float32(loss.item())
        out[f"{tag}_logits"] = o["logits"].detach().numpy()
        named = [(n, p) for n, p in m.named_parameters() if p.requires_grad]
        for n, p in named:
            out[f"{tag}_grad_{n}"] = p.grad.numpy().copy()
        if with_step:
            # method.py:165-182: parameters outside model.visual get `lr`, those inside get `clip_lr`
            opt = torch.optim.Adam([{"params": [p for n, p in named if "model.visual" not in n], "lr": 1e-3},
                                    {"params": [p for n, p in named if "model.visual" in n], "lr": 5e-4}])
            opt.step()
            for n, p in named:
                out[f"{tag}_step1_{n}"] = p.detach().numpy().copy()
        print("ft_train", tag, "loss", loss.item(), "trainable", len(named), sum(p.numel() for _, p in named))
    np.savez_compressed(os.path.join(HERE, "ft_train_golden.npz"), **out)


def make_tta():
    """Row F4: executes the reference's OWN source lines -- the TTA aggregation / confidence filter of gen_data.py:141-164
    and the per-class top-k of gen_data.py:201-226 (lines 204-226, minus the path -> ground-truth lookups) -- on seeded
    synthetic probabilities, and stores inputs + outputs."""
    import textwrap
    import types
    src = open(os.path.join(ref_import.REF, "gen_data.py")).read().split("\n")
    block = textwrap.dedent("\n".join(src[140:164]))            # lines 141-164
    assert block.startswith("if tta:") and "sel_mask &= tta_mask" in block, block[:80]
    g = torch.Generator().manual_seed(123)
    B, n_cls = 64, 7
    logits4 = torch.randn(B, 4, n_cls, generator=g) * 2.0
    logits4[: B // 2] += 3.0 * torch.nn.functional.one_hot(torch.randint(0, n_cls, (B // 2,), generator=g), n_cls)[:, None].float()
    probs4 = logits4.softmax(-1)
    labels = torch.randint(0, n_cls, (B,), generator=g)
    out = {"probs4": probs4.numpy(), "labels": labels.numpy()}
    for ci, (thr, cons, minp) in enumerate([(-1.0, False, False), (0.5, True, False), (0.4, False, True), (0.6, True, True)]):
        ns = dict(torch=torch, tta=True, labels=labels, conf_thresh=thr,
                  args=types.SimpleNamespace(tta_consistent=cons, tta_min_prob=minp),
                  pred_probs=probs4.flatten(0, 1),
                  all_acc_meter=types.SimpleNamespace(update=lambda *a: None))
        exec(block, ns)
        out[f"case{ci}_cfg"] = np.array([thr, float(cons), float(minp)], np.float32)
        for k in ("probs", "max_probs", "pred_labels", "sel_mask"):
            out[f"case{ci}_{k}"] = ns[k].numpy()
    # per-class top-k: the reference walks a dict path -> {'cls', 'prob'}; run its loop on synthetic paths
    tk = textwrap.dedent("\n".join(src[203:226]))                # lines 204-226
    assert tk.startswith("for cls_name in class_names:"), tk[:60]
    sel = torch.from_numpy(out["case1_sel_mask"])
    names = [f"c{i}" for i in range(n_cls)]
    pred = {f"/d/{names[labels[i]]}/{i:03d}.npy": {"cls": names[int(out['case1_pred_labels'][i])], "prob": float(out["case1_max_probs"][i])}
            for i in range(B) if sel[i]}
    ns = dict(torch=torch, osp=os.path, class_names=names, pred_path2cls=pred, topk=3, is_nin=False,
              ev_dst=types.SimpleNamespace(folder2name={}), new_cnames=None, topk_pred_path2cls={}, sel_class_cnt={},
              sel_correct_class_cnt={})
    exec(tk, ns)
    keep = np.zeros(B, bool)
    for path in ns["topk_pred_path2cls"]:
        keep[int(os.path.basename(path)[:3])] = True
    out["topk3_keep"] = keep
    np.savez_compressed(os.path.join(HERE, "tta_golden.npz"), **out)
    print("tta_golden", {k: v.shape for k, v in out.items() if k.startswith("case1")}, "kept", int(keep.sum()))


def make_formats():
    """Row F2: the reference's own loaders (datasets/imagenet.py::load_event, NCaltech101._load_events) on synthetic files
    written in the two on-disk formats; stores SHA-256 of what they return."""
    import tempfile
    import textwrap
    import types
    # the dataset modules import nerv at the top; the two loaders are self-contained, so their source is executed alone
    src = open(os.path.join(ref_import.REF, "datasets", "imagenet.py")).read()
    ns = {"np": np}
    exec(src[src.index("def load_event"):src.index("class NImageNet")], ns)
    imagenet = types.SimpleNamespace(load_event=ns["load_event"])
    src = open(os.path.join(ref_import.REF, "datasets", "caltech.py")).read().split("\n")
    i0 = next(i for i, l in enumerate(src) if "def _load_events(event_path):" in l)
    ns2 = {"np": np}
    exec(textwrap.dedent("\n".join(src[i0:i0 + 3])), ns2)
    caltech = types.SimpleNamespace(NCaltech101=types.SimpleNamespace(_load_events=ns2["_load_events"]))
    out = []
    with tempfile.TemporaryDirectory() as td:
        for seed, (E, pol01) in enumerate([(5000, True), (777, False), (20000, True)]):
            rng = np.random.default_rng(900 + seed)
            rec = np.zeros(E, dtype=[("x", "<u2"), ("y", "<u2"), ("t", "<i8"), ("p", "?" if pol01 else "<i2")])
            rec["x"], rec["y"] = rng.integers(0, 640, E), rng.integers(0, 480, E)
            rec["t"] = np.sort(rng.integers(0, 55000, E))
            rec["p"] = rng.integers(0, 2, E) if pol01 else rng.choice([-1, 1], E)
            path = os.path.join(td, f"s{seed}.npz")
            np.savez(path, event_data=rec)
            ref = imagenet.load_event(path)
            out.append(dict(kind="npz", seed=900 + seed, E=E, pol01=pol01, dtype=str(ref.dtype), sha=sha(ref)))
        for seed, E in enumerate([3000, 1]):
            rng = np.random.default_rng(950 + seed)
            arr = np.stack([rng.integers(0, 240, E), rng.integers(0, 180, E), np.sort(rng.random(E)), rng.choice([-1, 1], E)], 1)
            path = os.path.join(td, f"c{seed}.npy")
            np.save(path, arr)                                                    # float64 on disk
            ref = caltech.NCaltech101._load_events(path)
            out.append(dict(kind="npy", seed=950 + seed, E=E, dtype=str(ref.dtype), sha=sha(ref)))
    json.dump(out, open(os.path.join(HERE, "formats_sha.json"), "w"), indent=1)
    print("formats", len(out))


if __name__ == "__main__":
    assert ref_import.available(), "/root/reference is required to (re)generate the golden fixtures"
    vis = ref_import.load_vis()
    make_split_cases(vis)
    make_gray_lut(vis)
    make_event2img(vis)
    make_event2img_shapes(vis)
    make_heads()
    make_event_transforms()
    make_ft_train()
    make_tta()
    make_formats()
    print("sizes:", {f: os.path.getsize(os.path.join(HERE, f)) for f in sorted(os.listdir(HERE)) if not f.endswith(".py")})
