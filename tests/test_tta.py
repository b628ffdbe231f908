"""Row F4 (flip-TTA of the pseudo-label generator).  CPU: the selection rules -- oracle restatement and the product's
bookkeeping -- against outputs of the reference's own source lines (tests/golden/tta_golden.npz).  GPU: the 4-variant
device batch bit-exact against the oracle's flips, and the TTA forward against oracle frames + fp32 oracle CLIP."""
import os

import numpy as np
import pytest
import torch

from oracle import tta_oracle


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "tta_golden.npz"))


@pytest.mark.parametrize("case", range(4))
def test_selection_rules_vs_reference_lines(G, case):
    from eventclip_b200 import tta
    thr, cons, minp = G[f"case{case}_cfg"].tolist()
    o = tta_oracle.tta_select(G["probs4"], thr, bool(cons), bool(minp))
    p = tta.tta_select(torch.from_numpy(G["probs4"]), thr, bool(cons), bool(minp))
    for k in ("probs", "max_probs", "pred_labels", "sel_mask"):
        ref = G[f"case{case}_{k}"]
        if ref.dtype.kind == "f":
            assert np.allclose(o[k], ref, rtol=0, atol=1e-7), k
            assert np.allclose(p[k].numpy(), ref, rtol=0, atol=1e-7), k
        else:
            assert np.array_equal(o[k], ref), k
            assert np.array_equal(p[k].numpy(), ref), k
    assert 0 < G[f"case{case}_sel_mask"].sum() <= 64


def test_topk_per_class_vs_reference_lines(G):
    from eventclip_b200 import tta
    keep_o = tta_oracle.topk_per_class(G["case1_pred_labels"], G["case1_max_probs"], G["case1_sel_mask"], 7, 3)
    keep_p = tta.topk_per_class(torch.from_numpy(G["case1_pred_labels"]), torch.from_numpy(G["case1_max_probs"]),
                                torch.from_numpy(G["case1_sel_mask"]), 7, 3)
    assert np.array_equal(keep_o, G["topk3_keep"]) and np.array_equal(keep_p.numpy(), G["topk3_keep"])


@pytest.mark.gpu
def test_tta_forward_vs_oracle(cuda_dev):
    from eventclip_b200 import clip, tta
    from eventclip_b200.models import ZSCLIPClassifier
    from eventclip_b200.synth import SENSORS, synth_batch
    from oracle import clip_oracle, heads_oracle
    from oracle import event2img as orc
    ds, arch, B = "n_caltech101", "ViT-B/32", 3
    cfg = SENSORS[ds]
    q = dict(max_imgs=10, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True,
             count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])
    ev, off = synth_batch(ds, B, 700, kind="clustered", E=45000)
    ev[:, 0] = np.minimum(ev[:, 0], cfg["shape"][1] * 0.7).astype(np.float32)      # off-centre content: the h-flip matters
    oracle = clip_oracle.build_clip(arch, seed=31)
    text = clip_oracle.synth_text_feats(cfg["n_cls"], oracle.visual.output_dim, 6)
    model = clip.CLIP(arch)
    model.load_state_dict(oracle.state_dict())
    zs = ZSCLIPClassifier(clip_dict=dict(clip_model=model.to(cuda_dev).eval(), prompt="a {}", class_names=None,
                                         agg_func="mean", text_feats=text)).to(cuda_dev).eval()
    zs.attach_event_frontend(q, cfg["shape"], cfg["max_n"])
    evd = torch.from_numpy(ev).to(cuda_dev)
    ev4, off4 = tta.tta_events(evd, off, cfg["shape"][1])
    ev4 = ev4.cpu().numpy()
    T = zs.event_frontend.max_imgs
    ref_probs = []
    for b in range(B):
        variants = tta_oracle.tta_variants(ev[off[b]:off[b + 1]], cfg["shape"])
        imgs, valids = [], []
        for v, e in enumerate(variants):
            got = ev4[off4[v * B + b]:off4[v * B + b + 1]]
            assert np.array_equal(got, e), (b, v)                       # device variants are bit-exact
            im, va, _ = orc.event2img_sample(e, cfg["shape"], cfg["N"], T, cfg["count_non_zero"], cfg["background_mask"])
            imgs.append(im)
            valids.append(va)
        imgs, valid = torch.from_numpy(np.stack(imgs)), torch.from_numpy(np.stack(valids))
        with torch.no_grad():
            feats = oracle.encode_image(imgs[valid])
        ref_probs.append(heads_oracle.zs_head(feats, valid, text, 100.0, "mean")["probs"])
    ref = torch.stack(ref_probs)                                         # [B, 4, n_cls]
    out = tta.tta_forward(zs, evd, off)
    assert out["probs"].shape == ref.shape
    assert np.abs(out["probs"].cpu().numpy() - ref.numpy()).max() < 5e-2
    a = tta.tta_select(out["probs"].cpu(), 0.0, True, True)
    o = tta_oracle.tta_select(ref.numpy(), 0.0, True, True)
    assert np.abs(a["probs"].numpy() - o["probs"]).max() < 5e-2
    # the flips change the frames: the four variants are not all identical
    assert (ref[:, 0] - ref[:, 1]).abs().max() > 0 or (ref[:, 0] - ref[:, 2]).abs().max() > 0
