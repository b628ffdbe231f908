"""CPU: host-side logic and the C-ABI surface (no device compute)."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

from eventclip_b200 import _lib, ops, clip
from eventclip_b200.synth import SENSORS, synth_batch, synth_events
from oracle import event2img as orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_abi_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "eventclip_b200.h")).read()
    declared = set(re.findall(r"EC_API\s+(?:const\s+)?\w+\s*\*?\s*(ec_\w+)\s*\(", hdr))
    assert len(declared) >= 18
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert built_lib.ec_version() >= 100
    # nothing but the C ABI is exported
    import subprocess
    syms = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in syms.splitlines() if " T " in l}
    assert exported == declared, exported ^ declared


def test_header_constants_match_binding():
    hdr = open(os.path.join(ROOT, "include", "eventclip_b200.h")).read()
    for name in ("EC_OK", "EC_ERR_ARG", "EC_ERR_CUDA", "EC_ERR_UNSUPPORTED", "EC_ERR_CAPACITY", "EC_STATUS_BAD_COORD",
                 "EC_STATUS_COUNT_OVERFLOW", "EC_FLAG_COUNT_NON_ZERO", "EC_FLAG_BACKGROUND_MASK", "EC_OUT_F32_NCHW",
                 "EC_OUT_BF16_NCHW", "EC_OUT_BF16_PATCH", "EC_EPI_BF16", "EC_EPI_BF16_QGELU", "EC_EPI_F32_RESADD",
                 "EC_EPI_F32", "EC_EPI_PATCH", "EC_EPI_F16_RESADD"):
        m = re.search(rf"#define {name}\s+(-?\d+)", hdr)
        assert m and int(m.group(1)) == getattr(_lib, name), name
    assert ctypes.sizeof(_lib.ECFrame) == 16


def test_event_kernel_selection_and_gemm_statistics_layout(built_lib):
    """Host-only entry points: which event kernel a sensor gets (one CTA + tensor-core passes when the bins fit one SM, rows
    are 4-byte aligned and the resize up-samples; clusters otherwise) and how many statistics slots a residual GEMM writes."""
    geo = {s: ops.event2img_geometry(s) for s in ((180, 240), (100, 120), (480, 640), (34, 34), (260, 346), (128, 128))}
    assert geo[(180, 240)] == dict(cluster=1, threads=1024, smem=217952)        # bins 172 800 + gray plane (188 rows + 32, 16-byte rounded)
    assert geo[(100, 120)]["cluster"] == 1 and geo[(100, 120)]["threads"] == 512 and geo[(100, 120)]["smem"] == 60992
    assert geo[(480, 640)]["cluster"] == 8 and geo[(260, 346)]["cluster"] == 2
    assert geo[(34, 34)]["cluster"] == 1 and geo[(128, 128)]["cluster"] == 1
    for s, g in geo.items():
        assert g["smem"] <= 227 * 1024 - 13 * 1024, (s, g)                      # leaves room for the kernels' static tables
    with pytest.raises(_lib.ECError):
        ops.event2img_geometry((4, 4))
    assert [ops.gemm_stats_parts(n) for n in (768, 1024, 512, 384, 1280)] == [6, 8, 4, 6, 10]


def _frames_np(frames):
    return np.frombuffer(frames.numpy().tobytes(), dtype=[("s", "<i8"), ("n", "<i4"), ("o", "<i4")])


def test_plan_frames_matches_reference_split(golden_dir, built_lib):
    cases = json.load(open(os.path.join(golden_dir, "split_cases.json")))
    for c in cases:
        K = len(c["idx0"])
        fr, valid, chunks, nv = ops.plan_frames([0, c["E"]], c["N"], K + 2, compact=True)
        rec = _frames_np(fr)
        assert chunks.tolist() == [K] and nv == K and valid[0].tolist() == [True] * K + [False] * 2
        assert [int(r["s"]) for r in rec] == c["idx0"]
        assert [int(r["s"] + r["n"]) for r in rec] == c["idx1"]
        assert [int(r["o"]) for r in rec] == list(range(K))


def test_plan_frames_padding_selection_and_errors(built_lib):
    off = [0, 100000, 104000, 154001]
    fr, valid, chunks, nv = ops.plan_frames(off, 20000, 4)
    rec = _frames_np(fr)
    assert len(rec) == 12 and chunks.tolist() == [5, 1, 3] and nv == 4 + 1 + 3
    assert valid.tolist() == [[True] * 4, [True, False, False, False], [True, True, True, False]]
    assert [int(r["o"]) for r in rec] == list(range(12))
    assert rec[5]["n"] == 0 and rec[11]["n"] == 0                       # padded view slots
    assert rec[10]["s"] == 104000 + 30001 and rec[10]["n"] == 20000       # overlapping tail chunk (vis.py:67-69)
    sel = np.tile(np.arange(4, dtype=np.int32), (3, 1))
    sel[0] = [4, 2, 0, 1]
    fr2, _, _, _ = ops.plan_frames(off, 20000, 4, sel=sel, compact=True)
    rec2 = _frames_np(fr2)
    assert [int(r["s"]) for r in rec2[:4]] == [80000, 40000, 0, 20000] and len(rec2) == 8
    sel[0, 0] = 5
    with pytest.raises(_lib.ECError):
        ops.plan_frames(off, 20000, 4, sel=sel)
    with pytest.raises(AssertionError):      # an empty stream never reaches the reference (caltech.py:181-182)
        ops.plan_frames([0, 10, 10], 20000, 4)


def test_view_slot_rule():
    from eventclip_b200.datasets.event2img import Event2Image
    for ds, want in (("n_caltech101", 10), ("n_cars", 1), ("n_imagenet", 2)):
        cfg = SENSORS[ds]
        q = dict(max_imgs=10, N=cfg["N"], split_method="event_count", convert_method="event_histogram", grayscale=True,
                 count_non_zero=cfg["count_non_zero"], background_mask=cfg["background_mask"])
        e = Event2Image(q, cfg["shape"], cfg["max_n"])
        assert e.max_imgs == want == orc.max_imgs(cfg["max_n"], cfg["N"], 10)
        q["max_imgs"] = 2
        assert Event2Image(q, cfg["shape"], cfg["max_n"]).max_imgs == min(want, 2)


def test_synth_streams_are_valid():
    ev, off = synth_batch("n_cars", 3, 10)
    assert ev.dtype == np.float32 and ev.shape == (12000, 4) and off.tolist() == [0, 4000, 8000, 12000]
    for kind in ("uniform", "clustered", "hotpixel"):
        e = synth_events((180, 240), 5000, 1, kind)
        assert (e[:, 0] >= 0).all() and (e[:, 0] < 240).all() and (e[:, 1] >= 0).all() and (e[:, 1] < 180).all()
        assert (np.diff(e[:, 2]) >= 0).all() and set(np.unique(e[:, 3])) == {-1.0, 1.0}
        assert (e[:, :2] == np.floor(e[:, :2])).all()


def test_product_fails_loudly_without_gpu(built_lib):
    """No CPU fallback: CPU tensors are rejected, and without a CUDA device the device check raises."""
    ev = torch.zeros(8, 4)
    fr = torch.zeros(1, 16, dtype=torch.uint8)
    with pytest.raises(_lib.ECError):
        ops.event2img(ev, fr, (100, 120), 1)
    with pytest.raises(_lib.ECError):
        ops.gemm_bf16(torch.zeros(128, 64, dtype=torch.bfloat16), torch.zeros(128, 64, dtype=torch.bfloat16))
    if not torch.cuda.is_available():
        with pytest.raises(_lib.ECError):
            _lib.require_device(0)
        m = clip.CLIP("ViT-tiny/32")
        with pytest.raises(_lib.ECError):
            m.encode_image(torch.zeros(1, 3, 224, 224))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "eventclip_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, f


def test_clip_state_dict_is_interchangeable_with_oracle():
    from oracle import clip_oracle
    a = clip_oracle.build_clip("ViT-tiny/16", seed=4)
    b = clip.CLIP("ViT-tiny/16")
    missing = b.load_state_dict(a.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    # the product's own seeded init draws the same weights as the oracle's for the same seed
    c = clip.init_weights_(clip.CLIP("ViT-tiny/16"), seed=4)
    for (k1, v1), (k2, v2) in zip(a.state_dict().items(), c.state_dict().items()):
        assert k1 == k2 and torch.equal(v1, v2), k1
    assert clip.flops_per_image("ViT-B/32") == pytest.approx(8.818e9, rel=1e-3)
    assert clip.flops_per_image("ViT-B/16") == pytest.approx(35.127e9, rel=1e-3)
    assert clip.flops_per_image("ViT-L/14") == pytest.approx(162.026e9, rel=1e-3)


def test_classifier_construction_and_state_dict_names(golden_dir):
    """Constructor dict schemas, parameter names and the CLIP-stripping state_dict follow the reference."""
    from eventclip_b200.models import build_model

    class P:
        pass

    G = np.load(os.path.join(golden_dir, "heads_golden.npz"))
    text = torch.from_numpy(G["text"])
    names = [f"class_{i}" for i in range(11)]
    p = P()
    p.model = "FSCLIP"
    p.clip_dict = dict(clip_model=clip.CLIP("ViT-tiny/32"), prompt="a {}", class_names=names, agg_func="mean",
                       text_feats=text)
    p.adapter_dict = dict(adapter_type="text-trans", in_dim=64, d_model=32, num_heads=2, ffn_dim=64, norm_first=True,
                          num_layers=2, residual=0.8)
    p.loss_dict = dict(use_logits_loss=True, use_probs_loss=False)
    m = build_model(p)
    want = {k[len("fs_trans_sd_"):] for k in G.files if k.startswith("fs_trans_sd_")}
    assert set(m.state_dict().keys()) == want
    m.load_state_dict({k: torch.from_numpy(G["fs_trans_sd_" + k]) for k in want})
    assert m.train().model.training is False and m.adapter.training is True

    p.model = "FTCLIP"
    p.clip_dict = dict(clip_model=clip.CLIP("ViT-tiny/32"), prompt="a {}", class_names=names, agg_func="mean",
                       text_feats=text, lora="qkvo-4", only_conv1=False, only_bias=False, only_ln=False)
    p.adapter_dict = dict(adapter_type="text-identity", residual=True)
    m = build_model(p)
    assert set(m.state_dict().keys()) == set(G["ft_lora_keys"].tolist())
    trainable = {n for n, q in m.named_parameters() if q.requires_grad}
    assert all(("lora_" in n) or n == "text_feats" for n in trainable) and len(trainable) == 2 * 8 + 1
    with pytest.raises(AssertionError):
        p.loss_dict = dict(use_logits_loss=True, use_probs_loss=True)
        p.model = "FSCLIP"
        p.adapter_dict = dict(adapter_type="identity", residual=True)
        build_model(p)
    p.model = "nope"
    with pytest.raises(NotImplementedError):
        build_model(p)


def test_layernorm_folding_operands_reproduce_layernorm_plus_linear():
    """pack_blocks_ln (host, pack time): with Wg = fp16(gamma * W_eff) row-centred and c = W_eff beta + b,
    rstd * (x Wg^T) + c equals LayerNorm(x) W_eff^T + b -- also with non-zero LoRA factors merged (models/lora.py:138-149)."""
    from eventclip_b200.models.lora import inject_trainable_lora
    torch.manual_seed(3)
    model = clip.CLIP("ViT-tiny/32")
    with torch.no_grad():
        for blk in model.visual.transformer.resblocks:
            for ln in (blk.ln_1, blk.ln_2):
                ln.weight.copy_(torch.rand_like(ln.weight) + 0.5)
                ln.bias.copy_(torch.randn_like(ln.bias) * 0.2)
    for with_lora in (False, True):
        if with_lora:
            inject_trainable_lora(model.visual, "qkvo-4")
            with torch.no_grad():
                for n, p in model.visual.named_parameters():
                    if "lora_up" in n:
                        p.copy_(torch.randn_like(p) * 0.05)
        d = model.visual.width
        packed = clip.pack_blocks_ln(model.visual.transformer.resblocks, d, torch.device("cpu"))
        x = torch.randn(37, d) * 2 + 0.7
        mean, var = x.mean(1, keepdim=True), x.var(1, unbiased=False, keepdim=True)
        rstd = torch.rsqrt(var + 1e-5)
        for blk, e in zip(model.visual.transformer.resblocks, packed):
            W, lora_in, ib, _, _, _ = clip._attn_weights(blk.attn)
            W = W.detach().float().clone()
            if lora_in is not None:
                for j, nme in enumerate("qkv"):
                    W[j * d:(j + 1) * d] += getattr(lora_in, f"lora_up_{nme}").detach() @ getattr(lora_in, f"lora_down_{nme}").detach()
            for name, Wf, bf_, ln in (("in", W, ib.detach(), blk.ln_1),
                                      ("fc", blk.mlp.c_fc.weight.detach(), blk.mlp.c_fc.bias.detach(), blk.ln_2)):
                want = torch.nn.functional.layer_norm(x, (d,), ln.weight, ln.bias, 1e-5) @ Wf.t() + bf_
                got = rstd * (x @ e["wg_" + name].float().t()) + e["c_" + name]
                assert e["wg_" + name].dtype == torch.float16
                assert e["wg_" + name].float().sum(1).abs().max() < 2e-3                  # centred rows (up to one fp16 rounding)
                assert (got - want).norm() / want.norm() < 2e-3, (with_lora, name)


@pytest.mark.parametrize("arch", ["ViT-tiny/16", "ViT-tiny/32", "ViT-L/14"])
def test_gray_folded_conv1_reproduces_conv1_on_the_normalised_channels(arch):
    """clip.VisionTransformer.packed_gray (CPU): event frames are grayscale, so CLIP's preprocess yields three channels that are
    affine in ONE byte g; conv1 folded onto the plane g / 128 (+ the constant that joins the positional embedding) must equal
    conv1 on ToTensor + Normalize of the gray image, for every byte value."""
    m = clip.init_weights_(clip.CLIP(arch), seed=5)
    vis = m.visual
    vis.operand_dtype = torch.float32                    # the algebra, without the 16-bit rounding of the packed operand
    pg = vis.packed_gray()
    P, d = vis.patch_size, vis.width
    g = torch.Generator().manual_seed(1)
    byte = torch.randint(0, 256, (3, 1, 224, 224), generator=g).float()
    byte[0, 0, :16, :16] = torch.arange(256.).view(16, 16)          # every byte value occurs
    mean = torch.tensor(clip.CLIP_MEAN).view(1, 3, 1, 1)
    std = torch.tensor(clip.CLIP_STD).view(1, 3, 1, 1)
    x = (byte / 255.0 - mean) / std                                  # [3, 3, 224, 224]: the reference's tensor
    ref = torch.nn.functional.conv2d(x.double(), vis.conv1.weight.double(), stride=P)          # [3, d, G, G]
    G = 224 // P
    rows = (byte / 128.0).view(3, G, P, G, P).permute(0, 1, 3, 2, 4).reshape(3 * G * G, P * P).double()
    got = rows @ pg["conv1"][:, :P * P].double().t()                 # [3 G G, d]
    got = got + (pg["pos"][1:].double() - vis.positional_embedding[1:].double()).repeat(3, 1)   # the folded constant
    want = ref.permute(0, 2, 3, 1).reshape(3 * G * G, d)
    assert pg["conv1"].shape == (d, vis.k_gray) and (pg["conv1"][:, P * P:] == 0).all()
    assert (got - want).abs().max() < 2e-5 * want.abs().max().clamp_min(1.0)
