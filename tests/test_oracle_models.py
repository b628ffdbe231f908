"""CPU: the classifier-head / CLIP oracles against golden outputs of the unmodified reference classifiers."""
import os

import numpy as np
import pytest
import torch

from oracle import clip_oracle, heads_oracle

ARCH = "ViT-tiny/32"


@pytest.fixture(scope="module")
def G(golden_dir):
    return np.load(os.path.join(golden_dir, "heads_golden.npz"))


def _close(a, b, tol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return np.abs(a - b).max() <= tol * max(1.0, np.abs(b).max())


@pytest.fixture(scope="module")
def feats(G):
    g = torch.Generator().manual_seed(77)
    imgs = torch.randn(6, 4, 3, 224, 224, generator=g)
    valid = torch.from_numpy(G["valid"])
    imgs = imgs * valid[:, :, None, None, None].float()
    with torch.no_grad():
        f = clip_oracle.build_clip(ARCH, seed=3).encode_image(imgs[valid])
    assert _close(f, G["img_feats"])
    return f


def test_zero_shot_head(G, feats):
    valid, text = torch.from_numpy(G["valid"]), torch.from_numpy(G["text"])
    for agg in ("mean", "sum"):
        o = heads_oracle.zs_head(feats, valid, text, 100.0, agg)
        for k in ("full_logits", "logits", "probs"):
            assert _close(o[k], G[f"zs_{agg}_{k}"]), (agg, k)


def test_few_shot_heads(G, feats):
    valid = torch.from_numpy(G["valid"])
    for tag in ("fs_trans", "fs_ident"):
        sd = {k[len(tag) + 4:]: torch.from_numpy(G[k]) for k in G.files if k.startswith(f"{tag}_sd_")}
        adapter = None
        if tag == "fs_trans":
            ap = {k[len("adapter."):]: v for k, v in sd.items() if k.startswith("adapter.")}
            adapter = lambda f, v: heads_oracle.adapter_forward(ap, f, v, num_heads=2, residual=0.8)
        o = heads_oracle.fs_head(feats, valid, sd["text_feats"], 100.0, "mean", adapter)
        for k in ("full_logits", "logits", "probs"):
            assert _close(o[k], G[f"{tag}_{k}"], 1e-4), (tag, k)


def test_fine_tune_lora_head(G):
    """LoRA-merged in_proj / out_proj weights reproduce the reference's LoraInjectedMHA forward."""
    valid = torch.from_numpy(G["valid"])
    g = torch.Generator().manual_seed(77)
    imgs = torch.randn(6, 4, 3, 224, 224, generator=g) * valid[:, :, None, None, None].float()
    clip = clip_oracle.build_clip(ARCH, seed=3)
    sd = {k[len("ft_lora_sd_"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("ft_lora_sd_")}
    d = clip.visual.transformer.width
    with torch.no_grad():
        for i, blk in enumerate(clip.visual.transformer.resblocks):
            pre = f"model.visual.transformer.resblocks.{i}.attn."
            lora = {n: sd[pre + f"in_proj_weight.{n}"] for n in
                    ("lora_up_q", "lora_down_q", "lora_up_k", "lora_down_k", "lora_up_v", "lora_down_v")}
            blk.attn.in_proj_weight.copy_(heads_oracle.lora_merged_in_proj(blk.attn.in_proj_weight, d, lora))
            blk.attn.out_proj.weight.add_(sd[pre + "out_proj.lora_up.weight"] @ sd[pre + "out_proj.lora_down.weight"])
        f = clip.encode_image(imgs[valid])
    o = heads_oracle.fs_head(f, valid, sd["text_feats"], 100.0, "mean", None)
    for k in ("full_logits", "logits", "probs"):
        assert _close(o[k], G[f"ft_lora_{k}"], 1e-4), k
    keys = set(G["ft_lora_keys"].tolist())
    assert "model.visual.transformer.resblocks.0.attn.in_proj_weight.lora_down_q" in keys
    assert "model.visual.transformer.resblocks.1.attn.out_proj.lora_up.weight" in keys
    assert "model.visual.transformer.resblocks.0.attn.out_proj.linear.weight" in keys


def lora_from_sd(sd, n_blocks, requires_grad=True):
    """{(block, 'q'|'k'|'v'|'o'): (up, down)} leaf tensors from a reference-named state dict, and the name of each."""
    lora, names = {}, {}
    for i in range(n_blocks):
        pre = f"model.visual.transformer.resblocks.{i}.attn."
        for n in "qkvo":
            ku, kd = (pre + "out_proj.lora_up.weight", pre + "out_proj.lora_down.weight") if n == "o" else \
                (pre + f"in_proj_weight.lora_up_{n}", pre + f"in_proj_weight.lora_down_{n}")
            if ku in sd:
                lora[(i, n)] = (sd[ku].clone().requires_grad_(requires_grad), sd[kd].clone().requires_grad_(requires_grad))
                names[(i, n)] = (ku, kd)
    return lora, names


@pytest.mark.parametrize("agg", ["mean", "sum"])
def test_fine_tune_step_gradients(G, golden_dir, agg):
    """oracle ft_train_loss + autograd == the unmodified reference's train-mode forward / calc_train_loss / backward."""
    T = np.load(os.path.join(golden_dir, "ft_train_golden.npz"))
    valid = torch.from_numpy(G["valid"])
    g = torch.Generator().manual_seed(77)
    imgs = torch.randn(6, 4, 3, 224, 224, generator=g) * valid[:, :, None, None, None].float()
    clip = clip_oracle.build_clip(ARCH, seed=3)
    sd = {k[len("ft_lora_sd_"):]: torch.from_numpy(G[k]) for k in G.files if k.startswith("ft_lora_sd_")}
    lora, names = lora_from_sd(sd, 2)
    text = sd["text_feats"].clone().requires_grad_(True)
    loss, o = heads_oracle.ft_train_loss(clip.visual, lora, imgs[valid], valid, text, torch.from_numpy(T["labels"]), 100.0, agg)
    loss.backward()
    assert abs(loss.item() - float(T[f"{agg}_loss"])) < 1e-4 * float(T[f"{agg}_loss"])
    assert _close(o["logits"].detach(), T[f"{agg}_logits"], 1e-4)
    assert _close(text.grad, T[f"{agg}_grad_text_feats"], 1e-4)
    for key, (up, down) in lora.items():
        assert _close(up.grad, T[f"{agg}_grad_{names[key][0]}"], 1e-4), key
        assert _close(down.grad, T[f"{agg}_grad_{names[key][1]}"], 1e-4), key


@pytest.mark.parametrize("tag", ["full", "subset"])
def test_fine_tune_step_gradients_whole_tower(G, golden_dir, tag):
    """Same for the non-LoRA trainable sets of clip_cls_ft.py:45-80: the whole tower ('full') and the union of the only_*
    switches ('subset': conv1, biases, LayerNorms, proj, class token)."""
    T = np.load(os.path.join(golden_dir, "ft_train_golden.npz"))
    valid = torch.from_numpy(G["valid"])
    g = torch.Generator().manual_seed(77)
    imgs = torch.randn(6, 4, 3, 224, 224, generator=g) * valid[:, :, None, None, None].float()
    clip = clip_oracle.build_clip(ARCH, seed=3)
    names = [k[len(tag) + 6 + len("model.visual."):] for k in T.files if k.startswith(f"{tag}_grad_model.visual.")]
    params = dict(clip.visual.named_parameters())
    for n, p in params.items():
        p.requires_grad_(n in names)
    text = torch.from_numpy(G["ft_lora_sd_text_feats"]).clone().requires_grad_(True)
    loss, o = heads_oracle.ft_train_loss(clip.visual, {}, imgs[valid], valid, text, torch.from_numpy(T["labels"]), 100.0, "mean")
    loss.backward()
    assert abs(loss.item() - float(T[f"{tag}_loss"])) < 1e-4 * float(T[f"{tag}_loss"])
    assert _close(text.grad, T[f"{tag}_grad_text_feats"], 1e-4)
    assert len(names) == (32 if tag == "full" else 23)
    for n in names:
        assert _close(params[n].grad, T[f"{tag}_grad_model.visual.{n}"], 2e-4), n
