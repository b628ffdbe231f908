"""oracle/heads_oracle.py -- TEST INFRASTRUCTURE ONLY.

Functional fp32 restatement of the reference classifier heads (rows A8-A13 of SURVEY.md section 8):
  zero-shot head     models/clip_cls.py:131-162
  logits aggregation models/clip_cls.py:104-121,  probs aggregation 123-129
  few-shot head      models/clip_cls.py:308-350  (adapter models/adapter.py:82-105, blend 22-25)
  fine-tune head     models/clip_cls_ft.py:214-256 (adapter call skipped at :228)
  LoRA merge         models/lora.py:138-149 (q/k/v), 49-52 (out_proj); init 8-11
Checked against the UNMODIFIED reference modules by tests/golden/make_golden.py.
"""
import torch
import torch.nn.functional as F


def aggregate_logits(full_logits, valid, agg):
    v = valid.float()
    if agg == "sum":
        return full_logits.sum(1)
    if agg == "mean":
        return full_logits.sum(1) / v.sum(1, keepdim=True)
    if agg == "max":
        return (full_logits - (1.0 - v)[..., None] * 1e6).max(1)[0]
    raise NotImplementedError(agg)


def aggregate_probs(full_logits, valid):
    v = valid.float()
    p = full_logits.softmax(-1) * v[..., None]
    return p.sum(1) / v.sum(1, keepdim=True)


def zs_head(img_feats, valid, text_feats, scale, agg):
    """img_feats [Nv,C] of the valid views in row-major (b,t) order; NOT normalised (clip_cls.py:148)."""
    B, T = valid.shape
    logits = scale * img_feats @ text_feats.T
    full = torch.zeros(B, T, text_feats.shape[0], dtype=logits.dtype)
    full[valid] = logits
    return dict(full_logits=full, valid_masks=valid, logits=aggregate_logits(full, valid, agg),
                probs=aggregate_probs(full, valid))


def adapter_forward(p, feats, valid, num_heads, residual, eps=1e-5):
    """TransformerAdapter (adapter.py:82-105) with norm_first encoder layers, eval mode (no dropout).
    p: state dict of the adapter (in_proj.*, transformer_encoder.layers.i.*, out_proj.*)."""
    x = F.linear(feats, p["in_proj.weight"], p["in_proj.bias"])
    B, T, D = x.shape
    hd = D // num_heads
    i = 0
    while f"transformer_encoder.layers.{i}.norm1.weight" in p:
        q = f"transformer_encoder.layers.{i}."
        h = F.layer_norm(x, (D,), p[q + "norm1.weight"], p[q + "norm1.bias"], eps)
        qkv = F.linear(h, p[q + "self_attn.in_proj_weight"], p[q + "self_attn.in_proj_bias"])
        qq, kk, vv = [t.view(B, T, num_heads, hd).transpose(1, 2) for t in qkv.chunk(3, -1)]
        s = qq @ kk.transpose(-1, -2) / hd ** 0.5
        s = s.masked_fill(~valid[:, None, None, :], float("-inf"))
        a = (s.softmax(-1) @ vv).transpose(1, 2).reshape(B, T, D)
        x = x + F.linear(a, p[q + "self_attn.out_proj.weight"], p[q + "self_attn.out_proj.bias"])
        h = F.layer_norm(x, (D,), p[q + "norm2.weight"], p[q + "norm2.bias"], eps)
        h = F.linear(F.relu(F.linear(h, p[q + "linear1.weight"], p[q + "linear1.bias"])),
                     p[q + "linear2.weight"], p[q + "linear2.bias"])
        x = x + h
        i += 1
    x = F.linear(x, p["out_proj.weight"], p["out_proj.bias"])
    return feats * residual + x * (1.0 - residual)


def fs_head(img_feats, valid, text_param, scale, agg, adapter=None, normalize_text=True):
    """Few-shot / fine-tune head.  adapter: None (identity / FT) or a callable (feats,valid)->feats."""
    B, T = valid.shape
    C = img_feats.shape[-1]
    full = torch.zeros(B, T, C, dtype=img_feats.dtype, device=img_feats.device)
    full[valid] = img_feats
    if adapter is not None:
        full = adapter(full, valid)
    full = F.normalize(full, p=2, dim=-1) * valid.float()[..., None]
    text = F.normalize(text_param, p=2, dim=-1) if normalize_text else text_param
    fl = scale * full @ text.T
    return dict(full_logits=fl, valid_masks=valid, logits=aggregate_logits(fl, valid, agg),
                probs=aggregate_probs(fl, valid))


def lora_merged_in_proj(W, d, lora):
    """lora.py:138-149: lora = dict with lora_up_{q,k,v} [d,r], lora_down_{q,k,v} [r,d] (k optional)."""
    parts = []
    for j, n in enumerate("qkv"):
        w = W[j * d:(j + 1) * d]
        if f"lora_up_{n}" in lora:
            w = w + lora[f"lora_up_{n}"] @ lora[f"lora_down_{n}"]
        parts.append(w)
    return torch.cat(parts, 0)


def ft_train_loss(visual, lora, imgs, valid, text_param, labels, scale, agg):
    """Loss of one fine-tune step, differentiable w.r.t. the LoRA factors and text_param (autograd is the backward the
    reference uses): models/clip_cls_ft.py:214-269 with models/lora.py's merged weights.
      visual  oracle VisionTransformer (fp32);  lora {(block, 'q'|'k'|'v'|'o'): (up [rows,r], down [r,d])}
      imgs    [n_valid, 3, 224, 224] views of the valid slots;  valid bool [B,T]
    Pinned against the unmodified reference by tests/golden/ft_train_golden.npz (tests/test_oracle_models.py)."""
    d = visual.transformer.width
    over = {}
    for i, blk in enumerate(visual.transformer.resblocks):
        qkv = {n: lora[(i, n)] for n in "qkv" if (i, n) in lora}
        if qkv:
            W = blk.attn.in_proj_weight.detach()
            over[f"transformer.resblocks.{i}.attn.in_proj_weight"] = torch.cat(
                [W[j * d:(j + 1) * d] + (qkv[n][0] @ qkv[n][1] if n in qkv else 0) for j, n in enumerate("qkv")], 0)
        if (i, "o") in lora:
            up, down = lora[(i, "o")]
            over[f"transformer.resblocks.{i}.attn.out_proj.weight"] = blk.attn.out_proj.weight.detach() + up @ down
    feats = torch.func.functional_call(visual, over, (imgs,))
    o = fs_head(feats, valid, text_param, scale, agg)
    return F.cross_entropy(o["logits"], labels), o
