"""oracle/clip_oracle.py -- TEST INFRASTRUCTURE ONLY.

Plain-PyTorch fp32 restatement of the third-party CLIP pieces the reference calls:
openai/CLIP `clip/model.py` (git HEAD, unpinned -- docs/install.md:34, environment.yml:87 `clip==1.0`) is not
vendored under /root/reference, so its published image-encoder algorithm is restated here with the same
parameter names, and parity is anchored on the reference's call sites:
  encode_image   models/clip_cls.py:101, models/clip_cls_ft.py:180
  logit_scale    models/clip_cls.py:44
  visual.*       models/clip_cls_ft.py:53-80, models/lora.py:388-402
tests/golden/make_golden.py cross-checks this file against HF transformers' independent CLIP vision tower
(weights remapped) and drives the UNMODIFIED reference classifiers with it to produce the golden logits.
"""
import math
from collections import OrderedDict

import torch
import torch.nn as nn

ARCHS = {
    # name: (patch, width, layers, heads, embed_dim)
    "ViT-B/32": (32, 768, 12, 12, 512),
    "ViT-B/16": (16, 768, 12, 12, 512),
    "ViT-L/14": (14, 1024, 24, 16, 768),
    # tiny shapes for fast CPU tests (not CLIP releases)
    "ViT-tiny/32": (32, 128, 2, 2, 64),
    "ViT-tiny/16": (16, 128, 2, 2, 64),
}


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class Block(nn.Module):
    """x += MHA(ln_1(x)); x += c_proj(QuickGELU(c_fc(ln_2(x)))) on [L, N, d] input."""

    def __init__(self, d, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(OrderedDict(c_fc=nn.Linear(d, 4 * d), gelu=QuickGELU(), c_proj=nn.Linear(4 * d, d)))
        self.ln_2 = nn.LayerNorm(d)

    def forward(self, x):
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class Tower(nn.Module):
    def __init__(self, d, layers, heads):
        super().__init__()
        self.width, self.layers = d, layers
        self.resblocks = nn.Sequential(*[Block(d, heads) for _ in range(layers)])

    def forward(self, x):
        return self.resblocks(x)


class VisionTransformer(nn.Module):
    def __init__(self, patch, d, layers, heads, out_dim, res=224):
        super().__init__()
        self.input_resolution, self.output_dim, self.patch = res, out_dim, patch
        self.conv1 = nn.Conv2d(3, d, patch, patch, bias=False)
        g = res // patch
        s = d ** -0.5
        self.class_embedding = nn.Parameter(s * torch.randn(d))
        self.positional_embedding = nn.Parameter(s * torch.randn(g * g + 1, d))
        self.ln_pre = nn.LayerNorm(d)
        self.transformer = Tower(d, layers, heads)
        self.ln_post = nn.LayerNorm(d)
        self.proj = nn.Parameter(s * torch.randn(d, out_dim))

    def forward(self, x):
        x = self.conv1(x).flatten(2).transpose(1, 2)                     # [N, g*g, d]
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], 1) + self.positional_embedding.to(x.dtype)
        x = self.ln_pre(x)
        x = self.transformer(x.transpose(0, 1)).transpose(0, 1)          # tower runs in [L, N, d]
        return self.ln_post(x[:, 0]) @ self.proj


# text tower shapes of the released models: (width, layers, heads, context, vocab); embed dim = the image tower's
TEXT = {
    "ViT-B/32": (512, 12, 8, 77, 49408),
    "ViT-B/16": (512, 12, 8, 77, 49408),
    "ViT-L/14": (768, 12, 12, 77, 49408),
    "ViT-tiny/32": (64, 2, 1, 16, 97),
    "ViT-tiny/16": (64, 2, 1, 16, 97),
}


class CausalBlock(Block):
    """Text-tower block: same as Block with the additive upper-triangular -inf attention mask."""

    def forward(self, x):
        L = x.shape[0]
        mask = torch.full((L, L), float("-inf"), dtype=x.dtype).triu_(1)
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False, attn_mask=mask)[0]
        return x + self.mlp(self.ln_2(x))


class CLIP(nn.Module):
    """Shell exposing what the reference touches: .visual, .logit_scale, .encode_image, .encode_text, .dtype."""

    def __init__(self, arch, text=False):
        super().__init__()
        patch, d, layers, heads, out_dim = ARCHS[arch]
        self.arch = arch
        self.visual = VisionTransformer(patch, d, layers, heads, out_dim)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.has_text = text
        if text:      # openai-CLIP text tower, same parameter names
            w, tl, th, ctx, vocab = TEXT[arch]
            self.context_length, self.vocab_size = ctx, vocab
            self.token_embedding = nn.Embedding(vocab, w)
            self.positional_embedding = nn.Parameter(0.01 * torch.randn(ctx, w))
            self.transformer = Tower(w, tl, th)
            self.transformer.resblocks = nn.Sequential(*[CausalBlock(w, th) for _ in range(tl)])
            self.ln_final = nn.LayerNorm(w)
            self.text_projection = nn.Parameter(w ** -0.5 * torch.randn(w, out_dim))

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image.type(self.dtype))

    def encode_text(self, tokens):
        """tokens int [n, context] -> [n, embed]; features are read at the end-of-text position = argmax of the ids."""
        if not self.has_text:
            raise NotImplementedError("build the oracle with text=True, or pre-seed text_feats")
        x = self.token_embedding(tokens).type(self.dtype) + self.positional_embedding.type(self.dtype)
        x = self.transformer(x.transpose(0, 1)).transpose(0, 1)
        x = self.ln_final(x)
        return x[torch.arange(x.shape[0]), tokens.argmax(dim=-1)] @ self.text_projection


def init_clip_(model, seed, logit_scale=100.0):
    """Seeded CLIP-style init (std d^-0.5 attention, (2d)^-0.5 MLP, depth-scaled output projections) plus
    non-trivial biases / LayerNorm affine so every term of the forward is exercised.
    exp(logit_scale) = 100 as in the released checkpoints."""
    g = torch.Generator().manual_seed(seed)
    v = model.visual
    d, layers = v.transformer.width, v.transformer.layers
    attn_std, proj_std, fc_std = d ** -0.5, (d ** -0.5) * ((2 * layers) ** -0.5), (2 * d) ** -0.5

    def rn(t, std, mean=0.0):
        with torch.no_grad():
            t.copy_(torch.randn(t.shape, generator=g) * std + mean)

    rn(v.conv1.weight, (3 * v.patch ** 2) ** -0.5)
    rn(v.class_embedding, d ** -0.5)
    rn(v.positional_embedding, d ** -0.5)
    rn(v.proj, d ** -0.5)
    for ln in [v.ln_pre, v.ln_post] + [m for b in v.transformer.resblocks for m in (b.ln_1, b.ln_2)]:
        rn(ln.weight, 0.1, 1.0)
        rn(ln.bias, 0.05)
    for b in v.transformer.resblocks:
        rn(b.attn.in_proj_weight, attn_std)
        rn(b.attn.in_proj_bias, 0.02)
        rn(b.attn.out_proj.weight, proj_std)
        rn(b.attn.out_proj.bias, 0.02)
        rn(b.mlp.c_fc.weight, fc_std)
        rn(b.mlp.c_fc.bias, 0.02)
        rn(b.mlp.c_proj.weight, proj_std)
        rn(b.mlp.c_proj.bias, 0.02)
    if getattr(model, "has_text", False):
        w, tl = model.transformer.width, model.transformer.layers
        a_std, p_std, f_std = w ** -0.5, (w ** -0.5) * ((2 * tl) ** -0.5), (2 * w) ** -0.5
        rn(model.token_embedding.weight, 0.02)
        rn(model.positional_embedding, 0.01)
        rn(model.text_projection, w ** -0.5)
        rn(model.ln_final.weight, 0.1, 1.0)
        rn(model.ln_final.bias, 0.05)
        for b in model.transformer.resblocks:
            rn(b.ln_1.weight, 0.1, 1.0); rn(b.ln_1.bias, 0.05); rn(b.ln_2.weight, 0.1, 1.0); rn(b.ln_2.bias, 0.05)
            rn(b.attn.in_proj_weight, a_std); rn(b.attn.in_proj_bias, 0.02)
            rn(b.attn.out_proj.weight, p_std); rn(b.attn.out_proj.bias, 0.02)
            rn(b.mlp.c_fc.weight, f_std); rn(b.mlp.c_fc.bias, 0.02)
            rn(b.mlp.c_proj.weight, p_std); rn(b.mlp.c_proj.bias, 0.02)
    with torch.no_grad():
        model.logit_scale.fill_(math.log(logit_scale))
    return model


def build_clip(arch, seed=0, text=False):
    return init_clip_(CLIP(arch, text=text), seed).eval()


def synth_tokens(n, context, vocab, seed):
    """Random prompts in CLIP's token format: <sot> body <eot> 0-padding; <eot> = vocab-1 is the largest id."""
    g = torch.Generator().manual_seed(seed)
    tok = torch.zeros(n, context, dtype=torch.long)
    for i in range(n):
        ln = int(torch.randint(3, context - 1, (1,), generator=g))
        tok[i, 0] = vocab - 2
        tok[i, 1:ln] = torch.randint(1, vocab - 2, (ln - 1,), generator=g)
        tok[i, ln] = vocab - 1
    return tok


def synth_text_feats(n_cls, C, seed):
    """L2-normalised seeded Gaussian [n_cls, C] standing in for encode_text output (clip_cls.py:84-85)."""
    g = torch.Generator().manual_seed(seed)
    t = torch.randn(n_cls, C, generator=g)
    return t / t.norm(dim=-1, keepdim=True)
