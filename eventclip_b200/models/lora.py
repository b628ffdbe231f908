"""LoRA containers for the CLIP image tower (reference models/lora.py).

Parameter names follow the reference so its fine-tuned checkpoints load:
  ...attn.in_proj_weight.merged_proj / lora_down_{q,k,v} / lora_up_{q,k,v}    (lora.py:101-159)
  ...attn.out_proj.linear.{weight,bias} / lora_down.weight / lora_up.weight   (lora.py:14-57)
The B200 encoder never calls these modules' forward: it merges W + up@down into the packed bf16 weights
(ec_lora_merge) whenever a parameter version changes.
"""
import torch
import torch.nn as nn


def lora_w_init_(lora_down, lora_up, r):
    """lora.py:8-11: down ~ N(0, 1/r), up = 0 (so the injected model starts identical)."""
    nn.init.normal_(lora_down, std=1.0 / r)
    nn.init.zeros_(lora_up)


def parse_lora_spec(r):
    """lora.py:357-370: int r == 'qkv-r'; strings 'qv-r', 'qkv-r', 'qkvo-r'."""
    if isinstance(r, int):
        lora_k, lora_o = True, False
    else:
        assert "q" in r and "v" in r
        lora_k, lora_o = "k" in r, "o" in r
        r = int(r.split("-")[-1])
    assert r > 0
    return r, lora_k, lora_o


class LoraInjectedLinear(nn.Module):
    def __init__(self, linear, r=4):
        super().__init__()
        if r > min(linear.in_features, linear.out_features):
            raise ValueError(f"LoRA rank {r} must be less or equal than {min(linear.in_features, linear.out_features)}")
        self.r = r
        self.linear = linear
        for p in self.linear.parameters():
            p.requires_grad = False
        kw = dict(device=linear.weight.device, dtype=linear.weight.dtype)
        self.lora_down = nn.Linear(linear.in_features, r, bias=False, **kw)
        self.lora_up = nn.Linear(r, linear.out_features, bias=False, **kw)
        lora_w_init_(self.lora_down.weight, self.lora_up.weight, r)

    @property
    def weight(self):
        return self.linear.weight + self.lora_up.weight @ self.lora_down.weight

    @property
    def bias(self):
        return self.linear.bias


class LoraInjectedMergedProj(nn.Module):
    def __init__(self, merged_proj, r=4, lora_k=True):
        super().__init__()
        d3, in_dim = merged_proj.shape
        assert d3 % 3 == 0, "MergedProj's dim must be divisible by 3"
        d = d3 // 3
        if r > min(d, in_dim):
            raise ValueError(f"LoRA rank {r} must be less or equal than {min(d, in_dim)}")
        self.d_model, self.r, self.lora_k = d, r, lora_k
        self.merged_proj = merged_proj
        self.merged_proj.requires_grad = False
        kw = dict(device=merged_proj.device, dtype=merged_proj.dtype)
        for n in ("q", "v") + (("k",) if lora_k else ()):
            down = nn.Parameter(torch.empty(r, in_dim, **kw))
            up = nn.Parameter(torch.empty(d, r, **kw))
            lora_w_init_(down, up, r)
            setattr(self, f"lora_down_{n}", down)
            setattr(self, f"lora_up_{n}", up)

    def forward(self):
        """The merged [3d, in_dim] weight (fp32 autograd view, for inspection / tests)."""
        d = self.d_model
        parts = []
        for j, n in enumerate("qkv"):
            w = self.merged_proj[j * d:(j + 1) * d]
            if hasattr(self, f"lora_up_{n}"):
                w = w + getattr(self, f"lora_up_{n}") @ getattr(self, f"lora_down_{n}")
            parts.append(w)
        return torch.cat(parts, dim=0)


class LoraAttention(nn.Module):
    """Stands where nn.MultiheadAttention stood in a residual block, with LoRA-wrapped projections."""

    def __init__(self, mha, r):
        super().__init__()
        r, lora_k, lora_o = parse_lora_spec(r)
        self.embed_dim, self.num_heads = mha.embed_dim, mha.num_heads
        for p in mha.parameters():
            p.requires_grad = False
        self.in_proj_weight = LoraInjectedMergedProj(mha.in_proj_weight, r=r, lora_k=lora_k)
        self.in_proj_bias = mha.in_proj_bias
        self.out_proj = LoraInjectedLinear(mha.out_proj, r=r) if lora_o else mha.out_proj


def inject_trainable_lora(model, r=4):
    """lora.py:385-403: replace every nn.MultiheadAttention named `...attn` under `model`."""
    targets = [(name, m) for name, m in model.named_modules() if isinstance(m, nn.MultiheadAttention)]
    for name, mha in targets:
        assert name.endswith("attn")
        parent = model.get_submodule(name.rsplit(".", 1)[0]) if "." in name else model
        parent.attn = LoraAttention(mha, r)
    if hasattr(model, "invalidate_packed"):
        model.invalidate_packed()
    return model
