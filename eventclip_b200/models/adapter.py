"""Feature adapters of the few-shot classifier (reference models/adapter.py) on the B200 path.

The modules own the parameters under the reference's names (`in_proj.*`, `transformer_encoder.layers.i.*`,
`out_proj.*`) so its checkpoints load unchanged; `forward` runs the library's fp32 kernels
(ec_gemm_f32 / ec_layernorm_f32 / ec_adapter_attention / ec_blend), never nn.TransformerEncoder itself.
"""
import torch
import torch.nn as nn

from .. import ops


def _residual_weight(residual):
    """adapter.py:13-18: bool -> 0.5 / 0.0, float must lie in [0, 1]."""
    assert isinstance(residual, (bool, float))
    if isinstance(residual, bool):
        return 0.5 if residual else 0.0
    assert 0.0 <= residual <= 1.0
    return residual


class IdentityAdapter(nn.Module):
    """adapter.py:35-50 -- passes features through; the dummy parameter records dtype/device."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        self.residual = 0.0
        self.dummy = nn.Parameter(torch.zeros(1), requires_grad=False)

    def forward(self, feats, valid_masks):
        return feats

    @property
    def dtype(self):
        return self.dummy.dtype


def _adapter_params(ad):
    """Parameters of a TransformerAdapter in the fixed order _AdapterFn passes them to autograd."""
    ps = [ad.in_proj.weight, ad.in_proj.bias]
    for l in ad.transformer_encoder.layers:
        ps += [l.norm1.weight, l.norm1.bias, l.self_attn.in_proj_weight, l.self_attn.in_proj_bias, l.self_attn.out_proj.weight,
               l.self_attn.out_proj.bias, l.norm2.weight, l.norm2.bias, l.linear1.weight, l.linear1.bias, l.linear2.weight,
               l.linear2.bias]
    return ps + [ad.out_proj.weight, ad.out_proj.bias]


class _AdapterFn(torch.autograd.Function):
    """Training route of the pre-norm TransformerAdapter: the forward below with its activations kept, and an explicit
    backward through the library's fp32 kernels (the reference differentiates nn.TransformerEncoder with autograd,
    models/adapter.py:82-105).  Dropout is not applied (the reference's 0.1 dropout is a train-time regulariser with its own
    RNG stream; nn.Module.eval() semantics are what the kernels implement)."""

    @staticmethod
    def forward(fctx, ad, flat, valid_u8, B, T, *params):
        D, H = ad.d_model, ad.num_heads
        w = [p.detach().contiguous() for p in params]
        saved = []
        x = ops.gemm_f32(flat, w[0], w[1])
        for i in range(len(ad.transformer_encoder.layers)):
            n1w, n1b, wi, bi, wo, bo, n2w, n2b, w1, b1, w2, b2 = w[2 + 12 * i: 14 + 12 * i]
            h1 = ops.layernorm_f32(x, n1w, n1b)
            qkv = ops.gemm_f32(h1, wi, bi)
            a = ops.adapter_attention(qkv, valid_u8, B, T, D, H)
            x2 = ops.gemm_f32(a, wo, bo, res=x)
            h2 = ops.layernorm_f32(x2, n2w, n2b)
            u = ops.gemm_f32(h2, w1, b1, act=1)
            x3 = ops.gemm_f32(u, w2, b2, res=x2)
            saved.append((x, h1, qkv, a, x2, h2, u))
            x = x3
        new = ops.gemm_f32(x, w[-2], w[-1])
        out = ops.blend(flat, new, ad.residual)
        fctx.ad, fctx.w, fctx.saved, fctx.x_last, fctx.flat, fctx.valid_u8, fctx.geom = ad, w, saved, x, flat, valid_u8, (B, T)
        return out

    @staticmethod
    def backward(fctx, d_out):
        ad, w, flat, valid_u8 = fctx.ad, fctx.w, fctx.flat, fctx.valid_u8
        B, T = fctx.geom
        D, H = ad.d_model, ad.num_heads
        M = B * T
        d_out = d_out.contiguous()
        zeros = torch.zeros_like(d_out)
        d_new = ops.blend(zeros, d_out, ad.residual)                      # (1 - r) d_out
        grads = [None] * len(w)
        grads[-2] = ops.mm_f32(d_new, fctx.x_last, trans_a=True)          # out_proj.weight [C, D]
        grads[-1] = ops.colsum(d_new, n_part=1)
        dx = ops.mm_f32(d_new, w[-2])                                     # [M, D]
        for i in reversed(range(len(fctx.saved))):
            x, h1, qkv, a, x2, h2, u = fctx.saved[i]
            n1w, n1b, wi, bi, wo, bo, n2w, n2b, w1, b1, w2, b2 = w[2 + 12 * i: 14 + 12 * i]
            g = [None] * 12
            # x3 = x2 + linear2(relu(linear1(LN2(x2))))
            g[10] = ops.mm_f32(dx, u, trans_a=True)
            g[11] = ops.colsum(dx, n_part=1)
            d_pre = ops.relu_bwd(u, ops.mm_f32(dx, w2))
            g[8] = ops.mm_f32(d_pre, h2, trans_a=True)
            g[9] = ops.colsum(d_pre, n_part=1)
            d_h2 = ops.mm_f32(d_pre, w1)
            g[6], g[7] = ops.layernorm_param_grad(x2, d_h2, M, D)
            dx2 = ops.layernorm_bwd(x2, d_h2, n2w, M, D, acc=dx)
            # x2 = x + out_proj(attention(in_proj(LN1(x))))
            g[4] = ops.mm_f32(dx2, a, trans_a=True)
            g[5] = ops.colsum(dx2, n_part=1)
            d_a = ops.mm_f32(dx2, wo)
            d_qkv = ops.adapter_attention_bwd(qkv, valid_u8, d_a, B, T, D, H)
            g[2] = ops.mm_f32(d_qkv, h1, trans_a=True)
            g[3] = ops.colsum(d_qkv, n_part=1)
            d_h1 = ops.mm_f32(d_qkv, wi)
            g[0], g[1] = ops.layernorm_param_grad(x, d_h1, M, D)
            dx = ops.layernorm_bwd(x, d_h1, n1w, M, D, acc=dx2)
            grads[2 + 12 * i: 14 + 12 * i] = g
        grads[0] = ops.mm_f32(dx, flat, trans_a=True)
        grads[1] = ops.colsum(dx, n_part=1)
        return (None, None, None, None, None, *grads)          # the features come from the frozen CLIP: no gradient


class TransformerAdapter(nn.Module):
    """adapter.py:53-109 -- in_proj -> encoder layers over the views (key-padding mask; pre-norm, or post-norm when
    norm_first=False) -> out_proj -> residual blend r*in + (1-r)*new."""

    def __init__(self, in_dim, d_model=256, num_heads=4, ffn_dim=256 * 4, norm_first=True, num_layers=2,
                 residual=False):
        super().__init__()
        self.norm_first = bool(norm_first)
        self.residual = _residual_weight(residual)
        self.d_model, self.num_heads = d_model, num_heads
        layer = nn.TransformerEncoderLayer(d_model=d_model, nhead=num_heads, dim_feedforward=ffn_dim,
                                           norm_first=norm_first, batch_first=True)
        self.transformer_encoder = nn.TransformerEncoder(encoder_layer=layer, num_layers=num_layers)
        self.in_proj = nn.Linear(in_dim, d_model)
        self.out_proj = nn.Linear(d_model, in_dim)

    @property
    def dtype(self):
        return self.in_proj.weight.dtype

    def forward(self, feats, valid_masks):
        """feats: CUDA float32 [B,T,C]; valid_masks: [B,T] bool.  Returns [B,T,C]."""
        B, T, C = feats.shape
        D, H = self.d_model, self.num_heads
        flat = feats.reshape(B * T, C).contiguous()
        valid_u8 = valid_masks.to(device=feats.device, dtype=torch.uint8).contiguous()
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            if not self.norm_first:
                raise NotImplementedError("training of post-norm (norm_first=False) adapters is not built; no shipped config uses it")
            if flat.requires_grad:
                raise NotImplementedError("the adapter's training route expects features of a frozen CLIP (few-shot setting)")
            return _AdapterFn.apply(self, flat.detach(), valid_u8, B, T, *_adapter_params(self)).view(B, T, C)
        w = lambda p: p.detach().contiguous()
        x = ops.gemm_f32(flat, w(self.in_proj.weight), w(self.in_proj.bias))
        for layer in self.transformer_encoder.layers:
            attn_w = (w(layer.self_attn.in_proj_weight), w(layer.self_attn.in_proj_bias))
            out_w = (w(layer.self_attn.out_proj.weight), w(layer.self_attn.out_proj.bias))
            if self.norm_first:        # x = x + sa(norm1(x)); x = x + ff(norm2(x))
                h = ops.layernorm_f32(x, w(layer.norm1.weight), w(layer.norm1.bias))
                a = ops.adapter_attention(ops.gemm_f32(h, *attn_w), valid_u8, B, T, D, H)
                x = ops.gemm_f32(a, *out_w, res=x)
                h = ops.layernorm_f32(x, w(layer.norm2.weight), w(layer.norm2.bias))
                h = ops.gemm_f32(h, w(layer.linear1.weight), w(layer.linear1.bias), act=1)
                x = ops.gemm_f32(h, w(layer.linear2.weight), w(layer.linear2.bias), res=x)
            else:                      # post-norm: x = norm1(x + sa(x)); x = norm2(x + ff(x))
                a = ops.adapter_attention(ops.gemm_f32(x, *attn_w), valid_u8, B, T, D, H)
                x = ops.layernorm_f32(ops.gemm_f32(a, *out_w, res=x), w(layer.norm1.weight), w(layer.norm1.bias))
                h = ops.gemm_f32(x, w(layer.linear1.weight), w(layer.linear1.bias), act=1)
                x = ops.layernorm_f32(ops.gemm_f32(h, w(layer.linear2.weight), w(layer.linear2.bias), res=x),
                                      w(layer.norm2.weight), w(layer.norm2.bias))
        new = ops.gemm_f32(x, w(self.out_proj.weight), w(self.out_proj.bias))
        return ops.blend(flat, new, self.residual).view(B, T, C)
