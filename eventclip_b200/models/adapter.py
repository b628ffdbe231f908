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


class TransformerAdapter(nn.Module):
    """adapter.py:53-109 -- in_proj -> pre-norm encoder layers over the views (key-padding mask) -> out_proj ->
    residual blend r*in + (1-r)*new."""

    def __init__(self, in_dim, d_model=256, num_heads=4, ffn_dim=256 * 4, norm_first=True, num_layers=2,
                 residual=False):
        super().__init__()
        if not norm_first:
            raise NotImplementedError("only norm_first=True (every shipped config) is built")
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
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and self.training:
            raise NotImplementedError("adapter training (backward) is not built on the B200 path yet")
        B, T, C = feats.shape
        D, H = self.d_model, self.num_heads
        flat = feats.reshape(B * T, C).contiguous()
        valid_u8 = valid_masks.to(device=feats.device, dtype=torch.uint8).contiguous()
        w = lambda p: p.detach().contiguous()
        x = ops.gemm_f32(flat, w(self.in_proj.weight), w(self.in_proj.bias))
        for layer in self.transformer_encoder.layers:
            h = ops.layernorm_f32(x, w(layer.norm1.weight), w(layer.norm1.bias))
            qkv = ops.gemm_f32(h, w(layer.self_attn.in_proj_weight), w(layer.self_attn.in_proj_bias))
            a = ops.adapter_attention(qkv, valid_u8, B, T, D, H)
            x = ops.gemm_f32(a, w(layer.self_attn.out_proj.weight), w(layer.self_attn.out_proj.bias), res=x)
            h = ops.layernorm_f32(x, w(layer.norm2.weight), w(layer.norm2.bias))
            h = ops.gemm_f32(h, w(layer.linear1.weight), w(layer.linear1.bias), act=1)
            x = ops.gemm_f32(h, w(layer.linear2.weight), w(layer.linear2.bias), res=x)
        new = ops.gemm_f32(x, w(self.out_proj.weight), w(self.out_proj.bias))
        return ops.blend(flat, new, self.residual).view(B, T, C)
