"""CLIP object of the B200 path: same attribute / parameter names as openai-CLIP's `CLIP` and `VisionTransformer`
(`clip/model.py`, third-party to the reference), so the reference's classifiers, LoRA injection and checkpoints see
the object they expect (.visual.conv1, .visual.proj, .visual.class_embedding,
.visual.transformer.resblocks[i].attn : nn.MultiheadAttention, .logit_scale, .encode_image, module-level load()).

nn.Module / nn.Parameter are used as parameter containers only.  `VisionTransformer.forward` does not run a single
PyTorch op on the data path: it packs weights to bf16 once and enqueues the library's kernels
(tcgen05 GEMMs with fused epilogues, LayerNorm, attention) through the C ABI.
"""
import math
import os
from collections import OrderedDict

import torch
import torch.nn as nn

from . import _lib as L
from . import ops

ARCHS = {
    # name: (patch, width, layers, heads, embed_dim) -- openai-CLIP released ViTs
    "ViT-B/32": (32, 768, 12, 12, 512),
    "ViT-B/16": (16, 768, 12, 12, 512),
    "ViT-L/14": (14, 1024, 24, 16, 768),
    # small shapes for tests
    "ViT-tiny/32": (32, 128, 2, 2, 64),
    "ViT-tiny/16": (16, 128, 2, 2, 64),
}

# algorithmic FLOPs per image (2*MAC), SURVEY.md section 8(d)
def flops_per_image(arch):
    P, d, layers, _, C = ARCHS[arch]
    Ltok = (224 // P) ** 2 + 1
    patch = 2 * (Ltok - 1) * (3 * P * P) * d
    blocks = layers * (2 * Ltok * d * 3 * d + 2 * Ltok * d * d + 2 * 2 * Ltok * d * 4 * d + 2 * 2 * Ltok * Ltok * d)
    return patch + blocks + 2 * d * C


class QuickGELU(nn.Module):
    """Parameter-free marker module (keeps the mlp Sequential's key names c_fc / gelu / c_proj)."""

    def forward(self, x):  # pragma: no cover - never on the B200 data path
        raise L.ECError("the B200 path applies QuickGELU inside the c_fc GEMM epilogue")


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d, d * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d * 4, d))]))
        self.ln_2 = nn.LayerNorm(d)


class Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers, self.heads = width, layers, heads
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])


def _attn_weights(attn):
    """(in_proj fp32 [3d,d] parts + optional LoRA factors, in_proj_bias, out_proj W/b + optional LoRA) of an
    nn.MultiheadAttention or of the LoRA-injected variant (eventclip_b200.models.lora / reference models/lora.py)."""
    ipw = attn.in_proj_weight
    lora_in = None
    if isinstance(ipw, nn.Module):   # LoraInjectedMergedProj: .merged_proj + lora_{up,down}_{q,k,v}
        lora_in = ipw
        W = ipw.merged_proj
    else:
        W = ipw
    op = attn.out_proj
    lora_out = None
    if hasattr(op, "lora_up"):       # LoraInjectedLinear: .linear + lora_down / lora_up
        lora_out = op
        oW, ob = op.linear.weight, op.linear.bias
    else:
        oW, ob = op.weight, op.bias
    return W, lora_in, attn.in_proj_bias, oW, ob, lora_out


def merge_attn_weights(blk, d, dev, w_in=None, w_out=None):
    """bf16 in_proj [3d,d] and out_proj [d,d] of a block with the LoRA factors merged (models/lora.py:138-149, 49-52);
    writes into w_in / w_out when given."""
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    W, lora_in, ib, oW, ob, lora_out = _attn_weights(blk.attn)
    W = f32(W)
    if lora_in is None:
        w_in = ops.f32_to_bf16(W, dst=w_in)
    else:
        if w_in is None:
            w_in = torch.empty((3 * d, d), dtype=torch.bfloat16, device=dev)
        for j, n in enumerate("qkv"):
            up = getattr(lora_in, f"lora_up_{n}", None)
            down = getattr(lora_in, f"lora_down_{n}", None)
            ops.lora_merge(W[j * d:(j + 1) * d], f32(up) if up is not None else None,
                           f32(down) if down is not None else None, out=w_in[j * d:(j + 1) * d])
    if lora_out is None:
        w_out = ops.f32_to_bf16(f32(oW), dst=w_out)
    else:
        w_out = ops.lora_merge(f32(oW), f32(lora_out.lora_up.weight), f32(lora_out.lora_down.weight), out=w_out)
    return w_in, w_out, ib, ob, (lora_in is not None or lora_out is not None)


def pack_blocks(resblocks, d, dev):
    """bf16 GEMM weights (LoRA factors merged) + fp32 biases / LayerNorm affine per block."""
    f32 = lambda t: t.detach().to(torch.float32).contiguous()
    blocks = []
    for blk in resblocks:
        w_in, w_out, ib, ob, has_lora = merge_attn_weights(blk, d, dev)
        blocks.append(dict(
            ln1=(f32(blk.ln_1.weight), f32(blk.ln_1.bias)), ln2=(f32(blk.ln_2.weight), f32(blk.ln_2.bias)),
            w_in=w_in, b_in=f32(ib), w_out=w_out, b_out=f32(ob), has_lora=has_lora,
            w_fc=ops.f32_to_bf16(f32(blk.mlp.c_fc.weight)), b_fc=f32(blk.mlp.c_fc.bias),
            w_proj=ops.f32_to_bf16(f32(blk.mlp.c_proj.weight)), b_proj=f32(blk.mlp.c_proj.bias)))
    return blocks


def pack_blocks_f16(resblocks, d, dev):
    """fp16 copies of the GEMM weights of every block (LoRA factors merged in fp32, models/lora.py:138-149, 49-52) for the
    fp16-operand inference forward.  Pack-time arithmetic in torch, like pack_blocks_ln."""
    f32 = lambda t: t.detach().to(torch.float32)
    out = []
    for blk in resblocks:
        W, lora_in, _, oW, _, lora_out = _attn_weights(blk.attn)
        W = f32(W).clone()
        if lora_in is not None:
            for j, n in enumerate("qkv"):
                up, down = getattr(lora_in, f"lora_up_{n}", None), getattr(lora_in, f"lora_down_{n}", None)
                if up is not None and down is not None:
                    W[j * d:(j + 1) * d] += f32(up) @ f32(down)
        oW = f32(oW)
        if lora_out is not None:
            oW = oW + f32(lora_out.lora_up.weight) @ f32(lora_out.lora_down.weight)
        h = lambda t: t.to(torch.float16).contiguous()
        out.append(dict(w_in=h(W), w_out=h(oW), w_fc=h(f32(blk.mlp.c_fc.weight)), w_proj=h(f32(blk.mlp.c_proj.weight))))
    return out


def pack_blocks_ln(resblocks, d, dev):
    """Per block, the operands of the LayerNorm-folded GEMMs (ec_gemm_ln): Wg = fp16(gamma * W_eff) with every row centred
    (LoRA factors merged as in models/lora.py:138-149) and c = W_eff beta + b, so that   LN(x) W^T + b = rstd (x Wg^T) + c
    (the mean term drops out because sum_k (x_k - mean) = 0).  Pack-time arithmetic (runs when a parameter changes, not
    per batch) is done with torch in fp32."""
    f32 = lambda t: t.detach().to(torch.float32)
    out = []
    for blk in resblocks:
        W, lora_in, ib, _, _, _ = _attn_weights(blk.attn)
        W = f32(W).clone()
        if lora_in is not None:
            for j, n in enumerate("qkv"):
                up, down = getattr(lora_in, f"lora_up_{n}", None), getattr(lora_in, f"lora_down_{n}", None)
                if up is not None and down is not None:
                    W[j * d:(j + 1) * d] += f32(up) @ f32(down)
        e = {}
        for name, Wf, bf_, ln in (("in", W, f32(ib), blk.ln_1), ("fc", f32(blk.mlp.c_fc.weight), f32(blk.mlp.c_fc.bias), blk.ln_2)):
            wg = Wf * f32(ln.weight)[None, :]
            wg = (wg - wg.mean(1, keepdim=True)).to(torch.float16).contiguous()
            # the fp16 rounding leaves a row sum r_j != 0 that would multiply the row mean of x: fold it into the largest
            # element of the row, so that |sum_k Wg[j,k]| drops to one rounding of that element
            r = wg.to(torch.float32).sum(1)
            kmax = wg.abs().argmax(1, keepdim=True)
            wg.scatter_(1, kmax, (wg.gather(1, kmax).to(torch.float32) - r[:, None]).to(torch.float16))
            e["wg_" + name] = wg
            e["c_" + name] = (Wf @ f32(ln.bias) + bf_).contiguous()
        _, _, _, _, ob, _ = _attn_weights(blk.attn)
        e["b_out"] = f32(ob).contiguous().clone()
        e["b_proj"] = f32(blk.mlp.c_proj.bias).contiguous().clone()
        out.append(e)
    return out


def run_blocks_ln(x, blocks, lnb, n_seq, Ltok, d, heads, x2=None):
    """x: fp16 residual stream [n_seq*Ltok, d]; x2: optional [2, M, d] (hi, lo) pair of fp16 planes with x = x2[0] -- the
    residual updates then keep the rounding residue of every update in x2[1] (ec_gemm_bf16_stats2).  No LayerNorm kernel: in_proj and c_fc read the stream itself as their A
    operand and apply the normalisation in their epilogues from the row statistics that the previous residual epilogue
    (out_proj / c_proj) wrote next to the rows.  blocks: bf16 or fp16 weight copies; the activations between the GEMMs
    (qkv, attention output, MLP hidden) take the same 16-bit dtype."""
    M, dev = n_seq * Ltok, x.device
    parts = ops.gemm_stats_parts(d)
    stats = torch.empty((M, parts, 2), dtype=torch.float32, device=dev)
    ops.row_stats_f16(x, stats, parts)                       # the rows ln_pre wrote
    dt = blocks[0]["w_out"].dtype
    qkv = torch.empty((M, 3 * d), dtype=dt, device=dev)
    att = torch.empty((M, d), dtype=dt, device=dev)
    hid = torch.empty((M, 4 * d), dtype=dt, device=dev)
    for b, f in zip(blocks, lnb):
        ops.gemm_ln(x, f["wg_in"], None, f["c_in"], stats, parts, "bf16", out=qkv)
        ops.attention(qkv, att, n_seq, Ltok, heads)
        if x2 is None:
            ops.gemm_bf16_stats(att, b["w_out"], f["b_out"], x, stats)
        else:
            ops.gemm_bf16_stats2(att, b["w_out"], f["b_out"], x2, stats)
        ops.gemm_ln(x, f["wg_fc"], None, f["c_fc"], stats, parts, "bf16_qgelu", out=hid)
        if x2 is None:
            ops.gemm_bf16_stats(hid, b["w_proj"], f["b_proj"], x, stats)
        else:
            ops.gemm_bf16_stats2(hid, b["w_proj"], f["b_proj"], x2, stats)
    return x


def run_blocks(x, blocks, n_seq, Ltok, d, heads, causal=False, w16=None):
    """x: residual stream [n_seq*Ltok, d], fp32 or fp16, updated in place by the GEMM epilogues.  w16: per-block fp16 weight
    copies (pack_blocks_f16) -> fp16 operands and activations; None -> the bf16 copies of `blocks`."""
    M, dev = n_seq * Ltok, x.device
    resadd = "f16_resadd" if x.dtype == torch.float16 else "f32_resadd"
    dt = torch.bfloat16 if w16 is None else torch.float16
    xn = torch.empty((M, d), dtype=dt, device=dev)
    qkv = torch.empty((M, 3 * d), dtype=dt, device=dev)
    att = torch.empty((M, d), dtype=dt, device=dev)
    hid = torch.empty((M, 4 * d), dtype=dt, device=dev)
    ln_out = dict(out_f16=xn) if dt == torch.float16 else dict(out_bf16=xn)
    for i, b in enumerate(blocks):
        w = b if w16 is None else w16[i]
        ops.layernorm(x, *b["ln1"], M, d, **ln_out)
        ops.gemm_bf16(xn, w["w_in"], b["b_in"], "bf16", out=qkv)
        ops.attention(qkv, att, n_seq, Ltok, heads, causal=causal)
        ops.gemm_bf16(att, w["w_out"], b["b_out"], resadd, out=x, res=x)
        ops.layernorm(x, *b["ln2"], M, d, **ln_out)
        ops.gemm_bf16(xn, w["w_fc"], b["b_fc"], "bf16_qgelu", out=hid)
        ops.gemm_bf16(hid, w["w_proj"], b["b_proj"], resadd, out=x, res=x)
    return x


# constants of CLIP's preprocess (Normalize), confirmed by the reference's method.py:17-18
CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


class VisionTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        if input_resolution != 224:
            raise L.ECError("the fused path is built for CLIP's 224-pixel ViTs")
        self.input_resolution, self.output_dim = input_resolution, output_dim
        self.patch_size, self.width, self.heads = patch_size, width, heads
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.grid = input_resolution // patch_size
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn(self.grid ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))
        self._packed = None
        self._packed_key = None
        # dtype of the residual stream of the inference forward.  fp16 = the reference's own precision on CUDA (clip.load
        # keeps the model in fp16, test.py:26-29) at half the residual traffic; EC_RESIDUAL=fp32 keeps it in float32.
        # The fine-tune forward (train.py) always uses fp32, as the reference's train.py:27-29 does.
        self.residual_dtype = torch.float32 if os.environ.get("EC_RESIDUAL", "fp16") == "fp32" else torch.float16
        # fp16 stream kept as a (hi, lo) pair of fp16 planes (EC_RESIDUAL=fp16x2): hi = fp16(x) is the GEMM operand, lo the
        # residue of that rounding, carried through every residual update (ec_gemm_bf16_stats2).  The accuracy of a float32
        # stream without the LayerNorm kernels a float32 stream needs; used with the folded LayerNorm only.
        self.residual_split = os.environ.get("EC_RESIDUAL", "fp16") == "fp16x2"
        # fused events route: the event kernel writes ONE gray plane per patch row and conv1 is folded onto it (packed_gray);
        # EC_GRAY_FOLD=0 keeps the three normalised channels (the reference's tensor, rounded to 16 bits)
        self.gray_fold = os.environ.get("EC_GRAY_FOLD", "1") != "0"
        self._packed_gray, self._packed_gray_key = None, None
        # ln_1 / ln_2 folded into the in_proj / c_fc GEMMs (fp16 stream only); EC_LN_FOLD=0 keeps the LayerNorm kernels
        self.fold_ln = os.environ.get("EC_LN_FOLD", "1") != "0"
        self._packed_ln, self._packed_ln_key = None, None
        # 16-bit format of the tensor-core operands and of the activations between the GEMMs in the INFERENCE forward.  fp16 is
        # what the reference itself runs on CUDA (clip.load keeps fp16 weights, test.py:26-29) and rounds 8 times finer than bf16
        # at the same tcgen05 rate: on event frames the encoder's error relative to the part of the features that differs
        # between samples drops accordingly (tests/test_bench_geometry_gpu.py).  EC_OPERANDS=bf16 selects bf16 (wider exponent
        # range).  The fine-tune forward / backward (train.py) always uses bf16 operands and an fp32 residual stream.
        self.operand_dtype = torch.bfloat16 if os.environ.get("EC_OPERANDS", "fp16") == "bf16" else torch.float16
        self._packed16, self._packed16_key = None, None
        # bumped by everything that rewrites parameters behind the version counters' back (the fused optimizer kernel of
        # train.FineTuner writes through raw pointers): the LayerNorm-folded operands and captured graphs key on it
        self._epoch = 0

    # -- weight packing -------------------------------------------------------------------------------------------
    @property
    def k_patch(self):
        """Row length of the im2col / patch matrix: 3*P*P rounded up to a multiple of 8 (16-byte TMA stride)."""
        k = 3 * self.patch_size ** 2
        return (k + 7) // 8 * 8

    def _version_key(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def invalidate_packed(self):
        """Call after mutating weights in a way the version counters cannot see (e.g. .data swaps)."""
        self._packed = None
        self._packed_gray = None
        self._packed_ln = None
        self._packed16 = None
        self._epoch += 1

    def validate_ln_fold(self, patches, n_img, tol=1e-2):
        """Guard of the LayerNorm folding (ln_1 -> in_proj, ln_2 -> c_fc): the folded GEMMs evaluate var = E[x^2] - mean^2 on
        the un-centred fp16 rows, which loses log2(|mean| / sigma) bits where the LayerNorm kernel would not.  Random-init and
        pretrained CLIP streams (per-row means of the order of sigma, a few outlier CHANNELS) are far from that regime, but a
        tower whose rows carry |mean| >> sigma is not.  Runs the encoder once with and once without the folding on `patches`
        and switches the folding off for this tower when the features differ by more than `tol` (relative L2) or are not
        finite.  Returns the measured difference.  GraphedClassifier calls this on its first warm-up batch."""
        if not (self.fold_ln and self.residual_dtype == torch.float16 and ops.gemm_stats_parts(self.width) <= 8):
            self._fold_checked = True
            return 0.0
        with torch.no_grad():
            a = self.forward_patches(patches, n_img).float()
            self.fold_ln = False
            b = self.forward_patches(patches, n_img).float()
            self.fold_ln = True
        err = float(((a - b).norm() / b.norm().clamp_min(1e-30)).item())
        if not (err <= tol) or not bool(torch.isfinite(a).all()):
            import warnings
            warnings.warn(f"LayerNorm folding disabled for this tower: folded vs unfolded features differ by {err:.3g} (> {tol})")
            self.fold_ln = False
        if not bool(torch.isfinite(b).all()):
            # the fp16 residual stream itself leaves its range (|x| > 65504 somewhere): keep the stream in float32 for this tower,
            # as EC_RESIDUAL=fp32 would (the reference's own fp16 CUDA inference overflows on such a tower too)
            import warnings
            warnings.warn("fp16 residual stream overflows on this tower: switching its residual stream to float32")
            self.residual_dtype = torch.float32
        self._fold_checked = True
        return err

    @property
    def k_gray(self):
        """Row length of the single-plane (gray) patch matrix: P*P rounded up to a multiple of 8."""
        return (self.patch_size ** 2 + 7) // 8 * 8

    @property
    def patch_fmt(self):
        """ec_event2img output format that feeds this tower's patch GEMM in the inference forward."""
        f16 = self.operand_dtype == torch.float16
        if self.gray_fold:
            return "gray_f16" if f16 else "gray"
        return "patch_f16" if f16 else "patch"

    @property
    def patch_ldk(self):
        """Row stride of that format: k_gray (one plane, conv1 folded) or k_patch (three normalised channels)."""
        return self.k_gray if self.gray_fold else self.k_patch

    def packed_gray(self):
        """conv1 folded onto the gray plane.  Event frames are grayscale (datasets/vis.py:94-104): the three channels CLIP's
        preprocess yields are x_c = (g / 255 - mean_c) / std_c of ONE resampled byte g, so
            conv1(x)[j] = sum_c sum_k W[j,c,k] x_c[k] = sum_k (sum_c W[j,c,k] / (255 std_c)) g[k] - sum_c (mean_c / std_c) sum_k W[j,c,k].
        The event kernel writes g / 128 exactly (EC_OUT_GRAY_*_PATCH); Wg = 128 sum_c W_c / (255 std_c) keeps the weights at their
        original scale, and the constant term joins the positional embedding of the patch tokens.  A third of the patch rows'
        bytes and of the patch GEMM's K; the A operand carries no rounding error at all.  SURVEY section 8(d) names this folding;
        FLOPs and bytes are still reported against the three-channel formulas."""
        key = (self._version_key(), self._epoch, self.operand_dtype)
        if getattr(self, "_packed_gray", None) is None or key != self._packed_gray_key:
            d, P, dev = self.width, self.patch_size, self.proj.device
            w = self.conv1.weight.detach().to(torch.float64).reshape(d, 3, P * P)
            mean = torch.tensor(CLIP_MEAN, dtype=torch.float64, device=dev)
            std = torch.tensor(CLIP_STD, dtype=torch.float64, device=dev)
            wg = (w * (128.0 / (255.0 * std)).view(1, 3, 1)).sum(1)
            bias = -(w.sum(2) * (mean / std).view(1, 3)).sum(1)
            wp = torch.zeros((d, self.k_gray), dtype=torch.float32, device=dev)
            wp[:, :P * P] = wg.to(torch.float32)
            pos = self.positional_embedding.detach().to(torch.float32).clone()
            pos[1:] += bias.to(torch.float32)
            new = dict(conv1=wp.to(self.operand_dtype).contiguous(), pos=pos.contiguous())
            if getattr(self, "_packed_gray", None) is None or self._packed_gray["conv1"].dtype != new["conv1"].dtype:
                self._packed_gray = new
            else:                                   # same buffers: captured graphs stay valid
                self._packed_gray["conv1"].copy_(new["conv1"])
                self._packed_gray["pos"].copy_(new["pos"])
            self._packed_gray_key = key
        return self._packed_gray

    def mark_weights_changed(self):
        """Parameters were updated in place without touching their version counters (ec_adam on the flat buffer).  The
        bf16 copies are rewritten by refresh_packed(); the LayerNorm-folded operands (merged LoRA q/k/v, gamma*W, W beta + b)
        are recomputed -- into the same buffers -- by the next packed_ln() call, and GraphedClassifier captures again."""
        self._epoch += 1

    def packed(self):
        """bf16 copies of the GEMM weights in the layout the kernels read (LoRA factors merged, models/lora.py:138-149).
        Rebuilt when any parameter's version counter changes (optimizer step, load_state_dict)."""
        key = self._version_key()
        if self._packed is not None and key == self._packed_key:
            return self._packed
        dev = self.proj.device
        if dev.type != "cuda":
            raise L.ECError("VisionTransformer weights must live on a CUDA device (B200); there is no CPU path")
        d, P = self.width, self.patch_size
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        pk = {}
        w = f32(self.conv1.weight).reshape(d, 3 * P * P)
        if self.k_patch != 3 * P * P:
            wp = torch.zeros((d, self.k_patch), dtype=torch.float32, device=dev)
            wp[:, :3 * P * P] = w
            w = wp
        pk["conv1"] = ops.f32_to_bf16(w)
        pk["cls"], pk["pos"] = f32(self.class_embedding), f32(self.positional_embedding)
        pk["ln_pre"] = (f32(self.ln_pre.weight), f32(self.ln_pre.bias))
        pk["ln_post"] = (f32(self.ln_post.weight), f32(self.ln_post.bias))
        pk["proj"] = ops.f32_to_bf16(f32(self.proj).t().contiguous())   # [C, d] so that out = x @ proj
        blocks = pack_blocks(self.transformer.resblocks, d, dev)
        pk["blocks"] = blocks
        self._packed, self._packed_key = pk, key
        return pk

    def packed_ln(self):
        """Operands of the LayerNorm-folded GEMMs, rebuilt when any parameter's version counter changes."""
        key = (self._version_key(), self._epoch)
        if self._packed_ln is None or key != self._packed_ln_key:
            new = pack_blocks_ln(self.transformer.resblocks, self.width, self.proj.device)
            if self._packed_ln is None:
                self._packed_ln = new
            else:       # rewrite in place: captured CUDA graphs keep reading the same buffers
                for old_b, new_b in zip(self._packed_ln, new):
                    for k, v in new_b.items():
                        old_b[k].copy_(v)
            self._packed_ln_key = key
        return self._packed_ln

    def packed_f16(self):
        """fp16 copies of the GEMM weights for the fp16-operand inference forward (conv1, proj, per block in_proj / out_proj /
        c_fc / c_proj with the LoRA factors merged).  Rebuilt -- into the same buffers, so captured graphs stay valid -- when a
        parameter's version counter or the tower's epoch (mark_weights_changed) moves."""
        key = (self._version_key(), self._epoch)
        if self._packed16 is None or key != self._packed16_key:
            d, P, dev = self.width, self.patch_size, self.proj.device
            f32 = lambda t: t.detach().to(torch.float32)
            w = f32(self.conv1.weight).reshape(d, 3 * P * P)
            if self.k_patch != 3 * P * P:
                wp = torch.zeros((d, self.k_patch), dtype=torch.float32, device=dev)
                wp[:, :3 * P * P] = w
                w = wp
            new = dict(conv1=w.to(torch.float16).contiguous(), proj=f32(self.proj).t().to(torch.float16).contiguous(),
                       blocks=pack_blocks_f16(self.transformer.resblocks, d, dev))
            if self._packed16 is None:
                self._packed16 = new
            else:
                self._packed16["conv1"].copy_(new["conv1"])
                self._packed16["proj"].copy_(new["proj"])
                for old_b, new_b in zip(self._packed16["blocks"], new["blocks"]):
                    for k, v in new_b.items():
                        old_b[k].copy_(v)
            self._packed16_key = key
        return self._packed16

    def packed_train(self):
        """packed() plus the transposed bf16 weights the data-gradient GEMMs read (dX = dY . W needs W^T K-major)."""
        pk = self.packed()
        if "proj_t" not in pk:
            pk["proj_t"] = ops.f32_to_bf16(self.proj.detach().to(torch.float32).contiguous())     # [d, C]
            for b in pk["blocks"]:
                for n in ("w_in", "w_out", "w_fc", "w_proj"):
                    b[n + "_t"] = ops.transpose_bf16(b[n])
        return pk

    def refresh_packed(self):
        """After an optimizer step: rewrite, IN PLACE, the bf16 GEMM copies (and their transposes) of every trainable
        weight -- LoRA factors are re-merged into in_proj / out_proj.  Pointer-stable, so CUDA graphs captured over the
        packed weights stay valid.  fp32 vectors (biases, LayerNorm affine, embeddings) alias the master parameters and
        need nothing."""
        if self._packed is None:
            return
        pk, d, P, dev = self._packed, self.width, self.patch_size, self.proj.device
        f32 = lambda t: t.detach().to(torch.float32).contiguous()
        if self.conv1.weight.requires_grad:
            w = f32(self.conv1.weight).reshape(d, 3 * P * P)
            if self.k_patch != 3 * P * P:
                wp = torch.zeros((d, self.k_patch), dtype=torch.float32, device=dev)
                wp[:, :3 * P * P] = w
                w = wp
            ops.f32_to_bf16(w, dst=pk["conv1"])
        if self.proj.requires_grad:
            ops.f32_to_bf16(f32(self.proj).t().contiguous(), dst=pk["proj"])
            if "proj_t" in pk:
                ops.f32_to_bf16(f32(self.proj), dst=pk["proj_t"])
        for blk, b in zip(self.transformer.resblocks, pk["blocks"]):
            if any(p.requires_grad for p in blk.attn.parameters()):
                merge_attn_weights(blk, d, dev, w_in=b["w_in"], w_out=b["w_out"])
                if "w_in_t" in b:
                    ops.transpose_bf16(b["w_in"], out=b["w_in_t"])
                    ops.transpose_bf16(b["w_out"], out=b["w_out_t"])
            for name, lin in (("w_fc", blk.mlp.c_fc), ("w_proj", blk.mlp.c_proj)):
                if lin.weight.requires_grad:
                    ops.f32_to_bf16(f32(lin.weight), dst=b[name])
                    if name + "_t" in b:
                        ops.transpose_bf16(b[name], out=b[name + "_t"])
        self._packed_key = self._version_key()

    refresh_lora_packed = refresh_packed

    # -- forward ----------------------------------------------------------------------------------------------------
    def forward_patches(self, patches, n_img):
        """patches: bf16 [n_img*G*G, k_patch] im2col rows (what ec_event2img's EC_OUT_BF16_PATCH writes).
        Returns fp32 [n_img, output_dim]."""
        if self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from . import train
            return train.encode_patches_autograd(self, patches, n_img)
        pk = self.packed()
        f16 = self.operand_dtype == torch.float16
        gray = patches.shape[1] == self.k_gray and self.k_gray != self.k_patch      # single-plane rows of the gray format
        if patches.dtype != self.operand_dtype:
            raise L.ECError(f"patch rows are {patches.dtype} but this tower computes with {self.operand_dtype} operands "
                            f"(ec_event2img out='{self.patch_fmt}'; EC_OPERANDS / visual.operand_dtype select the format)")
        p16 = self.packed_f16() if f16 else None
        d, G2, heads = self.width, self.grid ** 2, self.heads
        Ltok = G2 + 1
        M = n_img * Ltok
        dev = patches.device
        x0 = torch.empty((M, d), dtype=torch.float32, device=dev)       # tokens before ln_pre
        if gray:
            pg = self.packed_gray()
            ops.gemm_bf16(patches, pg["conv1"], None, "patch", out=x0, res=pg["pos"], row_map=G2, M=n_img * G2)
        else:
            ops.gemm_bf16(patches, p16["conv1"] if f16 else pk["conv1"], None, "patch", out=x0, res=pk["pos"], row_map=G2, M=n_img * G2)
        ops.cls_rows(x0, pk["cls"], pk["pos"], n_img, Ltok, d)
        # residual stream: fp16 like the reference's CUDA inference (test.py:26-29 keeps CLIP in fp16) or fp32
        fold = self.fold_ln and self.residual_dtype == torch.float16 and ops.gemm_stats_parts(d) <= 8
        x2 = None
        if fold and self.residual_split and d <= 1024:
            x2 = torch.empty((2, M, d), dtype=torch.float16, device=dev)
            ops.layernorm_f16x2(x0, *pk["ln_pre"], M, d, x2)
            x = x2[0]
        elif self.residual_dtype == torch.float16:
            x = torch.empty((M, d), dtype=torch.float16, device=dev)
            ops.layernorm(x0, *pk["ln_pre"], M, d, out_f16=x)
        else:
            x = torch.empty((M, d), dtype=torch.float32, device=dev)
            ops.layernorm(x0, *pk["ln_pre"], M, d, out_f32=x)
        del x0
        if fold:
            run_blocks_ln(x, p16["blocks"] if f16 else pk["blocks"], self.packed_ln(), n_img, Ltok, d, heads, x2=x2)
        else:
            run_blocks(x, pk["blocks"], n_img, Ltok, d, heads, w16=p16["blocks"] if f16 else None)
        cls = torch.empty((n_img, d), dtype=self.operand_dtype, device=dev)
        ops.layernorm(x, *pk["ln_post"], n_img, d, row_stride=Ltok * d, **(dict(out_f16=cls) if f16 else dict(out_bf16=cls)))
        return ops.gemm_bf16(cls, p16["proj"] if f16 else pk["proj"], None, "f32")

    def forward(self, x):
        """x: CUDA [n,3,224,224] float32 / bfloat16 images (the reference's data_dict['img'] rows)."""
        if x.dtype == torch.float16:
            raise L.ECError("pass float32 or bfloat16 images; the B200 encoder computes in bf16 with fp32 accumulation")
        n = x.shape[0]
        if n == 0:
            return torch.empty((0, self.output_dim), dtype=torch.float32, device=x.device)
        training = self.training and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        patches = ops.im2col(x.contiguous(), self.patch_size, self.k_patch, dtype=torch.bfloat16 if training else self.operand_dtype)
        return self.forward_patches(patches, n)


# text tower shapes of the released models: (width, layers, heads, context, vocab)
TEXT = {
    "ViT-B/32": (512, 12, 8, 77, 49408),
    "ViT-B/16": (512, 12, 8, 77, 49408),
    "ViT-L/14": (768, 12, 12, 77, 49408),
    "ViT-tiny/32": (64, 2, 1, 16, 97),
    "ViT-tiny/16": (64, 2, 1, 16, 97),
}


class CLIP(nn.Module):
    """Image tower + logit scale (+ optional text tower).  With text=True the module also
    owns openai-CLIP's text parameters under their original names (token_embedding, positional_embedding,
    transformer.resblocks.*, ln_final, text_projection) and encode_text runs them through the same kernels with the
    causal tcgen05 attention.  The BPE tokenizer is not part of this (vocabulary absent offline): callers pass ids."""

    def __init__(self, arch, text=False):
        super().__init__()
        patch, width, layers, heads, embed = ARCHS[arch]
        self.arch = arch
        self.visual = VisionTransformer(224, patch, width, layers, heads, embed)
        self.logit_scale = nn.Parameter(torch.ones([]) * math.log(1 / 0.07))
        self.has_text = text
        if text:
            w, tl, th, ctx, vocab = TEXT[arch]
            self.context_length, self.vocab_size = ctx, vocab
            self.token_embedding = nn.Embedding(vocab, w)
            self.positional_embedding = nn.Parameter(0.01 * torch.randn(ctx, w))
            self.transformer = Transformer(w, tl, th)
            self.ln_final = nn.LayerNorm(w)
            self.text_projection = nn.Parameter(w ** -0.5 * torch.randn(w, embed))
            self._tpacked, self._tpacked_key = None, None

    @property
    def dtype(self):
        return self.visual.conv1.weight.dtype

    def encode_image(self, image):
        return self.visual(image)

    def _text_params(self):
        return [self.token_embedding.weight, self.positional_embedding, self.text_projection] + \
            list(self.transformer.parameters()) + list(self.ln_final.parameters())

    def encode_text(self, text):
        """text: int [n, context_length] token ids (clip.tokenize output) -> fp32 [n, embed_dim].
        Features are read at the end-of-text position = argmax of the ids, as in openai-CLIP."""
        if not self.has_text:
            raise NotImplementedError("this CLIP object was built without the text tower (CLIP(arch, text=True)); "
                                      "pass precomputed text features to the classifier instead")
        dev = self.text_projection.device
        key = tuple((p.data_ptr(), p._version) for p in self._text_params())
        if self._tpacked is None or key != self._tpacked_key:
            f32 = lambda t: t.detach().to(torch.float32).contiguous()
            w = self.transformer.width
            self._tpacked = dict(table=f32(self.token_embedding.weight), pos=f32(self.positional_embedding),
                                 blocks=pack_blocks(self.transformer.resblocks, w, dev),
                                 ln=(f32(self.ln_final.weight), f32(self.ln_final.bias)),
                                 proj=ops.f32_to_bf16(f32(self.text_projection).t().contiguous()))
            self._tpacked_key = key
        pk = self._tpacked
        n, ctx = text.shape
        w, heads = self.transformer.width, self.transformer.heads
        x = ops.embed_tokens(pk["table"], text.to(device=dev, dtype=torch.int32).contiguous(), pk["pos"])
        run_blocks(x, pk["blocks"], n, ctx, w, heads, causal=True)
        eot = (torch.arange(n) * ctx + text.cpu().argmax(dim=-1)).to(torch.int32)
        rows = ops.gather_rows(x, eot.to(dev), n)
        feat = torch.empty((n, w), dtype=torch.bfloat16, device=dev)
        ops.layernorm(rows, *pk["ln"], n, w, out_bf16=feat)
        return ops.gemm_bf16(feat, pk["proj"], None, "f32")


def init_weights_(model, seed=0, logit_scale=100.0):
    """Seeded random init with CLIP's published scales (no pretrained checkpoints offline).  Mirrors the oracle's
    init so both sides can be built from one seed without sharing tensors."""
    g = torch.Generator().manual_seed(seed)
    v = model.visual
    d, layers = v.width, v.transformer.layers
    attn_std, proj_std, fc_std = d ** -0.5, (d ** -0.5) * ((2 * layers) ** -0.5), (2 * d) ** -0.5

    def rn(t, std, mean=0.0):
        with torch.no_grad():
            t.copy_((torch.randn(t.shape, generator=g) * std + mean).to(t.device))

    rn(v.conv1.weight, (3 * v.patch_size ** 2) ** -0.5)
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


def _transform(n_px=224):
    """CLIP's preprocess (openai-CLIP `_transform`), for callers that still feed PIL images."""
    import torchvision.transforms as T
    return T.Compose([T.Resize(n_px, interpolation=T.InterpolationMode.BICUBIC), T.CenterCrop(n_px),
                      lambda im: im.convert("RGB"), T.ToTensor(),
                      T.Normalize((0.48145466, 0.4578275, 0.40821073), (0.26862954, 0.26130258, 0.27577711))])


def load(name, device="cuda", seed=0, state_dict=None, text=True):
    """clip.load(arch, device) -> (model, preprocess)  (test.py:26, train.py:26).
    Checkpoints cannot be downloaded offline: weights are seeded random unless `state_dict` is given."""
    model = CLIP(name, text=text)
    init_weights_(model, seed)
    if state_dict is not None:
        # openai-CLIP checkpoints carry non-parameter entries and (when text=False here) the text tower: those may be
        # unexpected.  A MISSING key would silently leave seeded-random weights in place, so it is an error.
        res = model.load_state_dict(state_dict, strict=False)
        if res.missing_keys:
            raise L.ECError("clip.load: checkpoint lacks %d parameter(s) of %s, e.g. %s"
                            % (len(res.missing_keys), name, ", ".join(res.missing_keys[:4])))
    return model.to(device).eval(), _transform(224)


def tokenize(texts, context_length=77):
    raise NotImplementedError("the BPE vocabulary file is not available offline; pass token ids to encode_text, a "
                              "tokenizer callable in clip_dict['tokenizer'], or text features in clip_dict['text_feats']")
