"""Fine-tune step of FTCLIPClassifier on the B200 path (BASELINE config 5).

The reference trains through `model.visual` with autograd (models/clip_cls_ft.py:214-269, LoRA factors of
models/lora.py:101-159 / 14-57 plus the prompt-tuned text features), DDP averaging the gradients and
torch.optim.Adam with two learning rates (method.py:150-191).  Here the forward keeps the activations it needs, the
backward is an explicit chain of the library's kernels (tcgen05 GEMMs for data and weight gradients, mma.sync
attention backward, LayerNorm / QuickGELU / normalize / cross-entropy backward kernels), gradients live in ONE flat fp32
buffer that is all-reduced with a single NCCL call, and Adam is one kernel launch per parameter group.

Two ways in:
  * torch.autograd compatible: `FTCLIPClassifier.forward` in training mode builds a graph of three autograd
    Functions (encoder, logit head, loss) whose backward methods run the kernels below, so the reference's
    `loss.backward(); optimizer.step()` loop works unchanged.
  * `FineTuner`: the fused step `events -> frames -> forward -> loss -> backward -> all-reduce -> Adam` without autograd.

Trainable sets: whatever clip_cls_ft.py:45-80 leaves with requires_grad -- LoRA factors on q/k/v/o
(`lora='qv-r' | 'qkv-r' | 'qkvo-r' | int`, configs/ftclip/*lora16.py), the only_conv1 / only_bias / only_ln / only_cls_fc /
only_cls_token subsets, or the whole image tower (`lora=-1`, configs/ftclip/*vitb16.py) -- plus the text features of
the 'text-identity' adapter.  Every parameter of model.visual has a gradient kernel path here.
"""
import torch

from . import ops
from . import _lib as L


# ------------------------------------------------------------------------------------------------ trainable set
def trainable_params(vis):
    """Parameters of the image tower that want a gradient, in named_parameters order: whatever clip_cls_ft.py:45-80 left
    trainable (LoRA factors, conv1 / bias / LayerNorm / proj / class token subsets, or the whole tower)."""
    return [p for _, p in vis.named_parameters() if p.requires_grad]


def _req(p):
    return p is not None and p.requires_grad


def _block_needs(blk):
    return any(p.requires_grad for p in blk.parameters())


# ------------------------------------------------------------------------------------------------ encoder forward/backward
def encoder_forward(vis, patches, n_img):
    """Training forward of VisionTransformer.forward_patches: same kernels, out-of-place residual updates, QuickGELU on the
    stored bf16 pre-activation.  Returns (feats fp32 [n_img, C], ctx)."""
    pk = vis.packed_train()
    d, G2, heads = vis.width, vis.grid ** 2, vis.heads
    Ltok = G2 + 1
    M, dev = n_img * Ltok, patches.device
    bf = lambda *shape: torch.empty(shape, dtype=torch.bfloat16, device=dev)
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32, device=dev)
    x0 = f32(M, d)
    ops.gemm_bf16(patches, pk["conv1"], None, "patch", out=x0, res=pk["pos"], row_map=G2, M=n_img * G2)
    ops.cls_rows(x0, pk["cls"], pk["pos"], n_img, Ltok, d)
    x = f32(M, d)
    ops.layernorm(x0, *pk["ln_pre"], M, d, out_f32=x)
    bottom = any(_req(p) for p in (vis.conv1.weight, vis.class_embedding, vis.positional_embedding, vis.ln_pre.weight,
                                   vis.ln_pre.bias))
    saved = []
    for b in pk["blocks"]:
        h1, qkv, att = bf(M, d), bf(M, 3 * d), bf(M, d)
        ops.layernorm(x, *b["ln1"], M, d, out_bf16=h1)
        ops.gemm_bf16(h1, b["w_in"], b["b_in"], "bf16", out=qkv)
        lse = None
        if Ltok <= 256:          # the tensor-memory backward needs the row log-sum-exp of the forward
            lse = ops.attention_fwd_lse(qkv, att, n_img, Ltok, heads)
        else:
            ops.attention(qkv, att, n_img, Ltok, heads)
        x2 = f32(M, d)
        ops.gemm_bf16(att, b["w_out"], b["b_out"], "f32_resadd", out=x2, res=x)
        h2, a = bf(M, d), bf(M, 4 * d)
        ops.layernorm(x2, *b["ln2"], M, d, out_bf16=h2)
        ops.gemm_bf16(h2, b["w_fc"], b["b_fc"], "bf16", out=a)
        g = ops.quickgelu(a)
        x3 = f32(M, d)
        ops.gemm_bf16(g, b["w_proj"], b["b_proj"], "f32_resadd", out=x3, res=x2)
        saved.append((x, h1, qkv, att, x2, h2, a, g, lse))
        x = x3
    cls = bf(n_img, d)
    cls32 = f32(n_img, d) if _req(vis.proj) else None
    ops.layernorm(x, *pk["ln_post"], n_img, d, row_stride=Ltok * d, out_bf16=cls, out_f32=cls32)
    feats = ops.gemm_bf16(cls, pk["proj"], None, "f32")
    return feats, dict(pk=pk, saved=saved, x_last=x, cls=cls, cls32=cls32, n_img=n_img, Ltok=Ltok, d=d, heads=heads,
                       x0=x0 if bottom else None, patches=patches if _req(vis.conv1.weight) else None)


def encoder_backward(vis, ctx, d_feats, dest=None):
    """d_feats fp32 [n_img, C] -> {id(param): fp32 gradient} for every trainable parameter of the image tower.
    dest(param) may return the tensor a gradient is to be written into (FineTuner's flat-buffer views).
    Frees ctx['saved'] as it goes; stops at the lowest block that has anything trainable."""
    pk, saved = ctx["pk"], ctx["saved"]
    n_img, Ltok, d, heads = ctx["n_img"], ctx["Ltok"], ctx["d"], ctx["heads"]
    G2 = Ltok - 1
    M, dev = n_img * Ltok, d_feats.device
    blocks = list(vis.transformer.resblocks)
    bottom = ctx["x0"] is not None
    first = 0 if bottom else next((i for i, blk in enumerate(blocks) if _block_needs(blk)), len(blocks))
    grads = {}
    f32w = lambda t: t.detach().to(torch.float32).contiguous()
    to = (lambda p: None) if dest is None else dest

    def wgrad(p, dy_bf16, x_bf16):                 # dW[out, in] = dY^T X over the tokens: split-K tcgen05 GEMM, MN-major operands
        out = to(p)
        view = None if out is None else out.view(p.shape[0], -1)
        g = ops.gemm_bf16_tn(dy_bf16, x_bf16, out=view)
        grads[id(p)] = g.view(p.shape)
        return g

    def bgrad(p, dy):                              # bias gradient: column sums over the tokens
        grads[id(p)] = ops.colsum(dy, out=to(p))

    def lngrad(ln, xin, dy, Mrows, x_stride=None):
        if _req(ln.weight) or _req(ln.bias):
            dg, db = ops.layernorm_param_grad(xin, dy, Mrows, d, x_stride=x_stride,
                                              dgamma=to(ln.weight) if _req(ln.weight) else None,
                                              dbeta=to(ln.bias) if _req(ln.bias) else None)
            if _req(ln.weight):
                grads[id(ln.weight)] = dg
            if _req(ln.bias):
                grads[id(ln.bias)] = db

    def factor_grads(dW, names, facs):             # dUp = dW . down^T,  dDown = up^T . dW   (W_eff = W + up . down)
        fac = [(f32w(f[0]), f32w(f[1])) if f is not None else None for f in facs]
        dst = [(to(f[0]), to(f[1])) if f is not None else None for f in facs]
        for f, g in zip(facs, ops.lora_grad(dW, d, fac, dst)):
            if g is not None:
                grads[id(f[0])], grads[id(f[1])] = g

    # feats = ln_post(x[cls rows]) @ proj
    dfb = ops.f32_to_bf16(d_feats.contiguous())
    if _req(vis.proj):                             # d_proj[d, C] = cls^T d_feats (n_img rows: tiny, fp32 SIMT)
        grads[id(vis.proj)] = ops.mm_f32(ctx["cls32"], d_feats.contiguous(), trans_a=True, out=to(vis.proj))
    d_cls = ops.gemm_bf16(dfb, pk["proj_t"], None, "f32")                       # [n_img, d]
    lngrad(vis.ln_post, ctx["x_last"], d_cls, n_img, x_stride=Ltok * d)
    dx = torch.zeros((M, d), dtype=torch.float32, device=dev)
    ops.layernorm_bwd(ctx["x_last"], d_cls, pk["ln_post"][0], n_img, d, x_stride=Ltok * d, dx=dx, dx_stride=Ltok * d)
    dxb = ops.f32_to_bf16(dx)
    for i in range(len(saved) - 1, first - 1, -1):
        b, blk = pk["blocks"][i], blocks[i]
        x, h1, qkv, att, x2, h2, a, g, lse = saved[i]
        saved[i] = None
        W, lora_in, ib, oW, ob, lora_out = _attn_parts(blk.attn)
        fc, pj = blk.mlp.c_fc, blk.mlp.c_proj
        # MLP:  x3 = x2 + c_proj(QuickGELU(c_fc(ln_2(x2))));  dxb = bf16 copy of dx (written by the previous LayerNorm backward)
        if _req(pj.weight):
            wgrad(pj.weight, dxb, g)
        if _req(pj.bias):
            bgrad(pj.bias, dx)
        dg = ops.gemm_bf16(dxb, b["w_proj_t"], None, "bf16")                 # [M, 4d]
        da = ops.quickgelu_bwd(a, dg, out=dg)
        if _req(fc.weight):
            wgrad(fc.weight, da, h2)
        if _req(fc.bias):
            bgrad(fc.bias, da)
        dh2 = ops.gemm_bf16(da, b["w_fc_t"], None, "f32")                    # [M, d]
        lngrad(blk.ln_2, x2, dh2, M)
        dx2b = torch.empty((M, d), dtype=torch.bfloat16, device=dev)
        dx2 = ops.layernorm_bwd(x2, dh2, b["ln2"][0], M, d, acc=dx, dx_bf16=dx2b)
        del dg, da, dh2, a, g, h2
        # attention:  x2 = x + out_proj(attn(in_proj(ln_1(x))))
        d_att = ops.gemm_bf16(dx2b, b["w_out_t"], None, "bf16")              # [M, d]
        if _req(ob):
            bgrad(ob, dx2)
        if _req(oW):
            wgrad(oW, dx2b, att)
        elif lora_out is not None and (_req(lora_out.lora_up.weight) or _req(lora_out.lora_down.weight)):
            dW = ops.gemm_bf16_tn(dx2b, att)                                                          # [d_out, d_in]
            factor_grads(dW, "o", [(lora_out.lora_up.weight, lora_out.lora_down.weight)])
        dqkv = ops.attention_bwd(qkv, att, d_att, n_img, Ltok, heads, lse=lse)
        if _req(ib):
            bgrad(ib, dqkv)
        facs = None
        if lora_in is not None:
            facs = [(getattr(lora_in, f"lora_up_{n}"), getattr(lora_in, f"lora_down_{n}"))
                    if hasattr(lora_in, f"lora_up_{n}") and (_req(getattr(lora_in, f"lora_up_{n}")) or
                                                             _req(getattr(lora_in, f"lora_down_{n}"))) else None for n in "qkv"]
        if _req(W):
            wgrad(W, dqkv, h1)
        elif facs is not None and any(f is not None for f in facs):
            dW = ops.gemm_bf16_tn(dqkv, h1)                                                           # [3d, d]
            factor_grads(dW, "qkv", facs)
        need_ln1 = _req(blk.ln_1.weight) or _req(blk.ln_1.bias)
        if i > first or bottom or need_ln1:
            dh1 = ops.gemm_bf16(dqkv, b["w_in_t"], None, "f32")
            lngrad(blk.ln_1, x, dh1, M)
            if i > first or bottom:                                          # nothing trainable below the first block otherwise
                dx = ops.layernorm_bwd(x, dh1, b["ln1"][0], M, d, acc=dx2, dx_bf16=dxb)
    ctx["saved"] = None
    if bottom:
        # x = ln_pre(x0),  x0[img, 0] = class_embedding + pos[0],  x0[img, 1 + t] = conv1(patch t) + pos[1 + t]
        x0 = ctx["x0"]
        lngrad(vis.ln_pre, x0, dx, M)
        dx0 = ops.layernorm_bwd(x0, dx, pk["ln_pre"][0], M, d)
        if _req(vis.positional_embedding) or _req(vis.class_embedding):
            dpos = ops.colsum(dx0.view(n_img, Ltok * d), out=to(vis.positional_embedding).view(-1)
                              if _req(vis.positional_embedding) and to(vis.positional_embedding) is not None else None,
                              n_part=1 if n_img < 128 else None).view(Ltok, d)
            if _req(vis.positional_embedding):
                grads[id(vis.positional_embedding)] = dpos
            if _req(vis.class_embedding):          # the class token sits in row 0 of every image next to pos[0]
                dst = to(vis.class_embedding)
                grads[id(vis.class_embedding)] = dpos[0].clone() if dst is None else dst.copy_(dpos[0])
        if _req(vis.conv1.weight):                 # conv1 as a GEMM over im2col rows: dW[d, 3PP] = d_x0[patch rows]^T patches
            w = vis.conv1.weight
            kp, k = vis.k_patch, 3 * vis.patch_size ** 2
            dyp = ops.patch_rows_bf16(dx0, n_img, G2, d)
            gw = ops.gemm_bf16_tn(dyp, ctx["patches"])                                                # [d, k_patch]
            gw = gw[:, :k].reshape(w.shape)
            dst = to(w)
            grads[id(w)] = gw.contiguous() if dst is None else dst.copy_(gw)
    return grads


def _attn_parts(attn):
    from .clip import _attn_weights
    return _attn_weights(attn)


# ------------------------------------------------------------------------------------------------ logit head
def head_forward(feats, plan, text_raw, scale, agg):
    """clip_cls_ft.py:224-248 with the kernels of the inference head.  feats fp32 [n_valid, C]; text_raw fp32 [K, C]
    (re-normalised every forward, :163).  Returns (out_dict, ctx)."""
    B, T = plan["B"], plan["T"]
    slots = feats if plan["row_of_slot"] is None else ops.gather_rows(feats, plan["row_of_slot"], B * T)
    slots = slots.contiguous()
    that = ops.l2norm_rows(text_raw)
    full, logits, probs, top = ops.head(slots, plan["valid_u8"], that, B, T, scale, True, agg)
    out = {"full_logits": full, "valid_masks": plan["valid_dev"], "logits": logits, "probs": probs,
           "top5_logits": top[:, 0], "top5_probs": top[:, 1]}
    return out, dict(slots=slots, that=that, text_raw=text_raw, plan=plan, scale=scale)


def head_backward(ctx, d_full):
    """d_full fp32 [B,T,K] -> (d_feats [n_valid, C], d_text_raw [K, C])."""
    plan, slots, that, scale = ctx["plan"], ctx["slots"], ctx["that"], ctx["scale"]
    B, T = plan["B"], plan["T"]
    S, K = B * T, that.shape[0]
    dz = d_full.reshape(S, K)
    fhat = ops.l2norm_rows(slots)                                   # zero rows (invalid views) stay zero
    d_fhat = ops.mm_f32(dz, that, alpha=scale)                      # [S, C]
    d_that = ops.mm_f32(dz, fhat, trans_a=True, alpha=scale)        # [K, C]; invalid rows of fhat are zero
    d_slots = ops.l2norm_rows_bwd(slots, d_fhat, plan["valid_u8"])
    d_text = ops.l2norm_rows_bwd(ctx["text_raw"], d_that)
    if plan["row_of_slot"] is None:
        return d_slots, d_text
    return ops.gather_rows(d_slots, plan["slot_of_row"], plan["n_valid"]), d_text


def add_slot_of_row(plan, dev):
    """Inverse of row_of_slot (valid row -> slot), needed to route feature gradients back."""
    if plan["row_of_slot"] is not None and "slot_of_row" not in plan:
        ros = plan["row_of_slot"]
        plan["slot_of_row"] = (ros >= 0).nonzero().squeeze(1).to(torch.int32).to(dev)
    return plan


# ------------------------------------------------------------------------------------------------ autograd plumbing
class _EncoderFn(torch.autograd.Function):
    """autograd node whose backward is encoder_backward; inputs after `n_img` are the trainable parameters of the tower."""

    @staticmethod
    def forward(fctx, vis, patches, n_img, *params):
        feats, ctx = encoder_forward(vis, patches, n_img)
        fctx.vis, fctx.ctx, fctx.params = vis, ctx, params
        return feats

    @staticmethod
    def backward(fctx, d_feats):
        grads = encoder_backward(fctx.vis, fctx.ctx, d_feats.contiguous())
        return (None, None, None, *[grads[id(p)].to(p.dtype).reshape(p.shape) for p in fctx.params])


def encode_patches_autograd(vis, patches, n_img):
    return _EncoderFn.apply(vis, patches, n_img, *trainable_params(vis))


class _HeadFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, feats, text_raw, plan, scale, agg):
        out, ctx = head_forward(feats.detach().float().contiguous(), plan, text_raw.detach().float().contiguous(), scale, agg)
        fctx.ctx = ctx
        fctx.mark_non_differentiable(out["logits"], out["probs"], out["top5_logits"], out["top5_probs"])
        return out["full_logits"], out["logits"], out["probs"], out["top5_logits"], out["top5_probs"]

    @staticmethod
    def backward(fctx, d_full, *unused):
        d_feats, d_text = head_backward(fctx.ctx, d_full.contiguous())
        return d_feats, d_text, None, None, None


class _CELossFn(torch.autograd.Function):
    """F.cross_entropy(aggregate(full_logits), labels) with the gradient produced by the same kernel."""

    @staticmethod
    def forward(fctx, full_logits, valid_u8, labels_i32, agg):
        if agg == "probs":        # use_probs_loss: nll of the view-averaged probabilities (clip_cls_ft.py:265-267)
            _, loss, d_full = ops.probs_loss_bwd(full_logits.detach().contiguous(), valid_u8, labels_i32)
        else:
            _, loss, d_full = ops.ce_loss_bwd(full_logits.detach().contiguous(), valid_u8, labels_i32, agg)
        fctx.d_full = d_full
        return loss.reshape(())

    @staticmethod
    def backward(fctx, g):
        return _scale(fctx.d_full, g), None, None, None


def _scale(t, g):
    """t * g for a scalar upstream gradient (exactly 1.0 for a plain loss.backward(), where t is returned as is)."""
    gv = float(g)
    if gv == 1.0:
        return t
    return ops.blend(t.reshape(-1), torch.zeros_like(t).reshape(-1), gv).reshape(t.shape)


def train_forward(model, feats, plan):
    """Head of FTCLIPClassifier.forward in training mode: returns the reference's out_dict with `full_logits` attached
    to the autograd graph (clip_cls_ft.py:224-256)."""
    add_slot_of_row(plan, feats.device)
    text_raw = model.text_feats if model.prompt_tuning else model.get_text_feats()
    full, logits, probs, t5l, t5p = _HeadFn.apply(feats, text_raw, plan, model.logit_scale, model.agg_func)
    return {"full_logits": full, "valid_masks": plan["valid_dev"], "logits": logits, "probs": probs,
            "top5_logits": t5l, "top5_probs": t5p, "_plan": plan}


def train_loss(model, data_dict, out_dict):
    """calc_train_loss in training mode (clip_cls_ft.py:258-269): cross-entropy of the aggregated logits (use_logits_loss) or
    nll of the view-averaged probabilities (use_probs_loss)."""
    plan = out_dict["_plan"]
    labels = data_dict["label"].to(device=out_dict["full_logits"].device, dtype=torch.int32).contiguous()
    kind = model.agg_func if model.use_logits_loss else "probs"
    return {"ce_loss": _CELossFn.apply(out_dict["full_logits"], plan["valid_u8"], labels, kind)}


# ------------------------------------------------------------------------------------------------ fused trainer
class FineTuner:
    """The whole fine-tune step without autograd: forward, loss, backward into one flat gradient buffer, ONE all-reduce
    (NCCL when torch.distributed is initialised; the reference's DDP averages the same gradients), Adam with the
    reference's two learning rates (method.py:155-182: `lr` for everything outside model.visual, `clip_lr` inside).

        ft = FineTuner(model, lr=2e-5, clip_lr=2e-5)
        loss = ft.step(events, offsets, labels)            # device events, host offsets, device/host labels
    """

    def __init__(self, model, lr, clip_lr=None, betas=(0.9, 0.999), eps=1e-8, process_group=None):
        if not getattr(model, "prompt_tuning", False):
            raise NotImplementedError("FineTuner expects the 'text-identity' adapter (prompt-tuned text features)")
        self.model, self.vis = model, model.model.visual
        self.lr, self.clip_lr = lr, (lr if clip_lr is None else clip_lr)
        self.betas, self.eps, self.pg = betas, eps, process_group
        # group 0: outside model.visual (text features), group 1: trainable parameters of the image tower -- each group is
        # one contiguous span of the flat buffers
        self.group0 = [model.text_feats]
        self.group1 = trainable_params(self.vis)

        from .dist import FlatParams
        self.flat = FlatParams([self.group0, self.group1])
        self.flat_p, self.flat_g = self.flat.flat_p, self.flat.flat_g
        # What DDP does at construction (the reference trains under nerv's DDP wrapper): every rank starts from rank 0's
        # parameters.  LoRA `lora_down` is drawn from the global RNG at injection, so ranks seeded differently would
        # otherwise average gradients of different weights and stay diverged.
        self.flat.broadcast(self.pg)
        self.m, self.v = torch.zeros_like(self.flat_g), torch.zeros_like(self.flat_g)
        self.vis.invalidate_packed()
        self.t = 0
        self.last = {}

    def _grad_view(self, p):
        return self.flat.grad_view(p)

    def forward_backward(self, events, offsets, labels, sel=None):
        model = self.model
        plan = model.plan_to_device(model.plan_events(offsets, sel), events.device)
        add_slot_of_row(plan, events.device)
        labels = labels.to(device=events.device, dtype=torch.int32).contiguous()
        return self.device_forward_backward(events, plan, labels)

    def device_forward_backward(self, events, plan, labels):
        """Device part of the step (kernel launches only, capturable in a CUDA graph): frames, forward, loss, backward
        into the flat gradient buffer.  Returns the mean loss as a device scalar [1]."""
        model = self.model
        with torch.no_grad():
            fe = model.event_frontend
            patches, st, _ = ops.event2img(events, plan["frames"], fe.resolution, plan["n_valid"], fe.count_non_zero,
                                           fe.background_mask, out="patch", patch=self.vis.patch_size, ldk=self.vis.k_patch)
            feats, ctx = encoder_forward(self.vis, patches, plan["n_valid"])
            out, hctx = head_forward(feats, plan, model.text_feats.detach(), model.logit_scale, model.agg_func)
            if model.use_logits_loss:
                _, loss, d_full = ops.ce_loss_bwd(out["full_logits"], plan["valid_u8"], labels, model.agg_func)
            else:     # use_probs_loss (clip_cls_ft.py:265-267)
                _, loss, d_full = ops.probs_loss_bwd(out["full_logits"], plan["valid_u8"], labels)
            d_feats, d_text = head_backward(hctx, d_full)
            encoder_backward(self.vis, ctx, d_feats, dest=self.flat.grad_view)
            self._grad_view(model.text_feats).copy_(d_text)
        self.last = {"out": out, "status": st}
        return loss

    def allreduce(self):
        self.flat.average_gradients(self.pg)

    def optimizer_step(self, lr=None, clip_lr=None, refresh=True):
        self.t += 1
        lr = self.lr if lr is None else lr
        clip_lr = self.clip_lr if clip_lr is None else clip_lr
        for (lo, hi), rate in zip(self.flat.spans, (lr, clip_lr)):
            if hi > lo:
                ops.adam(self.flat_p[lo:hi], self.flat_g[lo:hi], self.m[lo:hi], self.v[lo:hi], rate, self.t, self.betas, self.eps)
        # ec_adam writes through raw pointers: version counters do not move, so tell the tower (LayerNorm-folded operands,
        # captured inference graphs) that its weights changed
        self.vis.mark_weights_changed()
        if refresh:
            self.refresh_weights()

    def refresh_weights(self):
        """bf16 GEMM copies of the updated master weights, rewritten in place (pointer-stable, graph-capturable)."""
        self.vis.refresh_packed()

    def step(self, events, offsets, labels, sel=None, lr=None, clip_lr=None):
        loss = self.forward_backward(events, offsets, labels, sel)
        self.allreduce()
        self.optimizer_step(lr, clip_lr)
        return loss

    def check_status(self):
        """Raises what the reference's numpy path raises for the events of the last step (ValueError for coordinates
        outside the sensor, datasets/vis.py:9-14).  Synchronises on the status word; call it once per epoch / at will."""
        st = self.last.get("status")
        if st is not None:
            ops.raise_on_status(st)
