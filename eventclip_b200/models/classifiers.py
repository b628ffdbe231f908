"""Zero-shot / few-shot / fine-tuned EventCLIP classifiers on the B200 path.

Drop-in surface of the reference's models/clip_cls.py and models/clip_cls_ft.py: same constructor dicts
(`clip_dict`, `adapter_dict`, `loss_dict`), `forward(data_dict) -> {'full_logits','valid_masks','logits','probs'}`,
`get_text_feats`, `get_img_feats`, `calc_train_loss`, `calc_eval_loss`, `.dtype`, `.device`, `train()` keeping CLIP in
eval, and the CLIP-stripping `state_dict` / `load_state_dict`.  Differences, all additive:
  * `data_dict` may carry packed events (`'events'` float32 [sum E,4] + `'event_offsets'` int64 [B+1], optional
    `'sel_idx'`) instead of `'img'`; the frames then go straight from the fused event kernel into the patch GEMM.
  * `clip_dict['text_feats']` may carry precomputed text features (the tokenizer vocabulary is absent offline).
  * out_dict also holds `'top5_logits'` / `'top5_probs'` (int32 [B,5]) computed by the head kernel.
The arithmetic is in libeventclip_b200.so; nn.Module is the parameter container.
"""
import copy

import torch
import torch.nn as nn
from torch.nn import functional as F

from .. import ops
from .. import _lib as L
from ..datasets.event2img import Event2Image
from .adapter import IdentityAdapter, TransformerAdapter
from .lora import inject_trainable_lora


class _CLIPClassifierBase(nn.Module):
    normalize_img_feats = False

    def __init__(self, clip_dict, loss_dict):
        super().__init__()
        self.clip_dict = clip_dict
        self.loss_dict = loss_dict
        self.event_frontend = None
        self._build_clip()
        self._build_loss()

    # ---- construction ---------------------------------------------------------------------------------------
    def _freeze_clip(self, model):
        for p in model.parameters():
            p.requires_grad = False

    def _build_clip(self):
        model = self.clip_dict["clip_model"]
        self._freeze_clip(model)
        self.model = model.eval()
        self.logit_scale = model.logit_scale.exp().item()      # clip_cls.py:44
        self.prompt = self.clip_dict["prompt"]
        self.class_names = self.clip_dict["class_names"]
        tf = self.clip_dict.get("text_feats", None)
        self.text_feats = None if tf is None else tf.detach().to(model.logit_scale.device)
        self.agg_func = self.clip_dict["agg_func"]
        assert self.agg_func in ["sum", "mean", "max"]          # clip_cls.py:53

    def _build_loss(self):
        self.use_logits_loss = self.loss_dict["use_logits_loss"]
        self.use_probs_loss = self.loss_dict["use_probs_loss"]
        assert int(self.use_logits_loss) + int(self.use_probs_loss) == 1   # clip_cls.py:58

    def attach_event_frontend(self, quantize_args, resolution, max_n):
        """Tells the classifier how to turn packed events into views (the reference keeps this in the Dataset)."""
        self.event_frontend = Event2Image(quantize_args, resolution, max_n)
        return self

    # ---- features -------------------------------------------------------------------------------------------
    def _same_class_names(self, class_names):
        return all(c1 == c2 for c1, c2 in zip(class_names, self.class_names))

    def _clip_text_feats(self, class_names=None):
        """clip_cls.py:64-93: cached text features, else tokenize + encode_text + normalise."""
        if (class_names is None or self._same_class_names(class_names)) and self.text_feats is not None:
            return self.text_feats
        names = self.class_names if class_names is None else class_names
        from .. import clip
        prompts = [self.prompt.format(c.lower().replace("_", " ")) for c in names]
        tokenize = self.clip_dict.get("tokenizer", None) or clip.tokenize          # clip.tokenize raises offline (no BPE vocab)
        tokens = torch.cat([tokenize(p) for p in prompts]).to(self.device)
        feats = ops.l2norm_rows(self.model.encode_text(tokens).float().contiguous())
        if class_names is None or self._same_class_names(class_names):
            self.text_feats = feats
        return feats

    def get_text_feats(self, class_names=None):
        return self._clip_text_feats(class_names)

    def get_img_feats(self, imgs):
        """imgs: CUDA [N,3,224,224] -> fp32 [N,C] (model.encode_image, clip_cls.py:101)."""
        with torch.no_grad():
            return self.model.encode_image(imgs)

    def _adapt(self, full_feats, valid_masks):
        return full_feats

    # The events route is split into a host plan and a pure device part so that the latter can be captured in a
    # CUDA graph (eventclip_b200/graph.py): nothing in device_forward touches the host.
    def plan_events(self, offsets, sel=None):
        """Host side of the fused route: chunking + view selection (ec_plan_frames) and the slot map."""
        if self.event_frontend is None:
            raise L.ECError("call attach_event_frontend(quantize_args, resolution, max_n) before passing events")
        fe = self.event_frontend
        offsets = offsets.cpu().numpy() if isinstance(offsets, torch.Tensor) else offsets
        if sel is None:
            sel = fe.draw_selection(offsets)
        elif isinstance(sel, torch.Tensor):
            sel = sel.cpu().numpy()
        frames, valid, chunks, n_valid = ops.plan_frames(offsets, fe.N, fe.max_imgs, sel=sel, compact=True)
        B, T = valid.shape
        vflat = valid.reshape(-1)
        row_of_slot = None
        if not bool(vflat.all()):
            row_of_slot = torch.full((B * T,), -1, dtype=torch.int32)
            row_of_slot[vflat] = torch.arange(n_valid, dtype=torch.int32)
        return dict(frames=frames, valid=valid, valid_u8=valid.to(torch.uint8), row_of_slot=row_of_slot,
                    n_valid=n_valid, B=B, T=T)

    @staticmethod
    def plan_to_device(plan, dev):
        pin = lambda t: t if t.is_pinned() else t.pin_memory()
        d = dict(plan)
        d["frames"] = plan["frames"].to(dev, non_blocking=True)
        d["valid_u8"] = pin(plan["valid_u8"]).to(dev, non_blocking=True)
        d["valid_dev"] = d["valid_u8"].bool()
        if plan["row_of_slot"] is not None:
            d["row_of_slot"] = pin(plan["row_of_slot"]).to(dev, non_blocking=True)
        return d

    def device_forward(self, events, plan, status=None):
        """events: CUDA float32 [sum E, 4]; plan: output of plan_to_device.  Enqueues kernels only."""
        fe, visual = self.event_frontend, self.model.visual
        patches, st, _ = ops.event2img(events, plan["frames"], fe.resolution, plan["n_valid"], fe.count_non_zero,
                                       fe.background_mask, out=visual.patch_fmt, patch=visual.patch_size, ldk=visual.patch_ldk,
                                       status=status)
        self._last_status, self._last_patches = st, patches     # kept for status checks / parity checks of the frames
        feats = visual.forward_patches(patches, plan["n_valid"])
        out = self._head(feats, plan)
        out["status"] = st          # device status word of the event kernel (EC_STATUS_*); see check_status()
        return out

    def check_status(self):
        """The reference's numpy path raises ValueError on coordinates outside the sensor (datasets/vis.py:9-14).  The fused
        route only sets a bit in a device word; this reads it (synchronising) and raises the same exception.  Called by
        dist.AccuracyMeter before it reports, by GraphedClassifier.stream() with every batch it reads back, or at will."""
        st = getattr(self, "_last_status", None)
        if st is not None:
            ops.raise_on_status(st)

    def _head(self, feats, plan):
        B, T = plan["B"], plan["T"]
        dev = feats.device
        full = feats if plan["row_of_slot"] is None else ops.gather_rows(feats, plan["row_of_slot"], B * T)
        full = self._adapt(full.view(B, T, -1), plan["valid_dev"]).reshape(B * T, -1).contiguous()
        text = self.get_text_feats().to(device=dev, dtype=torch.float32).contiguous()
        full_logits, logits, probs, top = ops.head(full, plan["valid_u8"], text, B, T, self.logit_scale,
                                                   self.normalize_img_feats, self.agg_func)
        return {
            "full_logits": full_logits,      # [B, T, n_classes]
            "valid_masks": plan["valid_dev"],  # [B, T]
            "logits": logits,                # [B, n_classes]
            "probs": probs,                  # [B, n_classes]
            "top5_logits": top[:, 0],
            "top5_probs": top[:, 1],
        }

    # ---- forward --------------------------------------------------------------------------------------------
    def forward(self, data_dict):
        if "events" in data_dict:
            plan = self.plan_events(data_dict["event_offsets"], data_dict.get("sel_idx", None))
            events = data_dict["events"]
            dev = events.device if events.is_cuda else torch.device("cuda", torch.cuda.current_device())
            if not events.is_cuda:
                events = events.to(dev, non_blocking=True)
            return self.device_forward(events.contiguous(), self.plan_to_device(plan, dev))
        imgs = data_dict["img"]
        valid = data_dict["valid_mask"]
        valid_host = valid.cpu() if isinstance(valid, torch.Tensor) else torch.as_tensor(valid)
        B, T = valid_host.shape
        flat = imgs.reshape(B * T, *imgs.shape[2:])
        vflat = valid_host.reshape(-1)
        row_of_slot = None
        if not bool(vflat.all()):
            idx = vflat.nonzero().squeeze(1)
            flat = flat.index_select(0, idx.to(flat.device))          # the imgs[valid_masks] gather of clip_cls.py:139
            row_of_slot = torch.full((B * T,), -1, dtype=torch.int32)
            row_of_slot[vflat] = torch.arange(idx.numel(), dtype=torch.int32)
        feats = self.get_img_feats(flat)
        plan = dict(frames=torch.empty(0, 16, dtype=torch.uint8), valid=valid_host, valid_u8=valid_host.to(torch.uint8),
                    row_of_slot=row_of_slot, n_valid=int(vflat.sum()), B=B, T=T)
        return self._head(feats, self.plan_to_device(plan, feats.device))

    # ---- losses / metrics (host-side glue; clip_cls.py:164-192) ---------------------------------------------
    def calc_train_loss(self, data_dict, out_dict):
        labels = data_dict["label"].to(out_dict["logits"].device)
        loss = {}
        if self.use_logits_loss:
            loss["ce_loss"] = F.cross_entropy(out_dict["logits"], labels)
        if self.use_probs_loss:
            loss["ce_loss"] = F.nll_loss((out_dict["probs"] + 1e-6).log(), labels)
        return loss

    @torch.no_grad()
    def calc_eval_loss(self, data_dict, out_dict):
        loss = self.calc_train_loss(data_dict, out_dict)
        labels = data_dict["label"].to(out_dict["logits"].device)
        loss["probs_acc"] = (out_dict["probs"].argmax(dim=-1) == labels).float().mean()
        loss["logits_acc"] = (out_dict["logits"].argmax(dim=-1) == labels).float().mean()
        return loss

    @property
    def dtype(self):
        return self.model.logit_scale.dtype

    @property
    def device(self):
        return self.model.logit_scale.device

    def train(self, mode=True):
        nn.Module.train(self, mode)
        self.model.eval()          # keep CLIP in eval mode (clip_cls.py:202-206)
        return self

    def state_dict(self, *args, **kwargs):
        w = super().state_dict(*args, **kwargs)
        return {k: v for k, v in w.items() if self._keep_key(k)}

    def _keep_key(self, k):
        return not k.startswith("model.")

    def load_state_dict(self, state_dict, strict=True):
        clip_w = {f"model.{k}": v for k, v in self.model.state_dict().items() if not self._keep_key(f"model.{k}")}
        return super().load_state_dict({**clip_w, **state_dict}, strict=strict)


class ZSCLIPClassifier(_CLIPClassifierBase):
    """CLIP zero-shot classification (clip_cls.py:14-219): image features are used un-normalised (:148)."""

    def __init__(self, clip_dict=dict(clip_model=None, prompt="a point cloud image of a {}", class_names=None,
                                      agg_func="sum"),
                 loss_dict=dict(use_logits_loss=True, use_probs_loss=False)):
        super().__init__(clip_dict, loss_dict)


class _AdaptedClassifier(_CLIPClassifierBase):
    normalize_img_feats = True

    def __init__(self, adapter_dict, clip_dict, loss_dict):
        super().__init__(clip_dict, loss_dict)
        self.adapter_dict = copy.deepcopy(adapter_dict)
        self._build_adapter()

    allowed_adapters = ("identity", "trans")

    def _build_adapter(self):
        kind = self.adapter_dict.pop("adapter_type").lower()
        self.prompt_tuning = kind.startswith("text-")
        if self.prompt_tuning:      # tune the text features as the FC weight (clip_cls.py:253-259)
            print("Tune text features as well!")
            with torch.no_grad():
                tf = self._clip_text_feats().float()
            self.text_feats = nn.Parameter(tf.clone(), requires_grad=True)
            kind = kind[5:]
        self.adapter_type = kind
        if kind not in self.allowed_adapters:
            raise NotImplementedError(f"adapter {kind} not supported!")
        self.adapter = (IdentityAdapter if kind == "identity" else TransformerAdapter)(**self.adapter_dict)

    def get_text_feats(self, class_names=None):
        if self.prompt_tuning:
            assert self.text_feats.requires_grad or not self.training, "prompt should be trainable!"
            if torch.is_grad_enabled() and self.training:
                raise NotImplementedError("prompt-tuning backward is not built on the B200 path yet")
            return ops.l2norm_rows(self.text_feats.detach().float().contiguous())   # re-normalised every forward (:295)
        return self._clip_text_feats(class_names).to(self.dtype)

    @property
    def dtype(self):
        return self.adapter.dtype


class FSCLIPClassifier(_AdaptedClassifier):
    """Few-shot classifier with a feature adapter (clip_cls.py:222-354)."""

    def __init__(self, adapter_dict=dict(adapter_type="trans", residual=True),
                 clip_dict=dict(clip_model=None, prompt="a point cloud image of a {}", class_names=None, agg_func="sum"),
                 loss_dict=dict(use_logits_loss=False, use_probs_loss=True)):
        super().__init__(adapter_dict, clip_dict, loss_dict)

    def _adapt(self, full_feats, valid_masks):
        return self.adapter(full_feats, valid_masks)

    # ---- training (clip_cls.py:308-350 under autograd in the reference): CLIP stays frozen, gradients reach the adapter
    #      (models/adapter._AdapterFn) and, with a 'text-*' adapter type, the prompt-tuned text features (train._HeadFn) ----
    def _training_active(self):
        return self.training and torch.is_grad_enabled()

    def _head(self, feats, plan):
        if not self._training_active():
            return super()._head(feats, plan)
        from .. import train
        B, T = plan["B"], plan["T"]
        feats = feats.detach().float().contiguous()
        full = feats if plan["row_of_slot"] is None else ops.gather_rows(feats, plan["row_of_slot"], B * T)
        adapted = self._adapt(full.view(B, T, -1), plan["valid_dev"]).reshape(B * T, -1)
        slot_plan = dict(plan)
        slot_plan["row_of_slot"] = None                  # the head sees one feature row per (sample, view) slot
        text_raw = self.text_feats if self.prompt_tuning else self._clip_text_feats().float()
        full_l, logits, probs, t5l, t5p = train._HeadFn.apply(adapted, text_raw, slot_plan, self.logit_scale, self.agg_func)
        return {"full_logits": full_l, "valid_masks": plan["valid_dev"], "logits": logits, "probs": probs,
                "top5_logits": t5l, "top5_probs": t5p, "_plan": slot_plan}

    def calc_train_loss(self, data_dict, out_dict):
        if "_plan" in out_dict and torch.is_grad_enabled():
            from .. import train
            return train.train_loss(self, data_dict, out_dict)
        return super().calc_train_loss(data_dict, out_dict)


class FTCLIPClassifier(_AdaptedClassifier):
    """Fine-tuned CLIP (clip_cls_ft.py:15-333): trainable subsets of model.visual or LoRA; the adapter call is
    skipped in forward (:228) and only 'identity' adapters are accepted (:119)."""

    allowed_adapters = ("identity",)

    def __init__(self, adapter_dict=dict(adapter_type="text-identity", residual=True),
                 clip_dict=dict(clip_model=None, prompt="a point cloud image of a {}", class_names=None, agg_func="sum"),
                 loss_dict=dict(use_logits_loss=True, use_probs_loss=False)):
        super().__init__(adapter_dict, clip_dict, loss_dict)

    def _freeze_clip(self, model):
        """clip_cls_ft.py:45-80: freeze everything, then inject LoRA or unfreeze the selected subset."""
        super()._freeze_clip(model)
        cd = self.clip_dict
        lora = cd.get("lora", -1)
        if isinstance(lora, str) or lora > 0:
            model.visual = inject_trainable_lora(model.visual, r=lora)
        v = model.visual
        conv1, bias, ln = cd["only_conv1"], cd["only_bias"], cd["only_ln"]
        cls_fc, cls_token = cd.get("only_cls_fc", False), cd.get("only_cls_token", False)
        if conv1:
            for p in v.conv1.parameters():
                p.requires_grad = True
        if bias:
            for name, p in v.named_parameters():
                if "bias" in name and p is not None:
                    p.requires_grad = True
        if ln:
            for m in v.modules():
                if isinstance(m, nn.LayerNorm):
                    for p in m.parameters():
                        p.requires_grad = True
        if cls_fc:
            v.proj.requires_grad = True
        if cls_token:
            v.class_embedding.requires_grad = True
        if (isinstance(lora, int) and lora <= 0) and not (conv1 or bias or ln or cls_fc or cls_token):
            for p in v.parameters():
                p.requires_grad = True

    def get_img_feats(self, imgs):
        return self.model.encode_image(imgs)     # gradients flow here in training (clip_cls_ft.py:180; eventclip_b200/train.py)

    # ---- training mode: the same forward with activations kept, attached to autograd through train.py's Functions ----
    def _training_active(self):
        return self.training and torch.is_grad_enabled()

    def device_forward(self, events, plan, status=None):
        if not self._training_active():
            return super().device_forward(events, plan, status)
        fe, visual = self.event_frontend, self.model.visual
        patches, st, _ = ops.event2img(events, plan["frames"], fe.resolution, plan["n_valid"], fe.count_non_zero,
                                       fe.background_mask, out="patch", patch=visual.patch_size, ldk=visual.k_patch,
                                       status=status)
        self._last_status, self._last_patches = st, patches
        out = self._head(visual.forward_patches(patches, plan["n_valid"]), plan)
        out["status"] = st
        return out

    def _head(self, feats, plan):
        if not self._training_active():
            return super()._head(feats.detach(), plan)
        from .. import train
        return train.train_forward(self, feats, plan)

    def calc_train_loss(self, data_dict, out_dict):
        if "_plan" in out_dict and torch.is_grad_enabled():
            from .. import train
            return train.train_loss(self, data_dict, out_dict)
        return super().calc_train_loss(data_dict, out_dict)

    def train(self, mode=True):
        nn.Module.train(self, mode)
        self.model.eval()
        self.model.visual.train(mode)            # clip_cls_ft.py:305-311
        return self

    def _keep_key(self, k):
        return (not k.startswith("model.")) or k.startswith("model.visual.")
