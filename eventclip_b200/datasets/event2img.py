"""Event -> CLIP-input conversion on the B200.

`Event2Image` is the batched device operator (one fused kernel launch per batch).  `Event2ImageDataset` keeps the
reference's Dataset wrapper interface (datasets/event2img.py:14-145): same constructor arguments, same
`{'img','valid_mask', ...}` items, same view-count rule -- but `__getitem__` enqueues the fused kernel instead of
running numpy + PIL, and `collate_events` offers the packed form the fused classifier forward takes.
"""
import copy

import numpy as np
import torch

from .. import ops


def view_slots(max_n, N, hard_limit):
    """datasets/event2img.py:70-72."""
    return max(min(round(max_n / N), hard_limit), 1)


class Event2Image:
    """Packed event streams -> view images.

    quantize_args follows the reference's schema (max_imgs, N, split_method, convert_method, grayscale,
    count_non_zero, background_mask).  resolution = (H, W) and max_n come from the event dataset.
    """

    def __init__(self, quantize_args, resolution, max_n):
        q = copy.deepcopy(quantize_args)
        assert q.get("split_method", "event_count") == "event_count"
        if q.get("convert_method", "event_histogram") != "event_histogram":
            raise NotImplementedError(f"{q['convert_method']} not implemented!")
        if q.get("grayscale", True) is not True:
            raise NotImplementedError("only grayscale=True is built")
        self.N = int(q["N"])
        self.resolution = tuple(resolution)
        self.max_imgs = view_slots(max_n, self.N, q.get("max_imgs", 10))
        self.count_non_zero = bool(q.get("count_non_zero", False))
        self.background_mask = bool(q.get("background_mask", True))

    def draw_selection(self, offsets, generator=None):
        """The reference picks torch.randperm(K)[:T] when a sample has more chunks than view slots
        (event2img.py:83-86).  Returns int32 [B,T] (identity where K <= T)."""
        T = self.max_imgs
        _, _, chunks, _ = ops.plan_frames(offsets, self.N, T)
        sel = np.tile(np.arange(T, dtype=np.int32), (len(chunks), 1))
        for b, K in enumerate(chunks.tolist()):
            if K > T:
                sel[b] = torch.randperm(K, generator=generator)[:T].numpy().astype(np.int32)
        return sel

    def __call__(self, events, offsets, sel=None, out="f32", compact=False, patch=0, ldk=0, debug=False,
                 check=False):
        """events: float32 [sum E, 4] (CUDA, or CPU -- copied with non_blocking from pinned memory);
        offsets: int64 [B+1] on the host.  Returns a dict with img / valid_mask (CUDA) and bookkeeping."""
        frames, valid, chunks, n_valid = ops.plan_frames(offsets, self.N, self.max_imgs, sel=sel, compact=compact)
        dev = events.device if events.is_cuda else torch.device("cuda", torch.cuda.current_device())
        if not events.is_cuda:
            events = events.to(dev, non_blocking=True)
        B = len(offsets) - 1
        n_slots = n_valid if compact else B * self.max_imgs
        img, status, dbg = ops.event2img(events.contiguous(), frames.to(dev, non_blocking=True), self.resolution,
                                         n_slots, self.count_non_zero, self.background_mask, out=out, patch=patch,
                                         ldk=ldk, debug=debug)
        if check:
            ops.raise_on_status(status)
        if not compact and out != "patch":
            img = img.view(B, self.max_imgs, 3, 224, 224)
        return dict(img=img, valid_mask=valid, chunks=chunks, n_valid=n_valid, status=status, debug=dbg,
                    frames=frames)


class Event2ImageDataset(torch.utils.data.Dataset):
    """Wrapper of an event dataset yielding `{'img' [T,3,224,224] float32, 'valid_mask' [T] bool, 'label', 'data_idx'}`
    (datasets/event2img.py:130-145).  `transforms` is accepted for signature compatibility; the CLIP preprocess it
    stands for is fused into the kernel.  Image-space RandAugment and TTA are outside the hot path."""

    def __init__(self, transforms, event_dataset, quantize_args=dict(max_imgs=2, split_method="event_count",
                 convert_method="event_histogram", N=30000, grayscale=True, count_non_zero=False,
                 background_mask=True), augment=False, tta=False, device="cuda"):
        if augment or tta:
            raise NotImplementedError("RandAugment / flip-TTA run on PIL images in the reference and are out of scope")
        self.transforms = transforms
        self.event_dataset = event_dataset
        self.classes = event_dataset.classes
        self.resolution = event_dataset.resolution
        self.max_t = getattr(event_dataset, "max_t", None)
        self.max_n = event_dataset.max_n
        self.quantize_args = copy.deepcopy(quantize_args)
        self.quantize_args["shape"] = self.resolution
        self.e2i = Event2Image(quantize_args, self.resolution, self.max_n)
        self.max_imgs = self.e2i.max_imgs
        self.keep_events = False
        self.device = device

    def __len__(self):
        return len(self.event_dataset)

    def __getitem__(self, idx):
        data_dict = dict(self.event_dataset[idx])
        events = data_dict.pop("events")
        ev = torch.as_tensor(np.ascontiguousarray(events, dtype=np.float32))
        if self.keep_events:
            data_dict["events"] = copy.deepcopy(events)
        offsets = [0, ev.shape[0]]
        sel = self.e2i.draw_selection(offsets)
        r = self.e2i(ev.to(self.device), offsets, sel=sel, out="f32", check=True)
        data_dict["img"] = r["img"][0]
        data_dict["valid_mask"] = r["valid_mask"][0]
        return data_dict

    def collate_events(self, indices):
        """Packed batch for the fused classifier forward: {'events', 'event_offsets', 'label', 'data_idx'}."""
        items = [self.event_dataset[i] for i in indices]
        evs = [np.ascontiguousarray(it["events"], dtype=np.float32) for it in items]
        off = np.zeros(len(evs) + 1, np.int64)
        off[1:] = np.cumsum([len(e) for e in evs])
        packed = torch.from_numpy(np.concatenate(evs, 0))
        if torch.cuda.is_available():
            packed = packed.pin_memory()
        return dict(events=packed, event_offsets=torch.from_numpy(off),
                    label=torch.tensor([int(it["label"]) for it in items]),
                    data_idx=torch.tensor([int(it.get("data_idx", i)) for it, i in zip(items, indices)]))


def build_event2img_dataset(params, event_dataset, augment=False, tta=False):
    """datasets/event2img.py:148-156."""
    return Event2ImageDataset(transforms=getattr(params, "data_transforms", None), event_dataset=event_dataset,
                              quantize_args=params.quantize_args, augment=augment, tta=tta)
