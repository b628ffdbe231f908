"""Host-side mirror of the reference's data interface for the hot path (datasets/vis.py, datasets/event2img.py)."""
from .vis import events2frames, split_event_count
from .event2img import Event2Image, Event2ImageDataset, build_event2img_dataset

__all__ = ["events2frames", "split_event_count", "Event2Image", "Event2ImageDataset", "build_event2img_dataset"]
