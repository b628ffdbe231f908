"""eventclip_b200 -- B200-native (sm_100a) implementation of EventCLIP's inference hot path:
event stream -> frames -> CLIP ViT image encoder -> adapter / text-cosine logit head.

Layout:
  csrc/        hand-written CUDA kernels + the C ABI (include/eventclip_b200.h) -> libeventclip_b200.so
  _lib.py      ctypes binding (fails loudly when the library or a B200 is missing; there is no fallback)
  ops.py       tensor-level wrappers (PyTorch = device memory + streams only)
  clip.py      CLIP object with openai-CLIP's parameter names; encoder forward through the library
  datasets/    mirror of the reference's datasets/vis.py + datasets/event2img.py interfaces
  models/      mirror of the reference's models/ (build_model, ZS/FS/FT classifiers, adapter, LoRA)
  synth.py     seeded synthetic event streams of each dataset's sensor shape
"""
__version__ = "0.1.0"
