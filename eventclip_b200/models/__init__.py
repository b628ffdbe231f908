"""build_model(params) -- same dispatch as the reference's models/__init__.py:5-21."""
from .classifiers import ZSCLIPClassifier, FSCLIPClassifier, FTCLIPClassifier
from .adapter import IdentityAdapter, TransformerAdapter
from .lora import inject_trainable_lora


def build_model(params):
    if params.model == "ZSCLIP":
        return ZSCLIPClassifier(clip_dict=params.clip_dict)
    if params.model == "FSCLIP":
        return FSCLIPClassifier(adapter_dict=params.adapter_dict, clip_dict=params.clip_dict,
                                loss_dict=params.loss_dict)
    if params.model == "FTCLIP":
        return FTCLIPClassifier(adapter_dict=params.adapter_dict, clip_dict=params.clip_dict,
                                loss_dict=params.loss_dict)
    raise NotImplementedError(f"{params.model} is not implemented.")
