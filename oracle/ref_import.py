"""oracle/ref_import.py -- TEST INFRASTRUCTURE ONLY; usable only where /root/reference exists.

Imports the UNMODIFIED reference modules in the authoring container (they need `nerv` and `clip`, which
are absent offline: stub modules are placed in sys.modules).  Used solely by
tests/golden/make_golden.py to generate fixtures and by tests that skip when the reference is absent.
"""
import importlib
import importlib.util
import os
import sys
import types

REF = "/root/reference"


def available():
    return os.path.isdir(os.path.join(REF, "models"))


def load_vis():
    spec = importlib.util.spec_from_file_location("_ref_vis", os.path.join(REF, "datasets", "vis.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_utils():
    spec = importlib.util.spec_from_file_location("_ref_dutils", os.path.join(REF, "datasets", "utils.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def load_models():
    """Returns the reference `models` package (ZS/FS/FT classifiers, adapter, lora), unmodified."""
    import torch.nn as nn
    if "nerv" not in sys.modules:
        nerv = types.ModuleType("nerv")
        tr = types.ModuleType("nerv.training")

        class BaseModel(nn.Module):
            pass

        tr.BaseModel = BaseModel
        nerv.training = tr
        sys.modules["nerv"] = nerv
        sys.modules["nerv.training"] = tr
    if "clip" not in sys.modules:
        sys.modules["clip"] = types.ModuleType("clip")
    if REF not in sys.path:
        sys.path.insert(0, REF)
    try:
        return importlib.import_module("models")
    finally:
        sys.path.remove(REF)
