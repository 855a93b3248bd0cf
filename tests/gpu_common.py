"""Loader for the product binding used by the -m gpu tests and the ABI tests."""
import importlib.util
import os
import sys

from helpers import ROOT

_mod = None


def binding():
    global _mod
    if _mod is None:
        path = os.path.join(ROOT, "lzs-compression_b200", "python", "lzs_b200.py")
        spec = importlib.util.spec_from_file_location("lzs_b200", path)
        _mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(_mod)
        sys.modules.setdefault("lzs_b200", _mod)        # lzs_torch imports it by name
    return _mod


def torch_front_end():
    binding()
    path = os.path.join(ROOT, "lzs-compression_b200", "python", "lzs_torch.py")
    spec = importlib.util.spec_from_file_location("lzs_torch", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
