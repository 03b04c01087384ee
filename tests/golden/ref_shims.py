"""Shims that let the UNMODIFIED reference modules under /root/reference import and run on
CPU in the build container.  Used only by the golden-vector
generator (make_golden.py) and by reference-vs-oracle checks that run in this container;
nothing on the GPU box imports this file's targets (/root/reference does not exist there).
"""
import importlib
import os
import sys
import types

import numpy as np
import torch

REF_ROOT = os.environ.get("CRVQA_REFERENCE_ROOT", "/root/reference")


def reference_available():
    return os.path.isdir(os.path.join(REF_ROOT, "masking"))


def install():
    """Register stub packages and monkey patches; idempotent."""
    if getattr(install, "_done", False):
        return
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REF_ROOT}")
    # 1. hg_transformers as a namespace stub (its __init__ needs sacremoses)
    pkg = types.ModuleType("hg_transformers")
    pkg.__path__ = [os.path.join(REF_ROOT, "hg_transformers")]
    pkg.__version__ = "2.10.0"
    sys.modules["hg_transformers"] = pkg
    # 2. huggingface_hub symbols that no longer exist
    import huggingface_hub
    for name in ("HfFolder", "Repository"):
        if not hasattr(huggingface_hub, name):
            setattr(huggingface_hub, name, type(name, (), {}))
    # 3. dataset / h5py / torch._six stubs for the trainers
    ds = types.ModuleType("dataset_LXM")
    ds.Dictionary = type("Dictionary", (), {})
    ds.VQAFeatureDataset = type("VQAFeatureDataset", (), {})
    sys.modules.setdefault("dataset_LXM", ds)
    sys.modules.setdefault("h5py", types.ModuleType("h5py"))
    six = types.ModuleType("torch._six")
    six.string_classes = (str, bytes)
    six.int_classes = (int,)
    six.container_abcs = importlib.import_module("collections.abc")
    sys.modules.setdefault("torch._six", six)
    # 4. CPU: Tensor.cuda -> identity when no GPU
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
    # 6. np.float alias used by Trainer._log
    if not hasattr(np, "float"):
        np.float = float
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    install._done = True


def ref_module(name):
    install()
    return importlib.import_module(name)


def patch_get_init_scales(maskers_mod):
    """Reference bug: maskers_Robust.Masker.replace passes no init_scale,
    so MaskedLinearX.get_init_scales raises TypeError.  Its result is unused with
    controlled_init='magnitude', so return a dummy pair in that case."""
    orig = maskers_mod.MaskedLinearX.get_init_scales

    def tolerant(self, scheme_idx, init_sparsity, init_scale):
        if init_scale is None:
            return (0.0, 0.0)
        return orig(self, scheme_idx, init_sparsity, init_scale)

    maskers_mod.MaskedLinearX.get_init_scales = tolerant
