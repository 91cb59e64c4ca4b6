"""Loader for the UNMODIFIED reference modules -- TEST INFRASTRUCTURE, NOT PRODUCT.

`__graft_entry__.build()` copies /root/reference/GAT/{layers,models}.py (pure Python, nothing to compile) into the
git-ignored oracle/_ref/GAT/ when the reference tree is present (the build container); the directory travels to the
GPU box with the snapshot like a built .so. This module imports those two files as they are and pins the module
globals `CUDA` (GAT/layers.py:9, GAT/models.py:8, read at call time: layers.py:72, models.py:150,167,206,224) to False
so the reference takes its CPU branches on a box that has GPUs. Only bench.py's CPU legs and tests/
may use it; nothing under recon_b200/ does.
"""
import importlib.util
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref", "GAT")
REF_SRC = "/root/reference/GAT"
FILES = ("layers.py", "models.py")


def install():
    """Copy the two reference files into oracle/_ref/GAT (no-op when the reference tree is absent)."""
    import shutil
    if not all(os.path.exists(os.path.join(REF_SRC, f)) for f in FILES):
        return False
    os.makedirs(REF_DIR, exist_ok=True)
    for f in FILES:
        shutil.copyfile(os.path.join(REF_SRC, f), os.path.join(REF_DIR, f))
    return True


def available():
    return all(os.path.exists(os.path.join(REF_DIR, f)) for f in FILES)


def load():
    """Returns (layers, models) modules of the reference with CUDA pinned to False, or None if not installed."""
    if not available():
        return None
    mods = {}
    for name in ("layers", "models"):                      # models.py does `from layers import ...`
        key = "_recon_ref_" + name
        if key in sys.modules:
            mods[name] = sys.modules[key]
            continue
        spec = importlib.util.spec_from_file_location(key, os.path.join(REF_DIR, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        saved = sys.modules.get("layers")
        if name == "models":
            sys.modules["layers"] = mods["layers"]
        try:
            spec.loader.exec_module(mod)
        finally:
            if name == "models":
                if saved is not None:
                    sys.modules["layers"] = saved
                else:
                    sys.modules.pop("layers", None)
        mod.CUDA = False
        sys.modules[key] = mod
        mods[name] = mod
    return mods["layers"], mods["models"]
