"""Per-call device timing of the C-ABI entry points (CUDA events on the launching stream).
Used by bench.py for the roofline line; never active in normal runs."""
import torch

from . import _lib


def enable():
    _lib.load().timing = []


def disable():
    """Stops timing; returns {call name: (total ms, number of calls)}."""
    lib = _lib.load()
    rec, lib.timing = lib.timing or [], None
    torch.cuda.synchronize()
    out = {}
    for name, e0, e1 in rec:
        ms, cnt = out.get(name, (0.0, 0))
        out[name] = (ms + e0.elapsed_time(e1), cnt + 1)
    return out
