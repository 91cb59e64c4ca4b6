"""Shared helpers for the parity tests."""
import os
import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

MODEL_CASES = ["model_small_uniform", "model_small_zipf_nhop", "model_small_dropmask",
               "model_refdims", "model_batch_test", "model_3heads"]


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: z[k] for k in z.files}


def rel_l2(a, b):
    def _t(x):
        if isinstance(x, torch.Tensor):
            return x.detach().cpu().double().flatten()
        return torch.as_tensor(np.asarray(x)).double().flatten()
    a, b = _t(a), _t(b)
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)


def golden_params(g, dtype=torch.float32):
    return {k[len("param."):]: torch.as_tensor(v).to(dtype) for k, v in g.items() if k.startswith("param.")}


def golden_masks(g, dtype=torch.float32):
    m = {k[len("mask."):]: torch.as_tensor(v).to(dtype) for k, v in g.items() if k.startswith("mask.")}
    return m or None
