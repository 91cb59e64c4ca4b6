"""TEST INFRASTRUCTURE ONLY (checker for tests/) -- never imported by recon_b200.

CPU restatement of the reference's embedding export:
  * save_embed      /root/reference/GAT/main.py:406-413
  * CustomEncoder   /root/reference/GAT/main.py:127-144
Pinned against the reference's own function text (executed unmodified by tests/golden/make_golden.py ->
tests/golden/export_*.json) in tests/test_export.py.
"""
import datetime
import json

import numpy as np


class CustomEncoder(json.JSONEncoder):                       # main.py:127-144
    def default(self, obj):
        if isinstance(obj, (np.int32, np.int64)):
            return int(obj)
        if isinstance(obj, (np.float32, np.float64)):
            return float(obj)
        if isinstance(obj, np.ndarray):
            return obj.tolist()
        if isinstance(obj, datetime.datetime):
            return str(obj)
        if isinstance(obj, np.bool_):
            return bool(obj)
        return json.JSONEncoder.default(self, obj)


def save_embed(embeddings, save_path):                       # main.py:406-413
    emb_data = {}
    for idx in range(embeddings.shape[0]):
        emb_data[idx] = embeddings[idx].cpu().detach().numpy()      # np.array(tensor) in the reference: same values
    with open(save_path, "w") as f:
        json.dump(emb_data, f, indent=4, cls=CustomEncoder)
