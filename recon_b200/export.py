"""N2: the files the downstream half of RECON reads from the hot path's outputs.

  * `save_embed(embeddings, save_path)`           GAT/main.py:406-413 — `{idx: [floats]}` JSON, `indent=4`; the text is
    byte-identical to the reference's `json.dump(..., cls=CustomEncoder)` but produced by libspkbgat's multi-threaded
    formatter (`spk_export_json`) instead of a Python encoder walking 4x10^8 floats.
  * `save_embed_binary` / `load_embed`            side-car `<path>.bin` (64-byte header + raw fp32 rows) and a reader whose
    result indexes like the JSON dict the consumer uses (`gat_embeddings["17"]`, train.py:100-130,
    utils/context_utils.py:444-504) without parsing gigabytes of text.
  * `save_model(model, name, epoch, folder_name)` GAT/utils.py:26-31 — `trained_{epoch}.pth` state_dict.
  * `save_entity_relation_final_embeddings(model, output_folder)`  GAT/main.py:909-919.
"""
import ctypes as C
import json
import os

import numpy as np
import torch

from . import _lib

BIN_MAGIC = b"SPKEMB01"


def _host_rows(embeddings):
    t = embeddings.detach() if isinstance(embeddings, torch.Tensor) else torch.as_tensor(np.asarray(embeddings))
    if t.dim() != 2:
        raise ValueError("save_embed expects a [rows, width] table")
    t = t.to(device="cpu", dtype=torch.float32)
    if t.dim() == 2 and t.shape[1] and t.stride(1) != 1:
        t = t.contiguous()
    return t


def save_embed(embeddings, save_path, n_threads=0, binary_sidecar=False):
    """Drop-in for GAT/main.py:406-413. `binary_sidecar=True` also writes `save_path + '.bin'`."""
    lib = _lib.load()
    t = _host_rows(embeddings)
    rows, width = int(t.shape[0]), int(t.shape[1])
    ld = int(t.stride(0)) if rows > 1 and width else max(width, 1)
    ld = max(ld, width)
    _lib.check(lib.spk_export_json(t.data_ptr() if t.numel() else None, rows, width, ld, os.fsencode(save_path),
                                   int(n_threads)), "export_json")
    if binary_sidecar:
        save_embed_binary(t, save_path + ".bin")


def save_embed_binary(embeddings, save_path):
    lib = _lib.load()
    t = _host_rows(embeddings)
    rows, width = int(t.shape[0]), int(t.shape[1])
    ld = max(int(t.stride(0)) if rows > 1 and width else width, width)
    _lib.check(lib.spk_export_bin(t.data_ptr() if t.numel() else None, rows, width, ld, os.fsencode(save_path)),
               "export_bin")


def load_embed_array(path, n_threads=0):
    """The [rows, width] fp32 table stored in a `save_embed` JSON file (or any `{"<i>": [numbers]}` JSON), parsed by
    libspkbgat with all host threads instead of `json.load` building one Python float per number."""
    lib = _lib.load()
    rows, width = C.c_int64(0), C.c_int64(0)
    _lib.check(lib.spk_import_json_shape(os.fsencode(path), C.byref(rows), C.byref(width)), "import_json_shape")
    out = np.empty((rows.value, width.value), dtype=np.float32)
    _lib.check(lib.spk_import_json(os.fsencode(path), out.ctypes.data if out.size else None, rows.value, width.value,
                                   max(width.value, 1), int(n_threads)), "import_json")
    return out


class EmbeddingTable:
    """Read-only table that indexes like the reference's loaded JSON: keys are the row ids as `str` (ints accepted
    too), values the rows as lists of Python floats; `.array` is the [rows, width] fp32 array (a memmap of the side-car
    when built from a path, or any array handed in, e.g. `EmbeddingTable(load_embed_array(json_path))`)."""

    def __init__(self, path):
        if isinstance(path, np.ndarray):
            self.array = path
            return
        with open(path, "rb") as f:
            hdr = f.read(64)
        if len(hdr) != 64 or hdr[:8] != BIN_MAGIC:
            raise ValueError(f"{path} is not a recon_b200 embedding side-car")
        rows, width, dtype = (int(x) for x in np.frombuffer(hdr[8:32], dtype="<i8"))
        if dtype != 0:
            raise ValueError(f"{path}: unknown dtype code {dtype}")
        self.array = (np.memmap(path, dtype="<f4", mode="r", offset=64, shape=(rows, width)) if rows * width
                      else np.zeros((rows, width), dtype=np.float32))

    def __len__(self):
        return self.array.shape[0]

    def __contains__(self, key):
        try:
            return 0 <= int(key) < len(self)
        except (TypeError, ValueError):
            return False

    def __getitem__(self, key):
        i = int(key)
        if not 0 <= i < len(self):
            raise KeyError(key)
        return [float(x) for x in self.array[i]]       # same Python floats json.load returns

    def keys(self):
        return (str(i) for i in range(len(self)))

    def items(self):
        return ((str(i), self[i]) for i in range(len(self)))


def load_embed(path, as_table=False):
    """`path` may be the JSON written by `save_embed` (returns the dict `json.load` gives the reference consumer, or with
    `as_table=True` an `EmbeddingTable` filled by the native parser) or a side-car `.bin` (returns an `EmbeddingTable`)."""
    with open(path, "rb") as f:
        head = f.read(8)
    if head == BIN_MAGIC:
        return EmbeddingTable(path)
    if as_table:
        return EmbeddingTable(load_embed_array(path))
    with open(path, "r") as f:
        return json.load(f)


def save_model(model, name, epoch, folder_name):
    """GAT/utils.py:26-31."""
    print("Saving Model")
    os.makedirs(folder_name, exist_ok=True)
    torch.save(model.state_dict(), (folder_name + "trained_{}.pth").format(epoch))
    print("Done saving Model")


def save_entity_relation_final_embeddings(model_gat, output_folder, binary_sidecar=False):
    """GAT/main.py:909-919 (the model is passed in instead of being rebuilt from the globals + trained_0.pth)."""
    save_embed(model_gat.final_entity_embeddings, os.path.join(output_folder, "final_entity_embeddings.json"),
               binary_sidecar=binary_sidecar)
    save_embed(model_gat.final_relation_embeddings, os.path.join(output_folder, "final_relation_embeddings.json"),
               binary_sidecar=binary_sidecar)
    if getattr(model_gat, "W_ent2rel", None) is not None:                  # GAT_sep_space/main.py:982
        save_ent2rel(model_gat, output_folder)


def save_ent2rel(model_gat, output_folder):
    """GAT_sep_space/main.py:982: np.save(join(folder, 'W_ent2rel.json'), W_ent2rel) -- numpy appends the suffix, so the file
    is `W_ent2rel.json.npy` ([R, H*D, H*D] fp32, the relation-space projection the sep-space consumer loads)."""
    import numpy as np
    path = os.path.join(output_folder, "W_ent2rel.json")
    np.save(path, np.array(model_gat.W_ent2rel.detach().cpu()))
    return path + ".npy"
