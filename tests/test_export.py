"""N2: embedding export. Everything here is host-side (no GPU): the JSON writer must be byte-identical
to the reference's save_embed (GAT/main.py:406-413, fixtures written by the reference's own function text), the
side-car must round-trip, and the loaders must index like the dict the consumer uses."""
import json
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = ["export_small", "export_w200"]


@pytest.mark.parametrize("name", CASES)
def test_oracle_save_embed_matches_reference_bytes(name, tmp_path):
    from oracle import export as OE
    t = torch.as_tensor(np.load(os.path.join(GOLDEN, name + ".npy")))
    out = tmp_path / "o.json"
    OE.save_embed(t, str(out))
    assert out.read_bytes() == open(os.path.join(GOLDEN, name + ".json"), "rb").read()


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("threads", [1, 3, 0])
def test_save_embed_matches_reference_bytes(name, threads, tmp_path):
    from recon_b200.export import save_embed
    t = torch.as_tensor(np.load(os.path.join(GOLDEN, name + ".npy")))
    out = tmp_path / "p.json"
    save_embed(t, str(out), n_threads=threads)
    assert out.read_bytes() == open(os.path.join(GOLDEN, name + ".json"), "rb").read()


def test_save_embed_random_bit_patterns_vs_oracle(tmp_path):
    """200k random float32 bit patterns (all exponents, denormals, NaNs, infinities) + a strided input view."""
    from oracle import export as OE
    from recon_b200.export import save_embed
    rng = np.random.default_rng(0)
    bits = rng.integers(0, 2 ** 32, size=(1000, 200), dtype=np.uint64).astype(np.uint32)
    t = torch.as_tensor(bits.view(np.float32))
    a, b = tmp_path / "a.json", tmp_path / "b.json"
    save_embed(t, str(a))
    OE.save_embed(t, str(b))
    assert a.read_bytes() == b.read_bytes()
    wide = torch.randn(300, 40)
    view = wide[:, 3:28]                                      # row stride 40, width 25
    save_embed(view, str(a), n_threads=2)
    OE.save_embed(view, str(b))
    assert a.read_bytes() == b.read_bytes()


def test_save_embed_many_rows_multi_wave(tmp_path):
    """More rows than one wave of tasks: block boundaries and the trailing commas stay right."""
    from oracle import export as OE
    from recon_b200.export import save_embed
    t = torch.randn(70001, 3, generator=torch.Generator().manual_seed(1))
    a, b = tmp_path / "a.json", tmp_path / "b.json"
    save_embed(t, str(a), n_threads=2)
    OE.save_embed(t, str(b))
    assert a.read_bytes() == b.read_bytes()


@pytest.mark.parametrize("shape", [(0, 5), (1, 1), (4, 0)])
def test_save_embed_degenerate_shapes(shape, tmp_path):
    from oracle import export as OE
    from recon_b200.export import save_embed, save_embed_binary, load_embed
    t = torch.zeros(shape)
    a, b = tmp_path / "a.json", tmp_path / "b.json"
    save_embed(t, str(a))
    OE.save_embed(t, str(b))
    assert a.read_bytes() == b.read_bytes()
    save_embed_binary(t, str(tmp_path / "a.bin"))
    assert load_embed(str(tmp_path / "a.bin")).array.shape == shape


def test_sidecar_round_trip_and_consumer_indexing(tmp_path):
    from recon_b200.export import save_embed, load_embed, EmbeddingTable
    t = torch.randn(123, 200, generator=torch.Generator().manual_seed(2))
    p = str(tmp_path / "final_entity_embeddings.json")
    save_embed(t, p, binary_sidecar=True)
    as_json = load_embed(p)                                   # what train.py:103-104 does
    table = load_embed(p + ".bin")
    assert isinstance(as_json, dict) and isinstance(table, EmbeddingTable)
    assert len(table) == len(as_json) == 123
    assert np.array_equal(table.array, t.numpy())             # bit-exact fp32
    for key in ("0", "17", "122"):                            # context_utils.py:445,452 style access
        assert table[key] == as_json[key]
        assert len(table[key]) == 200
    assert "122" in table and "123" not in table
    with pytest.raises(KeyError):
        table["123"]
    assert list(table.keys())[:3] == list(as_json.keys())[:3]
    assert np.array_equal(np.asarray(json.load(open(p))["5"], dtype=np.float32), t[5].numpy())   # text round-trips fp32


def test_save_model_and_final_embeddings(tmp_path):
    from recon_b200 import SpKBGATModified
    from recon_b200.export import save_model, save_entity_relation_final_embeddings, load_embed
    m = SpKBGATModified(torch.randn(8, 4), torch.randn(3, 4), [4, 8], [4, 8], 0.0, 0.2, [2, 2], None)
    folder = str(tmp_path) + "/"
    save_model(m, "x", 0, folder)
    sd = torch.load(folder + "trained_0.pth")
    assert set(sd) == set(m.state_dict())
    save_entity_relation_final_embeddings(m, folder, binary_sidecar=True)
    ent = load_embed(folder + "final_entity_embeddings.json")
    assert len(ent) == 8 and len(ent["0"]) == 8
    assert np.array_equal(load_embed(folder + "final_relation_embeddings.json.bin").array,
                          m.final_relation_embeddings.detach().numpy())


# ---- native reader of the JSON text ------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_load_embed_array_reads_reference_files(name):
    from recon_b200.export import load_embed_array
    want = np.load(os.path.join(GOLDEN, name + ".npy"))
    got = load_embed_array(os.path.join(GOLDEN, name + ".json"))           # files written by the reference's own save_embed
    assert got.dtype == np.float32 and got.shape == want.shape
    assert np.array_equal(got, want, equal_nan=True)
    assert np.array_equal(np.signbit(got), np.signbit(want))               # -0.0 survives


@pytest.mark.parametrize("threads", [1, 2, 5, 0])
def test_load_embed_array_round_trip_multithreaded(threads, tmp_path):
    """3.5 MB of text: several byte ranges, cuts falling inside keys and numbers; random bit patterns incl. denormals."""
    from recon_b200.export import save_embed, load_embed_array, load_embed
    rng = np.random.default_rng(3)
    bits = rng.integers(0, 2 ** 32, size=(20011, 7), dtype=np.uint64).astype(np.uint32)
    t = bits.view(np.float32).copy()
    p = str(tmp_path / "t.json")
    save_embed(torch.as_tensor(t), p)
    assert os.path.getsize(p) > (1 << 20)
    got = load_embed_array(p, n_threads=threads)
    finite = np.isfinite(t)
    assert np.array_equal(got[finite].view(np.uint32), t[finite].view(np.uint32))      # bit-exact incl. denormals, -0.0
    assert np.array_equal(np.isnan(got), np.isnan(t)) and np.array_equal(got[np.isinf(t)], t[np.isinf(t)])
    table = load_embed(p, as_table=True)
    ref = json.load(open(p))
    for key in ("0", "9", "20010"):
        a, b = table[key], ref[key]
        assert len(a) == len(b) and all((x == y) or (x != x and y != y) for x, y in zip(a, b))


def test_load_embed_array_other_json_layouts_and_errors(tmp_path):
    from recon_b200.export import load_embed_array
    data = {str(i): [float(i) + 0.25, -1e-3 * i, 3e10] for i in range(50)}
    p = tmp_path / "compact.json"
    p.write_text(json.dumps(data))                                           # no indentation at all
    got = load_embed_array(str(p))
    assert np.array_equal(got, np.asarray([data[str(i)] for i in range(50)], dtype=np.float32))
    (tmp_path / "empty.json").write_text("{}")
    assert load_embed_array(str(tmp_path / "empty.json")).shape == (0, 0)
    (tmp_path / "ragged.json").write_text('{"0": [1.0, 2.0], "1": [3.0]}')
    with pytest.raises(RuntimeError):
        load_embed_array(str(tmp_path / "ragged.json"))
    (tmp_path / "bad.json").write_text('{"0": [1.0, oops]}')
    with pytest.raises(RuntimeError):
        load_embed_array(str(tmp_path / "bad.json"))
    with pytest.raises(RuntimeError):
        load_embed_array(str(tmp_path / "missing.json"))


def test_load_embed_array_rejects_duplicate_and_overlong_keys(tmp_path):
    """A repeated key leaves another row unwritten and a 20-digit key overflows int64: both must fail, not corrupt memory."""
    from recon_b200.export import load_embed_array
    (tmp_path / "dup.json").write_text('{"0": [1.0, 2.0], "0": [3.0, 4.0]}')
    with pytest.raises(RuntimeError):
        load_embed_array(str(tmp_path / "dup.json"))
    (tmp_path / "long.json").write_text('{"18446744073709551616": [1.0, 2.0], "1": [3.0, 4.0]}')
    with pytest.raises(RuntimeError):
        load_embed_array(str(tmp_path / "long.json"))
