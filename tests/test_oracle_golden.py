"""Pins oracle/ (the CPU restatement) against fixtures produced by the real reference
(tests/golden/make_golden.py). CPU only."""
import numpy as np
import pytest
import torch

from oracle import ref_torch as O
from oracle import edges as OE
from helpers import load_golden, rel_l2, golden_params, golden_masks, MODEL_CASES


@pytest.mark.parametrize("name", MODEL_CASES)
@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_model_oracle_matches_reference(name, prec):
    g = load_golden(name)
    dt = torch.float32 if prec == "f32" else torch.float64
    p = golden_params(g, dt)
    masks = golden_masks(g, dt)
    adj = (torch.as_tensor(g["edge"]), torch.as_tensor(g["edge_type"]))
    nhop = torch.as_tensor(g["nhop"])
    be = torch.as_tensor(g["batch_entities"])
    ge, gr = torch.as_tensor(g["g_ent"]).to(dt), torch.as_tensor(g["g_rel"]).to(dt)
    tol = 2e-5 if prec == "f32" else 1e-12
    if name == "model_batch_test":
        names = O.trainable_names(p)
        q = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
        ent_in = torch.as_tensor(g["entity_in"]).to(dt)
        out, out_rel, mask, _ = O.kbgat_forward(q, be, adj, nhop, 0.2, masks, entity_override=ent_in, detach_rel=True)
        ((out * ge).sum() + (out_rel * gr).sum()).backward()
        grads = {k: q[k].grad for k in names if q[k].grad is not None}
    else:
        out, out_rel, mask, ent, grads = O.fwd_bwd(p, be, adj, nhop, 0.2, ge, gr, masks)
        assert rel_l2(ent, g[f"{prec}.after.entity_embeddings"]) < tol
    assert rel_l2(out, g[f"{prec}.out_entity"]) < tol
    assert rel_l2(out_rel, g[f"{prec}.out_relation"]) < tol
    assert np.array_equal(mask.detach().numpy(), g[f"{prec}.mask"])
    checked = 0
    for k, v in g.items():
        if k.startswith(f"{prec}.grad."):
            nm = k[len(f"{prec}.grad."):]
            assert nm in grads and grads[nm] is not None, nm
            assert rel_l2(grads[nm], v) < 10 * tol, (nm, rel_l2(grads[nm], v))
            checked += 1
    assert checked >= 5


def test_index_add_variant_equals_coo():
    g = load_golden("model_small_zipf_nhop")
    p = golden_params(g, torch.float64)
    adj = (torch.as_tensor(g["edge"]), torch.as_tensor(g["edge_type"]))
    a = O.fwd_bwd(p, torch.as_tensor(g["batch_entities"]), adj, torch.as_tensor(g["nhop"]), 0.2,
                  torch.as_tensor(g["g_ent"]).double(), torch.as_tensor(g["g_rel"]).double(), None, O.seg_sum_index_add)
    assert rel_l2(a[0], g["f64.out_entity"]) < 1e-12
    for k, v in a[4].items():
        assert rel_l2(v, g["f64.grad." + k]) < 1e-11


@pytest.mark.parametrize("name", ["layer_concat", "layer_noconcat_nhop"])
def test_layer_oracle_matches_reference(name):
    g = load_golden(name)
    x = torch.as_tensor(g["x"]).requires_grad_(True)
    emb = torch.as_tensor(g["edge_embed"]).requires_grad_(True)
    a = torch.as_tensor(g["a"]).requires_grad_(True)
    a2 = torch.as_tensor(g["a_2"]).requires_grad_(True)
    has2 = g["edge_nhop"].size > 0
    e2 = torch.as_tensor(g["edge_nhop"]) if has2 else None
    emb2 = torch.as_tensor(g["edge_embed_nhop"]).requires_grad_(True) if has2 else None
    out = O.attention_layer(x, torch.as_tensor(g["edge"]), emb, e2, emb2, a, a2, 0.2, bool(g["concat"]))
    (out * torch.as_tensor(g["g"])).sum().backward()
    assert rel_l2(out, g["out"]) < 1e-6
    for nm, t in (("x", x), ("edge_embed", emb), ("a", a), ("a_2", a2)):
        assert rel_l2(t.grad, g["grad." + nm]) < 1e-5, nm
    if has2:
        assert rel_l2(emb2.grad, g["grad.edge_embed_nhop"]) < 1e-5


@pytest.mark.parametrize("name", ["spmm_f1", "spmm_f7"])
def test_spmm_oracle_matches_reference(name):
    g = load_golden(name)
    w = torch.as_tensor(g["w"]).requires_grad_(True)
    for fn in (O.seg_sum_coo, O.seg_sum_index_add):
        w.grad = None
        out = fn(torch.as_tensor(g["edge"]), w, g["out"].shape[0])
        (out * torch.as_tensor(g["g"])).sum().backward()
        assert rel_l2(out, g["out"]) < 1e-6
        assert np.array_equal(w.grad.numpy(), g["grad_w"])


@pytest.mark.parametrize("name", ["edges_a", "edges_b", "edges_partial"])
def test_edge_oracle_matches_corpus(name):
    g = load_golden(name)
    tr = g["triples"]
    rows, cols, data = OE.triples_to_adj(tr.tolist())
    graph = OE.build_graph(rows, cols, data)
    batch = g["batch"].tolist()
    idx, val = OE.batch_adj(graph, batch)
    assert np.array_equal(np.asarray(idx, dtype=np.int64).reshape(2, -1), g["adj_idx"])
    assert np.array_equal(np.asarray(val, dtype=np.int64), g["adj_val"])
    partial = bool(g["partial"])
    nh = OE.batch_nhop(graph, batch, partial)
    assert np.array_equal(np.asarray(nh, dtype=np.int32).reshape(-1, 4), g["nhop"])
    n = int(tr[:, [0, 2]].max()) + 1          # full-graph variants use sources 0..n-1 in id order
    fidx, fval = OE.batch_adj(graph, list(range(n)))
    assert np.array_equal(np.asarray(fidx, dtype=np.int64).reshape(2, -1), g["full_adj_idx"])
    fn = OE.batch_nhop(graph, list(range(n)), partial)
    assert np.array_equal(np.asarray(fn, dtype=np.int32).reshape(-1, 4), g["full_nhop"])


def test_edge_oracle_toy_kg():
    """The hand-checked toy KG of SURVEY.md 3.4."""
    tr = load_golden("edges_toy")["triples"]
    graph = OE.build_graph(*OE.triples_to_adj(tr.tolist()))
    idx, val = OE.batch_adj(graph, [0, 1, 2, 3])
    assert idx == [[1, 1, 2, 3, 0, 3, 4], [0, 0, 0, 1, 1, 2, 3]]
    assert val == [5, 6, 7, 8, 2, 9, 1]
    assert OE.batch_nhop(graph, [0, 1, 2, 3]) == [[0, 5, 8, 3], [1, 8, 1, 4], [1, 2, 7, 2], [2, 9, 1, 4]]


# ---- preprocess.load_data: triples -> adjacency (rows = tail, cols = head), all flag combinations -------------------
@pytest.mark.parametrize("directed", [True, False])
@pytest.mark.parametrize("unweighted", [False, True])
def test_triples_to_adj_matches_reference_load_data(directed, unweighted):
    import numpy as np
    from oracle import edges as OE
    from recon_b200 import triples_to_adj
    g = load_golden("load_data_small")
    want = g[f"adj.directed{int(directed)}.unweighted{int(unweighted)}"]
    rows, cols, data = OE.triples_to_adj(g["triples"].tolist(), unweighted, directed)
    assert np.array_equal(np.asarray([rows, cols, data]), want)
    edge, val = triples_to_adj(torch.as_tensor(g["triples"]), unweighted, directed)       # product helper (index glue)
    assert edge.dtype == torch.int64 and np.array_equal(edge.numpy(), want[:2]) and np.array_equal(val.numpy(), want[2])
