"""N4: ConvKB scoring, SpKBGATConvOnly (GAT and GAT_sep_space variants), W_ent2rel projection and the
relation-ranking evaluation. Fixtures tests/golden/convkb_*.npz were written by the reference's own modules
(tests/golden/make_golden.py convkb). CPU tests pin the oracle; -m gpu tests pin the CUDA path through the C ABI."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_l2

CASES = ["convkb_gat", "convkb_small", "convkb_sep"]


def _params(g, dtype):
    return {k[len("param."):]: torch.as_tensor(v).to(dtype) for k, v in g.items() if k.startswith("param.")}


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("tag,dtype,tol", [("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-12)])
def test_oracle_matches_reference(name, tag, dtype, tol):
    from oracle import convkb as O
    g = load_golden(name)
    p = _params(g, dtype)
    names = [k for k in p if k.startswith(("convKB.fc1", "convKB.fc2"))]
    q = {k: (v.clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    tri = torch.as_tensor(g["triples"])
    sep = "W_ent2rel" in g
    w = torch.as_tensor(g["W_ent2rel"]).to(dtype).requires_grad_(True) if sep else None
    preds = O.conv_only_forward_sep(q, tri, w) if sep else O.conv_only_forward(q, tri)
    loss = torch.nn.SoftMarginLoss()(preds.view(-1), torch.as_tensor(g["target"]).to(dtype))
    loss.backward()
    assert rel_l2(preds, g[tag + ".preds"]) < tol
    assert abs(float(loss) - float(g[tag + ".loss"])) < tol * max(1.0, abs(float(loss)))
    for k in names:
        assert rel_l2(q[k].grad, g[tag + ".grad." + k]) < 10 * tol, k
    if sep:
        assert rel_l2(w.grad, g[tag + ".grad.W_ent2rel"]) < 10 * tol
    test = torch.as_tensor(g["test_triples"])
    r = p["final_relation_embeddings"].shape[0]
    with torch.no_grad():
        sc = O.relation_scores(p, test, r, w.detach() if sep else None)
    assert rel_l2(sc, g[tag + ".scores"]) < tol


def test_relation_ranking_hand_checked():
    """create_batch.py:1413-1495 on a 3-row toy: pair (0,1) has two actual relations, predictions come from its first row."""
    from oracle import convkb as O
    test = torch.tensor([[0, 2, 1], [0, 3, 1], [4, 0, 5]])
    r = 12
    scores = torch.zeros(3, r)
    scores[0] = torch.arange(r, 0, -1).float()           # row 0 ranks relation 0 first, 1 second, ...
    scores[1] = torch.arange(r).float()
    scores[2] = torch.arange(r).float(); scores[2, 0] = 100.0     # the last row ranks relation 0 first, then 11, 10, ...
    idx, m = O.relation_ranking(scores, test)
    assert idx[0].tolist() == list(range(r))
    # pair (0,1): top-10 of row 0 = {0..9} contains 2 and 3 -> 2 hits; pair (4,5): top-10 of row 2 = {0, 11, 10, ..., 3} contains 0
    assert m["hits_at_10"] == 3
    # ranks are looked up in the LAST row's ordering (reference quirk): rel 2 -> position 11 (0-based 10+1)... checked numerically
    last = idx[2].tolist()
    exp = [last.index(2) + 1, last.index(3) + 1, last.index(0) + 1]
    assert abs(m["average_rank"] - sum(exp) / 3) < 1e-12


# ---- CUDA path ---------------------------------------------------------------------------------------
def _dev():
    return torch.device("cuda:0")


def _build(g, sep):
    from recon_b200 import SpKBGATConvOnly
    p = _params(g, torch.float32)
    n, d = p["final_entity_embeddings"].shape
    r = p["final_relation_embeddings"].shape[0]
    m = SpKBGATConvOnly(torch.zeros(n, 10), torch.zeros(r, 10), [d // 2, d], [d // 2, d], 0.0, 0.0, 0.2, 0.2, [2, 2], 50)
    m.load_state_dict(p)
    m = m.to(_dev())
    gat = None
    if sep:
        import types
        gat = types.SimpleNamespace(W_ent2rel=torch.nn.Parameter(torch.as_tensor(g["W_ent2rel"]).to(_dev())),
                                    nonlinearity_ent2rel=torch.tanh)
    return m, gat, r


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("frozen", [True, False])
def test_conv_only_matches_reference(name, frozen):
    """forward + SoftMarginLoss backward (GAT/main.py:753,818-835) vs the reference's fp64 run; with frozen embeddings
    (main.py:741-742) and with trainable ones (their gradients against the fp64 oracle)."""
    from oracle import convkb as O
    g = load_golden(name)
    sep = "W_ent2rel" in g
    m, gat, r = _build(g, sep)
    m.final_entity_embeddings.requires_grad = not frozen
    m.final_relation_embeddings.requires_grad = not frozen
    tri = torch.as_tensor(g["triples"])
    preds = m(None, None, tri, gat) if sep else m(None, None, tri)
    loss = torch.nn.SoftMarginLoss()(preds.view(-1), torch.as_tensor(g["target"]).to(_dev()))
    loss.backward()
    assert rel_l2(preds, g["f64.preds"]) < 2e-5 and rel_l2(preds, g["f32.preds"]) < 1e-4
    assert abs(float(loss) - float(g["f64.loss"])) < 2e-5
    for k in ("convKB.fc1.weight", "convKB.fc1.bias", "convKB.fc2.weight", "convKB.fc2.bias"):
        got = dict(m.named_parameters())[k].grad
        assert rel_l2(got, g["f64.grad." + k]) < 2e-5, k
    if sep:
        assert rel_l2(gat.W_ent2rel.grad, g["f64.grad.W_ent2rel"]) < 2e-5
    if not frozen:
        p = _params(g, torch.float64)
        q = {k: (v.clone().requires_grad_(True) if k.startswith("final_") else v) for k, v in p.items()}
        w = torch.as_tensor(g["W_ent2rel"]).double() if sep else None
        ref = O.conv_only_forward_sep(q, tri, w) if sep else O.conv_only_forward(q, tri)
        torch.nn.SoftMarginLoss()(ref.view(-1), torch.as_tensor(g["target"]).double()).backward()
        assert rel_l2(m.final_entity_embeddings.grad, q["final_entity_embeddings"].grad) < 2e-5
        assert rel_l2(m.final_relation_embeddings.grad, q["final_relation_embeddings"].grad) < 2e-5
    # deterministic
    m.zero_grad()
    preds2 = m(None, None, tri, gat) if sep else m(None, None, tri)
    assert torch.equal(preds, preds2)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_relation_scores_and_ranking(name):
    """All-relations scores (re-associated fc1 + streaming pass) vs the reference's batch_test on the tiled triples
    (create_batch.py:1367-1393), and the ranking metrics vs the oracle's restatement of lines 1413-1495."""
    from recon_b200 import rank_relations
    from oracle import convkb as O
    g = load_golden(name)
    sep = "W_ent2rel" in g
    m, gat, r = _build(g, sep)
    test = torch.as_tensor(g["test_triples"])
    scores, idx, metrics = rank_relations(m, test, r, gat)
    assert rel_l2(scores, g["f64.scores"]) < 2e-5 and rel_l2(scores, g["f32.scores"]) < 1e-4
    ref_idx, ref_metrics = O.relation_ranking(torch.as_tensor(g["f64.scores"]), test)
    # ties aside (none in these fixtures) the orderings agree, hence the metrics
    assert torch.equal(idx.cpu(), ref_idx)
    for k, v in ref_metrics.items():
        assert abs(metrics[k] - v) < 1e-12, k


@pytest.mark.gpu
def test_conv_only_errors_and_shapes():
    from recon_b200 import SpKBGATConvOnly
    m = SpKBGATConvOnly(torch.zeros(20, 10), torch.zeros(4, 10), [6, 12], [6, 12], 0.0, 0.0, 0.2, 0.2, [2, 2], 50).to(_dev())
    assert m(None, None, torch.zeros(0, 3, dtype=torch.long)).shape == (0, 1)
    with pytest.raises(IndexError):
        m.batch_test(torch.tensor([[0, 1, 20]]))                    # tail id out of range
    with pytest.raises(IndexError):
        m.batch_test(torch.tensor([[0, 4, 1]]))                     # relation id out of range
    with pytest.raises(RuntimeError):
        m.cpu().batch_test(torch.tensor([[0, 1, 2]]))               # no CPU fallback


def test_sep_space_state_dict_both_ways(tmp_path):
    """W_ent2rel [R, H*D, H*D] sits in the state dict of the sep-space variant; checkpoints of
    either variant load into either model; W_ent2rel.json.npy is what GAT_sep_space/main.py:982 writes."""
    from recon_b200 import SpKBGATModified
    from recon_b200.export import save_ent2rel
    args = (torch.randn(9, 6), torch.randn(4, 6), [5, 10], [5, 10], 0.0, 0.2, [2, 2], None)
    plain = SpKBGATModified(*args)
    sep = SpKBGATModified(*args, sep_space=True)
    assert "W_ent2rel" not in plain.state_dict() and tuple(sep.state_dict()["W_ent2rel"].shape) == (4, 10, 10)
    sd_sep = {k: v.clone() for k, v in sep.state_dict().items()}
    plain2 = SpKBGATModified(*args)
    plain2.load_state_dict(sd_sep)                                   # a sep-space checkpoint into a plain model
    assert torch.equal(plain2.W_ent2rel, sep.W_ent2rel)
    w0 = sep.W_ent2rel.detach().clone()
    sep.load_state_dict(plain.state_dict())                         # a plain checkpoint into a sep-space model
    assert torch.equal(sep.W_ent2rel, w0)
    path = save_ent2rel(sep, str(tmp_path))
    assert path.endswith("W_ent2rel.json.npy") and np.array_equal(np.load(path), w0.numpy())
