"""N1: batch_gat_loss + its backward + the SGD step.

CPU (`-m "not gpu"`): oracle/loss.py against the fixtures produced by the reference's own function
(GAT/main.py:344-376, executed unmodified by tests/golden/make_golden.py).
GPU (`-m gpu`): recon_b200.loss through the C ABI against the same fixtures (fp64 run of the reference,
rel <= 1e-4 = north_star tolerance, expected ~1e-6) and against the oracle on larger seeded inputs with hub
entities / relations; determinism; the reference's error behaviour."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_l2

LOSS_CASES = ["loss_small", "loss_refdims_hub", "loss_ratio3_oddwidth"]
TOL = 1e-4
TIGHT = 2e-5


def _case(name):
    g = load_golden(name)
    return (g, torch.as_tensor(g["train_indices"]), torch.as_tensor(g["entity_embed"]),
            torch.as_tensor(g["relation_embed"]), int(g["ratio"]), float(g["margin"]))


# ---- CPU: the oracle is pinned to the reference ------------------------------------------------------------------
@pytest.mark.parametrize("name", LOSS_CASES)
@pytest.mark.parametrize("tag,dt,tol", [("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-13)])
def test_loss_oracle_matches_reference(name, tag, dt, tol):
    from oracle import loss as OL
    g, tri, ent, rel, ratio, margin = _case(name)
    loss, d_ent, d_rel = OL.loss_fwd_bwd(tri, ent.to(dt), rel.to(dt), ratio, margin)
    assert abs(float(loss) - float(g[f"{tag}.loss"])) <= tol * abs(float(g[f"{tag}.loss"]))
    assert rel_l2(d_ent, g[f"{tag}.grad.entity_embed"]) <= tol
    assert rel_l2(d_rel, g[f"{tag}.grad.relation_embed"]) <= tol
    new_ent, new_rel = OL.sgd_step([ent.to(dt), rel.to(dt)], [d_ent, d_rel], float(g["lr"]))
    assert rel_l2(new_ent, g[f"{tag}.sgd.entity_embed"]) <= tol
    assert rel_l2(new_rel, g[f"{tag}.sgd.relation_embed"]) <= tol


def test_make_train_indices_layout():
    from oracle.loss import make_train_indices
    tri = make_train_indices(50, 4, 20, 2, seed=0)
    assert tri.shape == (100, 3)
    pos = tri[:20]
    assert torch.equal(tri[20:60, 1:], pos[:, 1:].repeat(2, 1))          # head-corrupted: relation, tail kept
    assert torch.equal(tri[60:, :2], pos[:, :2].repeat(2, 1))            # tail-corrupted: head, relation kept


def test_loss_has_no_cpu_fallback():
    from recon_b200.loss import batch_gat_loss, sgd_step
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU behaviour")
    _, tri, ent, rel, ratio, margin = _case("loss_small")
    with pytest.raises(RuntimeError):
        batch_gat_loss(torch.nn.MarginRankingLoss(margin=margin), tri, ent, rel, valid_invalid_ratio_gat=ratio)
    p = torch.nn.Parameter(torch.ones(4)); p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        sgd_step([p], 0.1)


# ---- GPU ---------------------------------------------------------------------------------------------------------------
def dev():
    return torch.device("cuda:0")


def _run_gpu(tri, ent, rel, ratio, margin, reduction="mean"):
    from recon_b200.loss import batch_gat_loss
    e = ent.to(dev()).requires_grad_(True)
    r = rel.to(dev()).requires_grad_(True)
    loss = batch_gat_loss(torch.nn.MarginRankingLoss(margin=margin, reduction=reduction), tri.to(dev()), e, r,
                          valid_invalid_ratio_gat=ratio)
    loss.backward()
    torch.cuda.synchronize()
    return loss.detach(), e, r


def _margin_without_ties(tri, ent, rel, ratio, margin):
    """Nudge the margin until no pair's hinge argument is within 1e-4 of 0: a pair that close can flip between fp32
    and fp64 arithmetic, which changes the gradient by a whole 1/M step (ill-conditioned by nature, not a kernel error)."""
    n_pos = tri.shape[0] // (2 * ratio + 1)
    e, r = ent.double(), rel.double()
    norm = (e[tri[:, 0]] + r[tri[:, 1]] - e[tri[:, 2]]).abs().sum(1)
    d = norm[:n_pos].repeat(2 * ratio) - norm[n_pos:]
    while float((d + margin).abs().min()) < 1e-4:
        margin += 1e-3
    return margin


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOSS_CASES)
def test_loss_golden(name):
    g, tri, ent, rel, ratio, margin = _case(name)
    loss, e, r = _run_gpu(tri, ent, rel, ratio, margin)
    assert loss.dim() == 0 and loss.dtype == torch.float32
    # the reference's own fp32 run sits 1.6e-5 from its fp64 run on the hub fixture (1884 sequential fp32 adds)
    for tag, tol in (("f64", TIGHT), ("f32", 5e-5)):
        assert abs(float(loss) - float(g[f"{tag}.loss"])) <= TIGHT * abs(float(g[f"{tag}.loss"])), (tag, float(loss))
        assert rel_l2(e.grad, g[f"{tag}.grad.entity_embed"]) <= tol, tag
        assert rel_l2(r.grad, g[f"{tag}.grad.relation_embed"]) <= tol, tag
    # entities / relations that occur in no triple get an exactly-zero dense gradient row
    used = torch.zeros(ent.shape[0], dtype=torch.bool); used[tri[:, 0]] = True; used[tri[:, 2]] = True
    assert float(e.grad.cpu()[~used].abs().sum()) == 0.0
    # the SGD step of main.py:524
    from recon_b200.loss import sgd_step
    pe = torch.nn.Parameter(ent.to(dev()).clone()); pe.grad = e.grad.clone()
    pr = torch.nn.Parameter(rel.to(dev()).clone()); pr.grad = r.grad.clone()
    sgd_step([pe, pr], float(g["lr"]))
    torch.cuda.synchronize()
    assert rel_l2(pe, g["f64.sgd.entity_embed"]) <= 1e-6
    assert rel_l2(pr, g["f64.sgd.relation_embed"]) <= 1e-6


@pytest.mark.gpu
@pytest.mark.parametrize("n_ent,n_rel,width,n_pos,ratio,margin,hub_share",
                         [(5000, 7, 200, 40000, 2, 5.0, 0.3),      # relation segments of ~28k, one entity heads 12k positives
                          (3000, 40, 100, 9000, 1, 1.0, 0.0),
                          (2000, 3, 36, 2500, 4, 0.25, 0.6),
                          (500, 2, 260, 600, 2, 2.0, 0.0)])         # width > 256: 3 chunks per lane
def test_loss_vs_oracle_seeded(n_ent, n_rel, width, n_pos, ratio, margin, hub_share):
    from oracle import loss as OL
    from recon_b200.loss import TripleIncidence
    gen = torch.Generator().manual_seed(7)
    ent = torch.nn.functional.normalize(torch.randn(n_ent, width, generator=gen), dim=1)
    rel = torch.randn(n_rel, width, generator=gen) * 0.1
    tri = OL.make_train_indices(n_ent, n_rel, n_pos, ratio, seed=8, hub_entity=11, hub_share=hub_share)
    margin = _margin_without_ties(tri, ent, rel, ratio, margin)
    loss, e, r = _run_gpu(tri, ent, rel, ratio, margin)
    ref_loss, ref_de, ref_dr = OL.loss_fwd_bwd(tri, ent.double(), rel.double(), ratio, margin)
    assert abs(float(loss) - float(ref_loss)) <= TIGHT * abs(float(ref_loss))
    err_e, err_r = rel_l2(e.grad, ref_de), rel_l2(r.grad, ref_dr)
    assert err_e <= TIGHT and err_r <= TIGHT, (err_e, err_r)
    inc = TripleIncidence(tri.to(dev()), n_ent, n_rel)
    if n_pos >= 9000 and n_rel <= 7:
        assert inc.rel_hubs.n_hubs > 0                                # hub tasks + finalize were exercised
    if hub_share > 0:
        assert inc.ent_hubs.n_hubs > 0
    # sum reduction
    loss_s, e_s, _ = _run_gpu(tri, ent, rel, ratio, margin, reduction="sum")
    m = tri.shape[0] - n_pos
    assert abs(float(loss_s) - float(ref_loss) * m) <= TIGHT * abs(float(ref_loss) * m)
    assert rel_l2(e_s.grad, ref_de * m) <= TIGHT


@pytest.mark.gpu
def test_loss_incidence_layout_bit_exact():
    """Sorted incidence lists = stable sort of (head, tail) / relation ids: exact integers."""
    from oracle.loss import make_train_indices
    from recon_b200.loss import TripleIncidence
    n_ent, n_rel = 700, 9
    tri = make_train_indices(n_ent, n_rel, 4000, 2, seed=3, hub_entity=5, hub_share=0.2)
    inc = TripleIncidence(tri.to(dev()), n_ent, n_rel)
    keys = tri[:, [0, 2]].reshape(-1)
    order = torch.sort(keys, stable=True).indices
    assert torch.equal(inc.ent_inc.cpu().long(), order)
    assert torch.equal(inc.ent_ptr.cpu().long(), torch.cat((torch.zeros(1, dtype=torch.long),
                                                          torch.cumsum(torch.bincount(keys, minlength=n_ent), 0))))
    rorder = torch.sort(tri[:, 1], stable=True).indices
    assert torch.equal(inc.rel_inc.cpu().long(), rorder)
    assert torch.equal(inc.rel_ptr.cpu().long(), torch.cat((torch.zeros(1, dtype=torch.long),
                                                          torch.cumsum(torch.bincount(tri[:, 1], minlength=n_rel), 0))))


@pytest.mark.gpu
def test_loss_run_to_run_bit_identical_and_scaled_upstream():
    from oracle.loss import make_train_indices
    from recon_b200.loss import batch_gat_loss
    gen = torch.Generator().manual_seed(1)
    ent = torch.randn(4000, 200, generator=gen); rel = torch.randn(5, 200, generator=gen)
    tri = make_train_indices(4000, 5, 30000, 2, seed=2, hub_entity=0, hub_share=0.5)
    a = _run_gpu(tri, ent, rel, 2, 5.0)
    b = _run_gpu(tri, ent, rel, 2, 5.0)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1].grad, b[1].grad) and torch.equal(a[2].grad, b[2].grad)
    # upstream gradient != 1 (loss scaled by the caller)
    e = ent.to(dev()).requires_grad_(True); r = rel.to(dev()).requires_grad_(True)
    (batch_gat_loss(torch.nn.MarginRankingLoss(margin=5.0), tri.to(dev()), e, r) * 0.25).backward()
    assert rel_l2(e.grad, a[1].grad * 0.25) <= 1e-7 and rel_l2(r.grad, a[2].grad * 0.25) <= 1e-7


@pytest.mark.gpu
def test_loss_errors_like_reference():
    from recon_b200.loss import batch_gat_loss
    _, tri, ent, rel, ratio, margin = _case("loss_small")
    f = torch.nn.MarginRankingLoss(margin=margin)
    e, r = ent.to(dev()), rel.to(dev())
    with pytest.raises(RuntimeError):                       # 149 rows are not (2*ratio+1) x P: the reference cannot broadcast
        batch_gat_loss(f, tri[:-1].to(dev()), e, r, valid_invalid_ratio_gat=ratio)
    bad = tri.clone(); bad[3, 2] = ent.shape[0]
    with pytest.raises(IndexError):                         # entity_embed[...] out of range, main.py:353-355
        batch_gat_loss(f, bad.to(dev()), e, r, valid_invalid_ratio_gat=ratio)
    nan_ent = e.clone(); nan_ent[int(tri[0, 0]), 0] = float("nan")
    with pytest.raises(AssertionError):                     # main.py:374
        batch_gat_loss(f, tri.to(dev()), nan_ent, r, valid_invalid_ratio_gat=ratio)
    # host int64 indices are accepted like everywhere else at the boundary
    loss = batch_gat_loss(f, tri, e, r, valid_invalid_ratio_gat=ratio)
    assert abs(float(loss) - float(load_golden("loss_small")["f64.loss"])) < 1e-5


@pytest.mark.gpu
def test_training_iteration_model_loss_sgd_vs_oracle():
    """One full iteration of train_gat (main.py:516-524): model forward -> batch_gat_loss -> backward -> SGD,
    against the fp64 oracle of the model chained with the fp64 oracle of the loss."""
    from recon_b200 import SpKBGATModified
    from recon_b200.loss import batch_gat_loss, sgd_step
    from recon_b200.synth import make_kg
    from oracle import ref_torch as O
    from oracle import loss as OL
    n, e, r, f, d, h = 3000, 30000, 20, 50, 100, 2
    edge, etype, nhop = make_kg(n, e, r, 1.1, 3000, seed=21)
    p = O.init_params(n, r, f, d, h, seed=21)
    tri = OL.make_train_indices(n, r, 4000, 2, seed=22)
    be = torch.unique(tri[:4000, 0])
    model = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [d, 2 * d], [d, 2 * d],
                            0.0, 0.2, [h, h], None)
    model.load_state_dict(p)
    model = model.to(dev())
    out_e, out_r, _ = model(None, be, (edge, etype), nhop)
    loss = batch_gat_loss(torch.nn.MarginRankingLoss(margin=5.0), tri, out_e, out_r)
    loss.backward()
    # oracle
    names = O.trainable_names(p)
    q = {k: (v.double().clone().requires_grad_(True) if k in names else v.double()) for k, v in p.items()}
    o_e, o_r, _, _ = O.kbgat_forward(q, be, (edge, etype), nhop, 0.2, None, O.seg_sum_index_add)
    ref_loss = OL.batch_gat_loss(tri, o_e, o_r, 2, 5.0)
    ref_loss.backward()
    assert abs(float(loss) - float(ref_loss)) <= TIGHT * abs(float(ref_loss))
    errs = {nm: rel_l2(prm.grad, q[nm].grad) for nm, prm in model.named_parameters() if nm in names and q[nm].grad is not None}
    assert len(errs) >= 10
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    lr = 1e-3
    before = {nm: prm.detach().clone() for nm, prm in model.named_parameters()}
    sgd_step(model.parameters(), lr)
    torch.cuda.synchronize()
    for nm, prm in model.named_parameters():
        if prm.grad is not None:
            want = before[nm].double() - lr * prm.grad.double()
            assert rel_l2(prm, want) <= 1e-6, nm
