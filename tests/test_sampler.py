"""N3: Corpus.get_iteration_triples_batch.

CPU: oracle/sampler.py reproduces the reference's own method bit for bit under the same numpy seed (fixtures written by
the reference's Corpus in tests/golden/make_golden.py).
GPU: recon_b200.sampler through the C ABI -- positives bit-exact (values and order); with the reference's initial draws
supplied every row the reference did not have to redraw is bit-exact, every other row obeys the reference's rule
(not a valid triple, only the designated column differs from its positive); layout, values, determinism."""
import numpy as np
import pytest
import torch

from helpers import load_golden

CASES = ["sampler_a", "sampler_dense_r1", "sampler_ratio3", "sampler_ratio1"]


def _valid_set(tri):
    return set(tuple(int(v) for v in row) for row in np.asarray(tri).tolist())


@pytest.mark.parametrize("name", CASES)
def test_sampler_oracle_matches_reference(name):
    from oracle import sampler as OS, edges as OE
    g = load_golden(name)
    rows, cols, data = OE.triples_to_adj(g["triples"].tolist())
    graph = OE.build_graph(rows, cols, data)
    np.random.seed(int(g["np_seed"]))
    trace = {}
    idx, val = OS.get_iteration_triples_batch(graph, g["batch"].tolist(), _valid_set(g["triples"]), int(g["n_entities"]),
                                              int(g["n_relations"]), int(g["ratio"]), trace=trace)
    assert idx.dtype == np.int32 and val.dtype == np.float32
    assert np.array_equal(idx, g["batch_indices"]) and np.array_equal(val, g["batch_values"])
    assert np.array_equal(trace["random_entities"], g["init_entities"])
    assert np.array_equal(trace["random_relations"], g["init_relations"])


def _check_rules(idx, val, p, ratio, valid, n_ent, n_rel):
    """The reference's row rules (create_batch.py:298-347) for any choice of random numbers."""
    idx = np.asarray(idx); val = np.asarray(val).reshape(-1)
    assert idx.shape == (p * (2 * ratio + 1), 3) and val.shape == (p * (2 * ratio + 1),)
    assert (val[:p] == 1).all()
    half = ratio // 2
    for q in range(2 * ratio * p):
        row, src = idx[p + q], idx[q % p]
        if q < p * half:
            col = 0
        elif q < 2 * p * half:
            col = 2
        elif q < p * ratio:
            col = None
        else:
            col = 1
        others = [c for c in range(3) if c != col]
        assert (row[others] == src[others]).all(), (q, row, src)
        if col is None:
            assert val[p + q] == 1
        elif col in (0, 2):
            assert val[p + q] == -1 and 0 <= row[col] < n_ent
            assert tuple(int(v) for v in row) not in valid, (q, row)
        else:
            assert 0 <= row[1] < n_rel
            if val[p + q] == -1:
                assert tuple(int(v) for v in row) not in valid, (q, row)
            else:                                   # gave up after n_rel redraws: untouched +1 copy (create_batch.py:339-345)
                assert val[p + q] == 1 and row[1] == src[1]


@pytest.mark.parametrize("name", CASES)
def test_reference_fixture_obeys_the_rules_checker(name):
    """The property checker used for the GPU path accepts the reference's own output."""
    g = load_golden(name)
    ratio = int(g["ratio"])
    p = g["batch_indices"].shape[0] // (2 * ratio + 1)
    _check_rules(g["batch_indices"], g["batch_values"], p, ratio, _valid_set(g["triples"]), int(g["n_entities"]),
                 int(g["n_relations"]))


def test_sampler_has_no_cpu_fallback():
    from recon_b200.sampler import TripleSampler
    if torch.cuda.is_available():
        pytest.skip("checks the no-GPU behaviour")
    g = load_golden("sampler_a")
    with pytest.raises(RuntimeError):
        TripleSampler(torch.as_tensor(g["triples"]), int(g["n_entities"]), int(g["n_relations"]), device=torch.device("cpu"))


# ---- GPU -------------------------------------------------------------------------------------------------------------
def dev():
    return torch.device("cuda:0")


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_sampler_golden(name):
    from recon_b200.sampler import TripleSampler
    g = load_golden(name)
    n_ent, n_rel, ratio = int(g["n_entities"]), int(g["n_relations"]), int(g["ratio"])
    tri = torch.as_tensor(g["triples"])
    smp = TripleSampler(tri, n_ent, n_rel, invalid_valid_ratio=ratio, device=dev())
    ref_idx, ref_val = g["batch_indices"], g["batch_values"]
    p = ref_idx.shape[0] // (2 * ratio + 1)
    pos = smp.positive_triples(g["batch"].tolist())
    assert pos.dtype == torch.int64 and np.array_equal(pos.cpu().numpy(), ref_idx[:p])          # values AND order
    idx, val = smp.get_iteration_triples_batch(g["batch"].tolist(), random_entities=g["init_entities"],
                                               random_relations=g["init_relations"], seed=5)
    idx, val = idx.cpu().numpy(), val.cpu().numpy()
    assert idx.shape == ref_idx.shape and val.shape == ref_val.shape and val.dtype == np.float32
    valid = _valid_set(g["triples"])
    _check_rules(idx, val, p, ratio, valid, n_ent, n_rel)
    # rows whose first candidate the reference accepted are bit-identical
    half = ratio // 2
    init_e, init_r = g["init_entities"], g["init_relations"]
    same = 0
    for q in range(2 * ratio * p):
        if q < p * half:
            untouched = ref_idx[p + q, 0] == init_e[q] and (int(init_e[q]), int(ref_idx[q % p, 1]), int(ref_idx[q % p, 2])) not in valid
        elif q < 2 * p * half:
            untouched = ref_idx[p + q, 2] == init_e[q] and (int(ref_idx[q % p, 0]), int(ref_idx[q % p, 1]), int(init_e[q])) not in valid
        elif q < p * ratio:
            untouched = True
        else:
            cr = q - p * ratio
            untouched = (int(ref_idx[q % p, 0]), int(init_r[cr]), int(ref_idx[q % p, 2])) not in valid
        if untouched:
            same += 1
            assert (idx[p + q] == ref_idx[p + q]).all() and val[p + q] == ref_val[p + q], q
    assert same > 0
    if name == "sampler_dense_r1":                    # n_rel == 1: every relation slot exhausts and stays a +1 copy
        assert (val[p + p * ratio:] == 1).all() and (ref_val[p + p * ratio:] == 1).all()


@pytest.mark.gpu
def test_sampler_generated_draws_rules_determinism_uniformity():
    from recon_b200.sampler import TripleSampler
    from recon_b200.synth import make_triples
    n_ent, n_rel, ratio = 3000, 12, 2
    tri = make_triples(n_ent, 60000, n_rel, seed=3)
    extra = make_triples(n_ent, 5000, n_rel, seed=4)                                   # "validation + test" triples
    smp = TripleSampler(tri, n_ent, n_rel, valid_triples=torch.cat((tri, extra)), invalid_valid_ratio=ratio, device=dev())
    batch = torch.randperm(n_ent, generator=torch.Generator().manual_seed(1))[:500].tolist()
    a_idx, a_val = smp.get_iteration_triples_batch(batch, seed=11)
    b_idx, b_val = smp.get_iteration_triples_batch(batch, seed=11)
    c_idx, _ = smp.get_iteration_triples_batch(batch, seed=12)
    assert torch.equal(a_idx, b_idx) and torch.equal(a_val, b_val)
    assert not torch.equal(a_idx, c_idx)
    p = a_idx.shape[0] // (2 * ratio + 1)
    assert p > 5000
    valid = _valid_set(torch.cat((tri, extra)).numpy())
    _check_rules(a_idx.cpu().numpy(), a_val.cpu().numpy(), p, ratio, valid, n_ent, n_rel)
    heads = a_idx[p:2 * p, 0].cpu().numpy()
    counts = np.bincount(heads, minlength=n_ent)
    assert counts.max() <= 25 and (counts > 0).sum() > 0.6 * n_ent                   # p/n_ent ~ 3.5 per entity on average
    rels = a_idx[p + p * ratio:, 1].cpu().numpy()
    rc = np.bincount(rels, minlength=n_rel) / rels.size
    assert np.abs(rc - 1.0 / n_rel).max() < 0.02
    # empty batch / entities without out-edges
    e_idx, e_val = smp.get_iteration_triples_batch([], seed=1)
    assert e_idx.shape == (0, 3) and e_val.shape == (0, 1)
    with pytest.raises(ValueError):
        smp.get_iteration_triples_batch(batch, random_entities=[1, 2, 3])


@pytest.mark.gpu
def test_sampler_feeds_a_training_iteration():
    """main.py:488-524 with every stage on the device: batch adjacency + 2-hop rows + triples -> model -> loss -> SGD."""
    from recon_b200 import SpKBGATModified
    from recon_b200.loss import batch_gat_loss, sgd_step
    from recon_b200.sampler import TripleSampler
    from recon_b200.synth import make_triples
    from oracle import ref_torch as O, loss as OL
    n, r, f, d, h = 800, 9, 50, 100, 2
    tri = make_triples(n, 6000, r, seed=8)
    smp = TripleSampler(tri, n, r, invalid_valid_ratio=2, device=dev())
    batch = torch.randperm(n, generator=torch.Generator().manual_seed(2))[:100].tolist()
    adj_idx, adj_val, nhop = smp.graph.batch_edges(batch)
    train_indices, train_values = smp.get_iteration_triples_batch(batch, seed=3)
    p0 = O.init_params(n, r, f, d, h, seed=4)
    model = SpKBGATModified(p0["entity_embeddings"].clone(), p0["relation_embeddings"].clone(), [d, 2 * d], [d, 2 * d],
                            0.0, 0.2, [h, h], None)
    model.load_state_dict(p0)
    model = model.to(dev())
    out_e, out_r, _ = model(None, torch.tensor(batch), (adj_idx, adj_val), nhop.long())
    loss = batch_gat_loss(torch.nn.MarginRankingLoss(margin=5.0), train_indices, out_e, out_r)
    loss.backward()
    names = O.trainable_names(p0)
    q = {k: (v.double().clone().requires_grad_(True) if k in names else v.double()) for k, v in p0.items()}
    o_e, o_r, _, _ = O.kbgat_forward(q, torch.tensor(batch), (adj_idx.cpu(), adj_val.cpu()), nhop.long().cpu(), 0.2, None,
                                     O.seg_sum_index_add)
    ref = OL.batch_gat_loss(train_indices.cpu(), o_e, o_r, 2, 5.0)
    ref.backward()
    assert abs(float(loss) - float(ref)) <= 2e-5 * abs(float(ref))
    worst = max(float((prm.grad.cpu().double() - q[nm].grad).norm() / q[nm].grad.norm())
                for nm, prm in model.named_parameters() if nm in names and q[nm].grad is not None and float(q[nm].grad.norm()) > 0)
    assert worst < 1e-4, worst
    sgd_step(model.parameters(), 1e-3)
    torch.cuda.synchronize()
