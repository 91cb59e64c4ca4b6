"""Full-size parity (-m gpu): BASELINE.json's C2 shape (N=2M entities, E=20M triples, R=1k) and a 2-hop shape.

The fp64 oracle cannot hold a 20M-edge graph (the reference algorithm materialises [2F+Rd, E]), so parity at full
size goes through a size-independent property of message passing, *locality*: with the loss restricted to a set S of
output rows, outputs and ALL gradients depend only on the edges into S and into the nodes S gathers from (two layers =
two hops). The same rows are therefore computed twice:
  * by the CUDA path on the FULL graph (every kernel runs at full size: 32-bit index ranges, hub tasks, grid sizes),
  * by the CPU fp64 oracle on the compacted 2-hop sub-problem (a few 10^4..10^5 edges),
and must agree to the north_star tolerance; the gradient of every entity row outside the sub-problem must be exactly 0.
Also at full size: the integer layouts (stable sort by row = sortedness + permutation + tie order) and bit-identical
re-runs.
"""
import pytest
import torch

from helpers import rel_l2

pytestmark = pytest.mark.gpu
TOL = 1e-4
F_IN, D_OUT, HEADS, ALPHA = 50, 100, 2, 0.2


def dev():
    return torch.device("cuda:0")


def _pick_rows(edge, nhop, n, seed, n_rows=48, max_sub_edges=400_000):
    """S = random rows + one hub-path row (in-degree in [600, 20000], i.e. > the 512-edge task threshold) whose 2-hop
    sub-problem stays small enough for the CPU oracle."""
    rows_all = torch.cat((edge[0], nhop[:, 3])) if nhop.numel() else edge[0]
    cols_all = torch.cat((edge[1], nhop[:, 0])) if nhop.numel() else edge[1]
    deg = torch.bincount(rows_all, minlength=n)
    cand = ((deg >= 600) & (deg <= 20000)).nonzero().flatten()
    g = torch.Generator().manual_seed(seed)
    for _ in range(20):
        s = torch.randperm(n, generator=g)[:n_rows]
        if cand.numel():
            s = torch.cat((s, cand[torch.randint(0, cand.numel(), (1,), generator=g)]))
        s = torch.unique(s)
        in_s = torch.zeros(n, dtype=torch.bool); in_s[s] = True
        a = in_s.clone(); a[cols_all[in_s[rows_all]]] = True                 # A = S + everything S gathers
        sub = a[rows_all]                                                     # every edge into A
        if int(sub.sum()) <= max_sub_edges:
            return s, a, sub
    raise RuntimeError("no small enough sub-problem found")


def _compact(edge, etype, nhop, sub_mask, n):
    e1 = edge.shape[1]
    m1, m2 = sub_mask[:e1], sub_mask[e1:]
    sedge, stype, snhop = edge[:, m1], etype[m1], nhop[m2] if nhop.numel() else nhop
    used = torch.zeros(n, dtype=torch.bool)
    used[sedge[0]] = True; used[sedge[1]] = True
    if snhop.numel():
        used[snhop[:, 0]] = True; used[snhop[:, 3]] = True
    v = used.nonzero().flatten()
    remap = torch.full((n,), -1, dtype=torch.int64); remap[v] = torch.arange(v.numel())
    sedge = remap[sedge]
    if snhop.numel():
        snhop = torch.stack((remap[snhop[:, 0]], snhop[:, 1], snhop[:, 2], remap[snhop[:, 3]]), dim=1)
    return v, remap, sedge, stype, snhop


@pytest.mark.parametrize("n,e1,e2,r,hub_frac", [(2_000_000, 20_000_000, 0, 1000, 0.2),        # C2 (BASELINE configs[1])
                                                (500_000, 4_000_000, 8_000_000, 1000, 0.2)])  # C3's 1:2 hop mix, quarter size
def test_full_size_locality_parity(n, e1, e2, r, hub_frac):
    from recon_b200 import SpKBGATModified
    from recon_b200.synth import make_kg
    from oracle import ref_torch as O
    edge, etype, nhop = make_kg(n, e1, r, 1.1, e2, seed=0, hub_frac=hub_frac)
    p = O.init_params(n, r, F_IN, D_OUT, HEADS, seed=5)
    s, a_mask, sub_mask = _pick_rows(edge, nhop, n, seed=6)
    gen = torch.Generator().manual_seed(9)
    g_s = torch.randn(s.numel(), D_OUT * HEADS, generator=gen)
    g_rel = torch.randn(r, D_OUT * HEADS, generator=gen)

    model = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [D_OUT, 2 * D_OUT],
                            [D_OUT, 2 * D_OUT], 0.0, ALPHA, [HEADS, HEADS], None)
    model.load_state_dict(p)
    model = model.to(dev())
    d_edge, d_type, d_nhop = edge.to(dev()), etype.to(dev()), nhop.to(dev())
    graph = model.prepare_graph((d_edge, d_type), d_nhop)

    # ---- integer layouts at full size: stable sort by aggregation row ----
    rows_all = torch.cat((d_edge[0], d_nhop[:, 3])) if e2 else d_edge[0]
    order = torch.sort(rows_all, stable=True).indices
    assert torch.equal(graph.perm.long(), order)
    assert torch.equal(graph.row.long(), rows_all[order])
    cols_all = torch.cat((d_edge[1], d_nhop[:, 0])) if e2 else d_edge[1]
    assert torch.equal(graph.col.long(), cols_all[order])
    deg = torch.bincount(rows_all, minlength=n)
    assert torch.equal(graph.rowptr.long()[1:], torch.cumsum(deg, 0)) and int(graph.rowptr[0]) == 0
    assert graph.row_hubs.n_hubs > 0 and int(deg.max()) > 100_000          # power-law rows really are there
    del order, rows_all, cols_all, deg

    def run():
        model.load_state_dict(p)          # forward overwrites entity_embeddings with its normalised rows (models.py:160-161)
        model.zero_grad(set_to_none=True)
        out_e, out_r, mask = model(None, s.to(dev()), graph, None)
        loss = (out_e[s.to(dev())] * g_s.to(dev())).sum() + (out_r * g_rel.to(dev())).sum()
        loss.backward()
        torch.cuda.synchronize()
        return out_e.detach().clone(), out_r.detach().clone(), mask, {k: v.grad.detach().clone() for k, v in model.named_parameters()
                                                      if v.grad is not None}

    out_e, out_r, mask, grads = run()
    assert int(mask.sum().item()) == s.numel()

    # ---- the same rows from the fp64 oracle on the compacted 2-hop sub-problem ----
    v, remap, sedge, stype, snhop = _compact(edge, etype, nhop, sub_mask, n)
    q = {k: val.double() for k, val in p.items()}
    q["entity_embeddings"] = q["entity_embeddings"][v].clone()
    q["final_entity_embeddings"] = q["final_entity_embeddings"][v].clone()
    names = O.trainable_names(q)
    q = {k: (val.clone().requires_grad_(True) if k in names else val) for k, val in q.items()}
    s_loc = remap[s]
    o_e, o_r, _, _ = O.kbgat_forward(q, s_loc, (sedge, stype), snhop if snhop.numel() else None, ALPHA, None,
                                     O.seg_sum_index_add)
    ((o_e[s_loc] * g_s.double()).sum() + (o_r * g_rel.double()).sum()).backward()

    errs = {"out_entity[S]": rel_l2(out_e[s.to(dev())], o_e[s_loc].detach()), "out_relation": rel_l2(out_r, o_r.detach())}
    for nm in names:
        ref = q[nm].grad
        got = grads[nm]
        if nm == "entity_embeddings":
            outside = torch.ones(n, dtype=torch.bool); outside[v] = False
            assert float(got[outside.to(dev())].abs().sum()) == 0.0        # nothing leaks outside the 2-hop sub-problem
            got = got[v.to(dev())]
        errs["grad." + nm] = rel_l2(got, ref)
    assert len(errs) >= 12
    bad = {k: e for k, e in errs.items() if not e < TOL}
    assert not bad, bad

    # ---- re-run on the same layouts: bit-identical (no atomics anywhere) ----
    out_e2, out_r2, _, grads2 = run()
    assert torch.equal(out_e, out_e2) and torch.equal(out_r, out_r2)
    for k in grads:
        assert torch.equal(grads[k], grads2[k]), k
