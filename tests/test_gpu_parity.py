"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against the golden
fixtures produced by the reference and against the CPU oracle on seeded inputs.

Tolerance: north_star asks rel-err <= 1e-4 on outputs and gradients against the reference fp32 path; we
compare against the reference's fp64 run (tests/golden, or oracle in fp64) with rel-L2 <= 1e-4 and
additionally expect ~1e-6 because every product is fp32-exact (SURVEY.md 8c calibration)."""
import numpy as np
import pytest
import torch

from helpers import load_golden, rel_l2, golden_params, golden_masks, MODEL_CASES

pytestmark = pytest.mark.gpu
TOL = 1e-4          # north_star tolerance
TIGHT = 2e-5        # what fp32-exact kernels should reach


def dev():
    return torch.device("cuda:0")


def build_model(g, p_drop=0.0):
    from recon_b200 import SpKBGATModified
    p = golden_params(g)
    n, f = p["entity_embeddings"].shape
    r = p["relation_embeddings"].shape[0]
    heads = sum(1 for k in p if k.startswith("sparse_gat_1.attention_") and k.endswith(".a"))
    d = p["sparse_gat_1.attention_0.a"].shape[0]
    m = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [d, 2 * d], [d, 2 * d],
                        p_drop, 0.2, [heads, heads], None)
    m.load_state_dict(p)
    return m.to(dev())


# ---- K0: graph construction ----------------------------------------------------------------------
@pytest.mark.parametrize("n,bits", [(0, 5), (1, 1), (33, 3), (5000, 13), (300000, 21), (1 << 20, 31)])
def test_radix_sort_pairs_is_stable_and_exact(n, bits):
    from recon_b200.graph import sort_pairs
    g = torch.Generator().manual_seed(n)
    keys = torch.randint(0, 1 << min(bits, 30), (n,), generator=g, dtype=torch.int32)
    vals = torch.arange(n, dtype=torch.int32)
    k2, v2 = sort_pairs(keys.to(dev()).clone(), vals.to(dev()).clone(), bits)
    ks, order = torch.sort(keys, stable=True)
    assert torch.equal(k2.cpu(), ks)
    assert torch.equal(v2.cpu().long(), order)


@pytest.mark.parametrize("alpha,n_nhop", [(None, 0), (1.1, 700), (1.5, 0)])
def test_kgraph_layouts_bit_exact(alpha, n_nhop):
    from recon_b200 import KGraph
    from recon_b200.synth import make_kg
    n, e, r = 3000, 40000, 17
    edge, etype, nhop = make_kg(n, e, r, alpha, n_nhop, seed=5)
    g = KGraph(edge, etype, nhop, n, r, device=dev())
    rows = torch.cat((edge[0], nhop[:, 3])); cols = torch.cat((edge[1], nhop[:, 0]))
    t1 = torch.cat((etype, nhop[:, 1])); t2 = torch.cat((torch.full((e,), -1, dtype=torch.long), nhop[:, 2]))
    srow, perm = torch.sort(rows, stable=True)
    assert torch.equal(g.perm.cpu().long(), perm)
    assert torch.equal(g.row.cpu().long(), srow)
    assert torch.equal(g.col.cpu().long(), cols[perm])
    assert torch.equal(g.t1.cpu().long(), t1[perm])
    if n_nhop:
        assert torch.equal(g.t2.cpu().long(), t2[perm])
    else:
        assert g.t2 is None
    rowptr = torch.searchsorted(srow, torch.arange(n + 1))
    assert torch.equal(g.rowptr.cpu().long(), rowptr)
    ccol, cpos = torch.sort(cols[perm], stable=True)
    assert torch.equal(g.csc_pos.cpu().long(), cpos)
    assert torch.equal(g.csc_row.cpu().long(), srow[cpos])
    assert torch.equal(g.colptr.cpu().long(), torch.searchsorted(ccol, torch.arange(n + 1)))
    # relation incidence: every (relation, edge) pair exactly once, segments stable
    relptr = g.relptr.cpu().long(); rpos = g.rel_pos.cpu().long()
    t1c, t2c = t1[perm], t2[perm]
    for k in range(r):
        seg = rpos[relptr[k]:relptr[k + 1]]
        want = torch.cat(((t1c == k).nonzero().flatten(), (t2c == k).nonzero().flatten()))
        assert torch.equal(seg, want)
    # hub tasks tile every hub segment exactly
    hubs = g.row_hubs
    deg = rowptr[1:] - rowptr[:-1]
    assert hubs.n_hubs == int((deg > hubs.thresh).sum())
    if hubs.n_tasks:
        tb, te, ts = hubs.task_beg.cpu().long(), hubs.task_end.cpu().long(), hubs.task_seg.cpu().long()
        assert int((te - tb).sum()) == int(deg[deg > hubs.thresh].sum())
        assert bool(((tb >= rowptr[ts]) & (te <= rowptr[ts + 1]) & (te > tb)).all())


def test_kgraph_rejects_out_of_range():
    from recon_b200 import KGraph
    edge = torch.tensor([[0, 5], [1, 2]]); et = torch.tensor([0, 0])
    with pytest.raises(IndexError):
        KGraph(edge, et, None, 4, 2, device=dev())
    with pytest.raises(IndexError):
        KGraph(torch.tensor([[0, 1], [1, 2]]), torch.tensor([0, 3]), None, 4, 2, device=dev())


def test_kgraph_empty():
    from recon_b200 import KGraph
    g = KGraph(torch.zeros(2, 0, dtype=torch.long), torch.zeros(0, dtype=torch.long), None, 7, 3, device=dev())
    assert g.rowptr.cpu().tolist() == [0] * 8 and g.colptr.cpu().tolist() == [0] * 8


# ---- K1/K5: GEMMs ---------------------------------------------------------------------------------
@pytest.fixture(params=["tc", "simt", "tc-pair", "tc-cluster"])
def gemm_path(request, monkeypatch):
    """Run the test on the tcgen05 3xTF32 GEMM, on the exact-fp32 SIMT GEMM, and on the two experimental NN variants
    (pair-CTA cta_group::2 kernel, 2-CTA clusters with multicast weight tiles; both off by default, see spk_gemm_tc.cu)."""
    from recon_b200 import functional as SF
    old = SF.USE_TC
    SF.USE_TC = request.param != "simt"
    if request.param == "tc-pair":
        monkeypatch.setenv("SPK_TC_PAIR", "1")
    if request.param == "tc-cluster":
        monkeypatch.setenv("SPK_TC_CLUSTER", "2")
    yield "simt" if request.param == "simt" else "tc"
    SF.USE_TC = old


@pytest.mark.parametrize("m,k,n", [(1, 1, 1), (130, 50, 416), (1000, 200, 416), (257, 416, 50), (3, 7, 5), (4096, 64, 200),
                                   (20000, 52, 416), (70001, 200, 208), (5000, 416, 200), (9999, 12, 24)])
def test_gemm_nn(m, k, n, gemm_path):
    from recon_b200.functional import gemm_nn
    g = torch.Generator().manual_seed(m + k + n)
    a = torch.randn(m, k, generator=g); b = torch.randn(k, n, generator=g)
    tol = 2e-6 if gemm_path == "simt" else 1e-5      # 3xTF32: dropped lo*lo term + tensor-core accumulation
    c = gemm_nn(a.to(dev()), b.to(dev()))
    assert rel_l2(c, a.double() @ b.double()) < tol
    c0 = torch.randn(m, n, generator=g)
    c2 = gemm_nn(a.to(dev()), b.to(dev()), out=c0.to(dev()), accumulate=True)
    assert rel_l2(c2, c0.double() + a.double() @ b.double()) < tol


def test_batch_index_range_check_without_host_sync():
    """models.py:167-173 indexes the mask with the batch entities: out-of-range -> IndexError, negative = from the end. Device
    index tensors are checked by the kernel through the model's sticky flag word (no min / max read-back per step)."""
    from recon_b200 import SpKBGATModified
    from recon_b200 import functional as SF
    from recon_b200.synth import make_kg
    n, r = 60, 5
    edge, etype, _ = make_kg(n, 400, r, seed=2)
    model = SpKBGATModified(torch.randn(n, 8), torch.randn(r, 8), [6, 12], [6, 12], 0.0, 0.2, [2, 2], None).to(dev())
    for bad in (torch.tensor([0, 3, n]), torch.tensor([0, -n - 1])):
        with pytest.raises(IndexError):
            model(None, bad.to(dev()), (edge, etype), None)                 # device tensor: flagged by the kernel
        with pytest.raises(IndexError):
            model(None, bad, (edge, etype), None)                           # host tensor: checked on the host
    _, _, mask = model(None, torch.tensor([1, -1, 1], device=dev()), (edge, etype), None)   # the flag word is clean again
    want = torch.zeros(n); want[1] = 1.0; want[n - 1] = 1.0
    assert torch.equal(mask.cpu(), want)
    with pytest.raises(IndexError):
        SF.mask_from_index(torch.tensor([n + 5], device=dev()), n, dev())   # stand-alone call: private flag word
    assert torch.equal(SF.mask_from_index(torch.tensor([-2], device=dev()), n, dev()).cpu().nonzero().flatten(), torch.tensor([n - 2]))


def test_gemm_nn_partial_last_tile_keeps_every_update():
    """157 M tiles on 148 persistent CTAs with a partial last tile: the epilogue warps whose rows lie past M used to rewrite
    their staging buffers while the previous tile's last TMA reduce-add was still reading them (lost C += rows, 30 % of runs)."""
    from recon_b200.functional import gemm_nn
    m, k, n = 20000, 52, 416
    g = torch.Generator().manual_seed(7)
    a = torch.randn(m, k, generator=g); b = torch.randn(k, n, generator=g); c0 = torch.randn(m, n, generator=g)
    ad, bd, c0d = a.to(dev()), b.to(dev()), c0.to(dev())
    ref = c0.double() + a.double() @ b.double()
    for _ in range(12):
        c = gemm_nn(ad, bd, out=c0d.clone(), accumulate=True)
        assert rel_l2(c, ref) < 1e-5


def test_gemm_nn_tc_wide_dynamic_range():
    """3xTF32 must stay fp32-accurate when magnitudes vary over many binades."""
    from recon_b200.functional import gemm_nn
    g = torch.Generator().manual_seed(7)
    a = torch.randn(3000, 200, generator=g) * torch.exp(4 * torch.randn(3000, 200, generator=g))
    b = torch.randn(200, 416, generator=g) * torch.exp(4 * torch.randn(200, 416, generator=g))
    c = gemm_nn(a.to(dev()), b.to(dev()))
    assert rel_l2(c, a.double() @ b.double()) < 1e-5


def test_gemm_nn_strided_views(gemm_path):
    from recon_b200.functional import gemm_nn
    g = torch.Generator().manual_seed(0)
    big = torch.randn(300, 96, generator=g).to(dev()); b = torch.randn(40, 24, generator=g).to(dev())
    a = big[:, 8:48]
    out = torch.zeros(300, 64, device=dev())
    gemm_nn(a, b, out=out[:, 16:40])
    assert rel_l2(out[:, 16:40], a.double().cpu() @ b.double().cpu()) < 1e-5
    assert float(out[:, :16].abs().sum()) == 0.0 and float(out[:, 40:].abs().sum()) == 0.0


@pytest.mark.parametrize("m,ka,nb", [(1, 1, 1), (5000, 52, 416), (70000, 200, 416), (333, 8, 4), (100000, 12, 208),
                                     (40, 200, 416), (33000, 416, 200), (1025, 300, 24), (333, 7, 5)])
def test_gemm_tn_deterministic(m, ka, nb, gemm_path):
    from recon_b200.functional import gemm_tn
    g = torch.Generator().manual_seed(m)
    a = torch.randn(m, ka, generator=g); b = torch.randn(m, nb, generator=g)
    c = gemm_tn(a.to(dev()), b.to(dev()))
    assert rel_l2(c, a.double().t() @ b.double()) < (5e-6 if gemm_path == "simt" else 2e-5)
    c2 = gemm_tn(a.to(dev()), b.to(dev()))
    assert torch.equal(c, c2)
    c0 = torch.randn(ka, nb, generator=g)
    c3 = gemm_tn(a.to(dev()), b.to(dev()), out=c0.to(dev()), accumulate=True)
    assert rel_l2(c3, c0.double() + a.double().t() @ b.double()) < 2e-5


def test_gemm_tn_tc_padded_view():
    """X given as a [:, :50] view of a 52-wide buffer (what tc_friendly produces for 50-dim embeddings)."""
    from recon_b200.functional import gemm_tn
    g = torch.Generator().manual_seed(1)
    xp = torch.randn(9000, 52, generator=g).to(dev()); gp = torch.randn(9000, 416, generator=g).to(dev())
    x = xp[:, :50]
    c = gemm_tn(x, gp)
    assert c.shape == (50, 416)
    assert rel_l2(c, x.double().cpu().t() @ gp.double().cpu()) < 2e-5


@pytest.mark.parametrize("H,F,Rd,D", [(1, 12, 12, 16), (2, 50, 50, 100), (3, 10, 6, 7), (4, 200, 200, 50), (1, 200, 200, 200)])
def test_attn_weights_kernels_match_torch_formulation(H, F, Rd, D):
    """spk_attn_weights_fwd / _bwd (the extended weight matrices and the gradients back to a, a_2) against the
    differentiable torch formulation they replace (GAT/layers.py:100-105 parameters), both layouts."""
    from recon_b200 import functional as SF
    g = torch.Generator().manual_seed(H * 1000 + F + D)
    a = [torch.randn(D, 2 * F + Rd, generator=g).to(dev()).requires_grad_(True) for _ in range(H)]
    a2 = [torch.randn(1, D, generator=g).to(dev()).requires_grad_(True) for _ in range(H)]
    cases = [("proj", lambda fn: fn(a, a2, F, SF.Geometry(H, D)), SF.extended_weights, SF._extended_weights_torch)]
    if H <= SF.AGG_MAX_HEADS and SF.AggGeometry.supported(H, F, Rd):
        ag = SF.AggGeometry(H, F, Rd, D)
        cases.append(("agg", lambda fn: fn(a, a2, ag), SF.agg_weights, SF._agg_weights_torch))
    for name, call, lib_fn, torch_fn in cases:
        outs, refs = call(lib_fn), call(torch_fn)
        seeds = [torch.randn(o.shape, generator=g).to(dev()) for o in refs]
        for o, r in zip(outs, refs):
            assert o.shape == r.shape and rel_l2(o, r.double().cpu()) < 1e-6, name   # data movement + one dot per column
        got = torch.autograd.grad(outs, a + a2, seeds)
        want = torch.autograd.grad(refs, a + a2, seeds)
        for x, y in zip(got, want):
            assert rel_l2(x, y.double().cpu()) < 2e-6, name


# ---- stand-alone op and layer ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["spmm_f1", "spmm_f7"])
def test_special_spmm_golden(name):
    from recon_b200 import SpecialSpmmFunctionFinal
    g = load_golden(name)
    w = torch.as_tensor(g["w"]).to(dev()).requires_grad_(True)
    n, f = g["out"].shape
    out = SpecialSpmmFunctionFinal.apply(torch.as_tensor(g["edge"]), w, n, w.shape[0], f)
    (out * torch.as_tensor(g["g"]).to(dev())).sum().backward()
    assert rel_l2(out, g["out"]) < 1e-6
    assert np.array_equal(w.grad.cpu().numpy(), g["grad_w"])


@pytest.fixture(params=["auto", "force", "off"])
def agg_path(request):
    """Layer groups run on the project-then-gather kernels (K2/K3), on the aggregate-then-project kernels (K2'/K3',
    input-narrow layers) or, with "auto", on whichever moves fewer bytes; every model test is run on all three."""
    from recon_b200 import functional as SF
    old = SF.AGG_MODE
    SF.AGG_MODE = request.param
    yield request.param
    SF.AGG_MODE = old


@pytest.fixture(params=["split", "fused", "rows"])
def bwd_path(request):
    """Backward schedule of the projected layer groups (functional.BWD_MODE): split-dot column + relation passes
    (default; graphs with 2-hop edges fall back to "rows"), column-major fused pass, or K3 rows + K4 columns."""
    from recon_b200 import functional as SF
    old = SF.BWD_MODE
    SF.BWD_MODE = request.param
    yield request.param
    SF.BWD_MODE = old


@pytest.mark.parametrize("name", ["layer_concat", "layer_noconcat_nhop"])
def test_attention_layer_golden(name, agg_path):
    from recon_b200 import SpGraphAttentionLayer
    g = load_golden(name)
    n, f = g["x"].shape
    d, k = g["a"].shape
    layer = SpGraphAttentionLayer(n, f, d, k - 2 * f, 0.0, 0.2, bool(g["concat"])).to(dev())
    with torch.no_grad():
        layer.a.copy_(torch.as_tensor(g["a"])); layer.a_2.copy_(torch.as_tensor(g["a_2"]))
    x = torch.as_tensor(g["x"]).to(dev()).requires_grad_(True)
    emb = torch.as_tensor(g["edge_embed"]).to(dev()).requires_grad_(True)
    has2 = g["edge_nhop"].size > 0
    e2 = torch.as_tensor(g["edge_nhop"]).long() if has2 else torch.tensor([])
    emb2 = torch.as_tensor(g["edge_embed_nhop"]).to(dev()).requires_grad_(True) if has2 else torch.tensor([])
    out = layer(x, torch.as_tensor(g["edge"]), emb, e2, emb2)
    (out * torch.as_tensor(g["g"]).to(dev())).sum().backward()
    assert rel_l2(out, g["out"]) < TIGHT
    assert rel_l2(x.grad, g["grad.x"]) < TIGHT
    assert rel_l2(emb.grad, g["grad.edge_embed"]) < TIGHT
    assert rel_l2(layer.a.grad, g["grad.a"]) < TIGHT
    assert rel_l2(layer.a_2.grad, g["grad.a_2"]) < TIGHT
    if has2:
        assert rel_l2(emb2.grad, g["grad.edge_embed_nhop"]) < TIGHT


# ---- full model against the reference's own outputs ---------------------------------------------------
@pytest.mark.parametrize("name", MODEL_CASES)
def test_model_golden(name, agg_path):
    g = load_golden(name)
    model = build_model(g, float(g["p_drop"]))
    masks = golden_masks(g)
    adj = (torch.as_tensor(g["edge"]), torch.as_tensor(g["edge_type"]))
    nhop = torch.as_tensor(g["nhop"])
    be = torch.as_tensor(g["batch_entities"])
    ge, gr = torch.as_tensor(g["g_ent"]).to(dev()), torch.as_tensor(g["g_rel"]).to(dev())
    if name == "model_batch_test":
        out_e, out_r, mask = model.batch_test(None, be, adj, nhop, torch.as_tensor(g["entity_in"]).to(dev()),
                                              dropout_masks=masks)
    else:
        out_e, out_r, mask = model(None, be, adj, nhop, dropout_masks=masks)
    ((out_e * ge).sum() + (out_r * gr).sum()).backward()
    errs = {"out_entity": rel_l2(out_e, g["f64.out_entity"]), "out_relation": rel_l2(out_r, g["f64.out_relation"])}
    assert np.array_equal(mask.cpu().numpy(), g["f64.mask"].astype(np.float32))
    n_grads = 0
    for k, v in g.items():
        if k.startswith("f64.grad."):
            nm = k[len("f64.grad."):]
            prm = dict(model.named_parameters())[nm]
            assert prm.grad is not None, nm
            errs["grad." + nm] = rel_l2(prm.grad, v)
            n_grads += 1
    assert n_grads >= 5
    if name != "model_batch_test":      # the three .data side effects (models.py:160-161,181-183)
        errs["after.entity_embeddings"] = rel_l2(model.entity_embeddings.data, g["f64.after.entity_embeddings"])
        errs["after.final_entity"] = rel_l2(model.final_entity_embeddings.data, g["f64.after.final_entity_embeddings"])
        errs["after.final_relation"] = rel_l2(model.final_relation_embeddings.data, g["f64.after.final_relation_embeddings"])
    bad = {k: v for k, v in errs.items() if not v < TIGHT}
    assert not bad, bad
    # and against the reference's fp32 numbers at the north_star tolerance
    assert rel_l2(out_e, g["f32.out_entity"]) < TOL and rel_l2(out_r, g["f32.out_relation"]) < TOL


def test_state_dict_keys_match_reference():
    g = load_golden("model_small_uniform")
    model = build_model(g)
    want = {k[len("param."):]: v.shape for k, v in g.items() if k.startswith("param.")}
    got = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    assert got == {k: tuple(s) for k, s in want.items()}


# ---- seeded synthetic inputs against the oracle (C1 shape, uniform and Zipf with hub rows) ------------
@pytest.mark.parametrize("alpha,n_nhop,p_drop", [(None, 0, 0.0), (1.1, 0, 0.0), (1.1, 20000, 0.3)])
def test_model_vs_oracle_c1(alpha, n_nhop, p_drop, agg_path, bwd_path):
    from recon_b200 import SpKBGATModified
    from recon_b200.synth import make_kg
    from oracle import ref_torch as O
    n, e, r, f, d, h = 10000, 100000, 200, 50, 100, 2
    edge, etype, nhop = make_kg(n, e, r, alpha, n_nhop, seed=11)
    p = O.init_params(n, r, f, d, h, seed=11)
    gen = torch.Generator().manual_seed(12)
    ge, gr = torch.randn(n, d * h, generator=gen), torch.randn(r, d * h, generator=gen)
    masks = None
    if p_drop > 0:
        et = e + n_nhop
        mk = lambda *s: (torch.rand(*s, generator=gen) >= p_drop).float() / (1 - p_drop)   # noqa: E731
        masks = {"att": mk(h, et), "out": mk(et), "x": mk(n, d * h)}
    be = torch.arange(0, n, 3)
    model = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [d, 2 * d], [d, 2 * d],
                            p_drop, 0.2, [h, h], None)
    model.load_state_dict(p)
    model = model.to(dev())
    out_e, out_r, mask = model(None, be, (edge, etype), nhop, dropout_masks=masks)
    ((out_e * ge.to(dev())).sum() + (out_r * gr.to(dev())).sum()).backward()
    p64 = {k: v.double() for k, v in p.items()}
    m64 = None if masks is None else {k: v.double() for k, v in masks.items()}
    ref = O.fwd_bwd(p64, be, (edge, etype), nhop, 0.2, ge.double(), gr.double(), m64, O.seg_sum_index_add)
    errs = {"out_entity": rel_l2(out_e, ref[0]), "out_relation": rel_l2(out_r, ref[1])}
    for nm, prm in model.named_parameters():
        if nm in ref[4]:
            errs["grad." + nm] = rel_l2(prm.grad, ref[4][nm])
    assert len(errs) >= 10
    bad = {k: v for k, v in errs.items() if not v < TOL}
    assert not bad, bad
    if alpha is not None:
        assert model.prepare_graph((edge, etype), nhop).row_hubs.n_hubs > 0      # the hub path was exercised


def test_run_to_run_bit_identical(agg_path, bwd_path):
    from recon_b200 import SpKBGATModified
    from recon_b200.synth import make_kg
    from oracle import ref_torch as O
    n, e, r = 4000, 60000, 31
    edge, etype, nhop = make_kg(n, e, r, 1.1, 5000, seed=3)
    p = O.init_params(n, r, 50, 100, 2, seed=3)
    outs = []
    for _ in range(2):
        model = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [100, 200],
                                [100, 200], 0.0, 0.2, [2, 2], None)
        model.load_state_dict(p)
        model = model.to(dev())
        oe, orel, _ = model(None, torch.arange(n), (edge, etype), nhop)
        (oe.sum() + orel.sum()).backward()
        outs.append([oe.detach().clone()] + [prm.grad.clone() for prm in model.parameters() if prm.grad is not None])
    for a, b in zip(*outs):
        assert torch.equal(a, b)


def test_overflow_raises_assertion_like_reference(agg_path):
    """exp without max-subtraction (GAT/layers.py:143-146): huge scores overflow to inf -> NaN -> AssertionError."""
    from recon_b200 import SpKBGATModified
    from recon_b200.synth import make_kg
    from oracle import ref_torch as O
    n, e, r = 200, 2000, 5
    edge, etype, nhop = make_kg(n, e, r, None, 0, seed=1)
    p = O.init_params(n, r, 12, 8, 2, seed=1)
    p["relation_embeddings"] = p["relation_embeddings"] * 1e4
    model = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [8, 16], [8, 16], 0.0,
                            0.2, [2, 2], None)
    model.load_state_dict(p)
    model = model.to(dev())
    with pytest.raises(AssertionError):
        model(None, torch.arange(n), (edge, etype), nhop)


def test_training_mode_dropout_runs_and_is_unbiased_in_shape(agg_path):
    from recon_b200 import SpKBGATModified
    from recon_b200.synth import make_kg
    from oracle import ref_torch as O
    n, e, r = 300, 3000, 7
    edge, etype, nhop = make_kg(n, e, r, 1.5, 100, seed=2)
    p = O.init_params(n, r, 12, 8, 2, seed=2)
    model = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [8, 16], [8, 16], 0.3,
                            0.2, [2, 2], None)
    model.load_state_dict(p)
    model = model.to(dev()).train()
    oe, orel, _ = model(None, torch.arange(n), (edge, etype), nhop)
    (oe.sum() + orel.sum()).backward()
    assert torch.isfinite(oe).all() and all(torch.isfinite(q.grad).all() for q in model.parameters() if q.grad is not None)
    model.eval()
    a = model(None, torch.arange(n), (edge, etype), nhop)[0]
    b = model(None, torch.arange(n), (edge, etype), nhop)[0]
    # eval mode: no dropout. (Not bit-equal: every forward re-normalises entity_embeddings.data in place,
    # models.py:160-161, and normalize(normalize(x)) differs from normalize(x) in the last ulp.)
    assert torch.allclose(a, b, atol=1e-6)


# ---- K0b: batch adjacency + 2-hop rows, bit-exact with Corpus (values and order) -----------------------
@pytest.mark.parametrize("name", ["edges_a", "edges_b", "edges_partial"])
def test_batch_edges_match_corpus_golden(name):
    from recon_b200.nhop import TripleGraph
    g = load_golden(name)
    tr = torch.as_tensor(g["triples"])
    n = int(tr[:, [0, 2]].max()) + 1
    tg = TripleGraph(tr, n, device=dev())
    partial = bool(g["partial"])
    idx, val, nhop = tg.batch_edges(g["batch"].tolist(), partial)
    assert np.array_equal(idx.cpu().numpy(), g["adj_idx"])
    assert np.array_equal(val.cpu().numpy(), g["adj_val"])
    assert np.array_equal(nhop.cpu().numpy(), g["nhop"].astype(np.int32))
    fidx, fval, fnhop = tg.batch_edges(list(range(n)), partial)
    assert np.array_equal(fidx.cpu().numpy(), g["full_adj_idx"])
    assert np.array_equal(fval.cpu().numpy(), g["full_adj_val"])
    assert np.array_equal(fnhop.cpu().numpy(), g["full_nhop"].astype(np.int32))


def test_batch_edges_toy_kg():
    from recon_b200.nhop import build_batch_edges
    tr = torch.as_tensor(load_golden("edges_toy")["triples"]).to(dev())
    idx, val, nhop = build_batch_edges(tr, 10, [0, 1, 2, 3])
    assert idx.cpu().tolist() == [[1, 1, 2, 3, 0, 3, 4], [0, 0, 0, 1, 1, 2, 3]]
    assert val.cpu().tolist() == [5, 6, 7, 8, 2, 9, 1]
    assert nhop.cpu().tolist() == [[0, 5, 8, 3], [1, 8, 1, 4], [1, 2, 7, 2], [2, 9, 1, 4]]


def test_batch_edges_vs_oracle_random_medium():
    """Larger random KG with multi-edges / self loops against the pure-Python restatement of Corpus."""
    from recon_b200.nhop import TripleGraph
    from recon_b200.synth import make_triples
    from oracle import edges as OE
    n, t, r = 400, 3000, 9
    tr = make_triples(n, t, r, seed=9)
    graph = OE.build_graph(*OE.triples_to_adj(tr.tolist()))
    gen = torch.Generator().manual_seed(3)
    batch = torch.randperm(n, generator=gen)[:150].tolist() + [5, 5]       # duplicates allowed
    want_idx, want_val = OE.batch_adj(graph, batch)
    want_nhop = OE.batch_nhop(graph, batch)
    tg = TripleGraph(tr, n, device=dev())
    idx, val, nhop = tg.batch_edges(batch)
    assert idx.cpu().tolist() == want_idx and val.cpu().tolist() == want_val
    assert nhop.cpu().tolist() == want_nhop
    # spk_nhop_build against the independent torch-glue construction, incl. partial_2hop / adjacency-only / empty batches
    for kw in (dict(), dict(partial_2hop=True), dict(want_nhop=False)):
        a = tg.batch_edges(batch, **kw); b = tg.batch_edges_torch(batch, **kw)
        assert all(torch.equal(x.long(), y.long()) for x, y in zip(a, b)), kw
    e_idx, e_val, e_nhop = tg.batch_edges([])
    assert e_idx.shape == (2, 0) and e_val.numel() == 0 and e_nhop.shape == (0, 4)
    with pytest.raises(IndexError):
        tg.batch_edges([n])
    # and the model consumes them directly
    from recon_b200 import KGraph
    kg = KGraph(idx, val, nhop.long(), n, r, device=dev())
    assert kg.n_edges == len(want_val) + len(want_nhop)


# ---- argument checks the raw-pointer C ABI cannot do -------------------------------------------
def test_mismatched_tables_raise_instead_of_reading_out_of_bounds():
    from recon_b200 import SpKBGATModified, KGraph, SpGraphAttentionLayer
    from recon_b200.synth import make_kg
    n, r = 50, 5
    edge, etype, _ = make_kg(n, 300, r, seed=1)
    model = SpKBGATModified(torch.randn(n, 8), torch.randn(r, 8), [6, 12], [6, 12], 0.0, 0.2, [2, 2], None).to(dev())
    big = KGraph(edge, etype, None, n + 7, r, device=dev())                 # graph built for more nodes than the model has
    with pytest.raises(IndexError):
        model(None, torch.arange(n), big, None)
    more_rel = KGraph(edge, etype, None, n, r + 3, device=dev())            # ... or for more relations
    with pytest.raises(IndexError):
        model(None, torch.arange(n), more_rel, None)
    with pytest.raises(IndexError):                                         # batch_test with a shorter entity table
        model.batch_test(None, torch.arange(n), (edge, etype), None, torch.randn(n - 1, 8, device=dev()))
    layer = SpGraphAttentionLayer(n, 8, 6, 4, 0.0, 0.2).to(dev())
    with pytest.raises(RuntimeError):                                       # fewer edge embeddings than edges
        layer(torch.randn(n, 8, device=dev()), edge.to(dev()), torch.randn(299, 4, device=dev()), None, None)


def test_forward_rebinds_entity_data_like_the_reference():
    """models.py:160-161 assigns a NEW tensor to entity_embeddings.data: a tensor the caller shares with the Parameter must
    keep its values (the reference builds several models from one global embedding tensor)."""
    from recon_b200 import SpKBGATModified
    from recon_b200.synth import make_kg
    n, r = 40, 4
    edge, etype, _ = make_kg(n, 200, r, seed=2)
    init = (torch.randn(n, 8) * 3).to(dev())
    keep = init.clone()
    model = SpKBGATModified(init, torch.randn(r, 8).to(dev()), [6, 12], [6, 12], 0.0, 0.2, [2, 2], None).to(dev())
    model(None, torch.arange(n), (edge, etype), None)
    assert torch.equal(init, keep)
    norms = model.entity_embeddings.detach().norm(dim=1)
    assert torch.allclose(norms, torch.ones_like(norms), atol=1e-5)


def test_phased_split_backward_is_bit_identical_to_one_call():
    """bench.py's per-kernel timing pass issues the split-dot backward as four phased calls (node / columns / relations /
    column sums, include/spkbgat.h spk_edge_bwd_split_args.phases): same kernels, same order, same bits."""
    from recon_b200 import SpKBGATModified, profiler
    from recon_b200.synth import make_kg
    n, r = 3000, 17
    edge, etype, _ = make_kg(n, 40000, r, alpha=1.1, seed=5)
    torch.manual_seed(3)
    model = SpKBGATModified(torch.randn(n, 50), torch.randn(r, 50), [100, 200], [100, 200], 0.0, 0.2, [2, 2], None).to(dev())
    ent0 = model.entity_embeddings.detach().clone()
    graph = model.prepare_graph((edge.to(dev()), etype.to(dev())), None)

    def run():
        model.entity_embeddings.data = ent0.clone()
        model.zero_grad(set_to_none=True)
        out_e, out_r, _ = model(None, torch.arange(n), graph, None)
        (out_e.sum() + (out_r * out_r).sum()).backward()
        return {k: v.grad.clone() for k, v in model.named_parameters() if v.grad is not None}

    a = run()
    profiler.enable()
    b = run()
    prof = profiler.disable()
    assert {"edge_attn_bwd_split:node", "edge_attn_bwd_split:cols", "edge_attn_bwd_split:rels",
            "edge_attn_bwd_split:colsums"} <= set(prof)
    assert all(torch.equal(a[k], b[k]) for k in a)


@pytest.mark.parametrize("n_nhop", [0, 700])
def test_host_staged_edges_give_the_same_layouts(n_nhop):
    """Host-resident int64 tensors go through spk_pack_index_host + a copy stream (graph._stage_host_edges); device-resident
    ones through spk_edges_concat. Same layouts bit for bit, twice in a row (staging buffer reuse), same IndexError."""
    from recon_b200 import KGraph
    from recon_b200.synth import make_kg
    n, r = 900, 13
    edge, etype, nhop = make_kg(n, 300_000, r, alpha=1.1, n_nhop=n_nhop, seed=4)
    nh = nhop if n_nhop else None
    g_dev = KGraph(edge.to(dev()), etype.to(dev()), nh.to(dev()) if n_nhop else None, n, r, device=dev())
    for _ in range(2):
        g_host = KGraph(edge, etype, nh, n, r, device=dev())
        for name in ("row", "perm", "rowptr", "col", "t1", "t2", "colptr", "csc_row", "csc_pos", "csc_t1", "relptr", "rel_row", "rel_pos"):
            a, b = getattr(g_host, name), getattr(g_dev, name)
            assert (a is None and b is None) or torch.equal(a, b), name
    bad = edge.clone(); bad[1, 5] = n
    with pytest.raises(IndexError):
        KGraph(bad, etype, nh, n, r, device=dev())
    bad_t = etype.clone(); bad_t[7] = -1
    with pytest.raises(IndexError):
        KGraph(edge, bad_t, nh, n, r, device=dev())


@pytest.mark.parametrize("n", [0, 1, 7, 4096, 1_000_003, 40_000_000])
def test_inner_product_deterministic(n):
    """spk_inner_product: the linear probe loss <out, G>, fixed-order reduction."""
    from recon_b200 import functional as SF
    g = torch.Generator().manual_seed(n + 5)
    a = torch.randn(n, generator=g); b = torch.randn(n, generator=g)
    ad, bd = a.to(dev()), b.to(dev())
    v = SF.inner_product([ad], [bd])
    ref = float((a.double() * b.double()).sum())
    scale = float((a.double() * b.double()).abs().sum()) + 1e-30
    assert abs(float(v) - ref) / scale < 1e-6
    assert float(SF.inner_product([ad], [bd])) == float(v)          # bit-stable
    two = SF.inner_product([ad, ad], [bd, bd])
    assert abs(float(two) - 2 * ref) / scale < 2e-6


def test_linear_loss_backward_matches_autograd():
    from recon_b200 import functional as SF
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1000, 40, generator=g).to(dev()).requires_grad_(True)
    y = torch.randn(7, 40, generator=g).to(dev()).requires_grad_(True)
    ge, gr = torch.randn(1000, 40, generator=g).to(dev()), torch.randn(7, 40, generator=g).to(dev())
    loss = SF.linear_loss_backward((x * 2.0, y * 3.0), (ge, gr))
    gx, gy = x.grad.clone(), y.grad.clone()
    x.grad = y.grad = None
    ref = ((x * 2.0) * ge).sum() + ((y * 3.0) * gr).sum()
    ref.backward()
    assert torch.equal(gx, x.grad) and torch.equal(gy, y.grad)
    assert abs(float(loss) - float(ref)) < 1e-3 * abs(float(ref)) + 1e-3
    # differentiable form
    x.grad = None
    SF.inner_product([x * 2.0], [ge]).backward()
    assert torch.allclose(x.grad, 2.0 * ge)
