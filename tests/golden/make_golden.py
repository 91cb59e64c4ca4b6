"""Generate the golden fixtures in this directory by running the UNMODIFIED reference
(/root/reference/GAT/{layers,models,create_batch}.py, imported, never copied) on seeded inputs.

Run in the build container only (the GPU box has no /root/reference):
    python tests/golden/make_golden.py
The committed *.npz files hold inputs, parameters, dropout masks, outputs, gradients and the
reference's .data side effects; tests/test_oracle_golden.py pins oracle/ against them and the
-m gpu tests pin the CUDA path against them.
"""
import os
import sys
import types
import copy
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/GAT")

# create_batch.py imports nltk (absent here) only for an unrelated tokenizer.
_nltk = types.ModuleType("nltk"); _tok = types.ModuleType("nltk.tokenize")
_tok.word_tokenize = lambda s: s.split(); _nltk.tokenize = _tok
sys.modules.setdefault("nltk", _nltk); sys.modules.setdefault("nltk.tokenize", _tok)

import layers as ref_layers          # noqa: E402
import models as ref_models          # noqa: E402
import create_batch as ref_batch     # noqa: E402
from recon_b200.synth import make_kg, make_triples   # noqa: E402

ref_layers.CUDA = ref_models.CUDA = False


class FixedMask(torch.nn.Module):
    """Stands in for an nn.Dropout instance of the reference so that a known mask is applied."""

    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask.to(x.dtype)


def drop_mask(shape, p, gen):
    return (torch.rand(shape, generator=gen) >= p).float() / (1.0 - p)


def model_case(name, n, e, r, in_dim, out_dim, nheads, alpha_zipf, n_nhop, batch, p_drop, seed,
               batch_test=False):
    edge, etype, nhop = make_kg(n, e, r, alpha_zipf, n_nhop, seed)
    torch.manual_seed(seed)
    ent0 = torch.randn(n, in_dim); rel0 = torch.randn(r, in_dim)
    model = ref_models.SpKBGATModified(ent0.clone(), rel0.clone(), [out_dim, 2 * out_dim], [out_dim, 2 * out_dim],
                                       p_drop, 0.2, [nheads, nheads], None)
    gen = torch.Generator().manual_seed(seed + 1)
    etot = e + n_nhop
    masks = {}
    if p_drop > 0:
        masks["att"] = torch.stack([drop_mask((etot,), p_drop, gen) for _ in range(nheads)])
        masks["out"] = drop_mask((etot,), p_drop, gen)
        masks["x"] = drop_mask((n, out_dim * nheads), p_drop, gen)
        for i in range(nheads):
            getattr(model.sparse_gat_1, f"attention_{i}").dropout = FixedMask(masks["att"][i])
        model.sparse_gat_1.out_att.dropout = FixedMask(masks["out"])
        model.sparse_gat_1.dropout_layer = FixedMask(masks["x"])
    params0 = {k: v.detach().clone() for k, v in model.state_dict().items()}
    g_ent = torch.randn(n, out_dim * nheads, generator=gen)
    g_rel = torch.randn(r, out_dim * nheads, generator=gen)
    batch_entities = torch.as_tensor(batch, dtype=torch.int64)
    out = {"edge": edge, "edge_type": etype, "nhop": nhop, "batch_entities": batch_entities,
           "g_ent": g_ent, "g_rel": g_rel, "alpha": np.float32(0.2), "p_drop": np.float32(p_drop)}
    for k, v in masks.items():
        out["mask." + k] = v
    for k, v in params0.items():
        out["param." + k] = v

    def run(m):
        if batch_test:
            ent_in = torch.randn(n, in_dim, generator=torch.Generator().manual_seed(seed + 2)).to(
                next(m.parameters()).dtype)
            o_e, o_r, msk = m.batch_test(None, batch_entities, (edge, etype), nhop, ent_in)
            return o_e, o_r, msk, ent_in
        o_e, o_r, msk = m(None, batch_entities, (edge, etype), nhop)
        return o_e, o_r, msk, None

    m64 = copy.deepcopy(model).double()
    for tag, m, ge, gr in (("f32", model, g_ent, g_rel), ("f64", m64, g_ent.double(), g_rel.double())):
        o_e, o_r, msk, ent_in = run(m)
        loss = (o_e * ge).sum() + (o_r * gr).sum()
        loss.backward()
        out[f"{tag}.out_entity"] = o_e.detach(); out[f"{tag}.out_relation"] = o_r.detach()
        out[f"{tag}.mask"] = msk
        if ent_in is not None and tag == "f32":
            out["entity_in"] = ent_in
        for k, prm in m.named_parameters():
            if prm.grad is not None:
                out[f"{tag}.grad.{k}"] = prm.grad.detach()
        for k, v in m.state_dict().items():          # side effects (models.py:160-161,181-183)
            if k in ("entity_embeddings", "final_entity_embeddings", "final_relation_embeddings"):
                out[f"{tag}.after.{k}"] = v.detach().clone()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in out.items()})
    print(name, "ok", {k: tuple(np.asarray(v).shape) for k, v in out.items() if k.startswith("f32.out")})


def layer_case(name, n, e, e2, f_in, d, rd, concat, seed):
    """SpGraphAttentionLayer stand-alone, per-edge embeddings given (layers.py:111)."""
    gen = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    layer = ref_layers.SpGraphAttentionLayer(n, f_in, d, rd, 0.0, 0.2, concat)
    x = torch.randn(n, f_in, generator=gen, requires_grad=True)
    edge = torch.randint(0, n, (2, e), generator=gen)
    edge[0, : e // 4] = 3                                   # one heavier row, a few empty rows remain
    emb = torch.randn(e, rd, generator=gen, requires_grad=True)
    edge2 = torch.randint(0, n, (2, e2), generator=gen) if e2 else torch.tensor([])
    emb2 = torch.randn(e2, rd, generator=gen, requires_grad=True) if e2 else torch.tensor([])
    g = torch.randn(n, d, generator=gen)
    out = layer(x, edge, emb, edge2, emb2)
    (out * g).sum().backward()
    res = {"x": x.detach(), "edge": edge, "edge_embed": emb.detach(), "edge_nhop": edge2,
           "edge_embed_nhop": emb2.detach(), "g": g, "a": layer.a.detach(), "a_2": layer.a_2.detach(),
           "concat": np.int32(concat), "out": out.detach(), "grad.x": x.grad, "grad.edge_embed": emb.grad,
           "grad.a": layer.a.grad, "grad.a_2": layer.a_2.grad}
    if e2:
        res["grad.edge_embed_nhop"] = emb2.grad
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in res.items()})
    print(name, "ok")


def spmm_case(name, n, e, f, seed):
    """SpecialSpmmFunctionFinal.apply stand-alone (layers.py:51-79)."""
    gen = torch.Generator().manual_seed(seed)
    edge = torch.randint(0, n, (2, e), generator=gen)
    w = torch.randn(e, f, generator=gen, requires_grad=True)
    g = torch.randn(n, f, generator=gen)
    out = ref_layers.SpecialSpmmFunctionFinal.apply(edge, w, n, e, f)
    (out * g).sum().backward()
    np.savez_compressed(os.path.join(HERE, name + ".npz"), edge=edge.numpy(), w=w.detach().numpy(), g=g.numpy(),
                        out=out.detach().numpy(), grad_w=w.grad.numpy())
    print(name, "ok")


def edges_case(name, n, t, r, seed, batch_size, partial=False):
    """Corpus.get_graph / bfs / get_batch_adj_data / get_batch_nhop_neighbors_all on random triples."""
    triples = make_triples(n, t, r, seed)
    rows = triples[:, 2].tolist(); cols = triples[:, 0].tolist(); data = triples[:, 1].tolist()
    c = ref_batch.Corpus.__new__(ref_batch.Corpus)
    c.train_adj_matrix = (torch.LongTensor([rows, cols]), torch.LongTensor(data))     # create_batch.py:28-31
    c.graph = c.get_graph(Train=True)
    c.node_neighbors_1hop = c.get_further_neighbors(nbd_size=1, Train=True)
    c.node_neighbors_2hop = c.get_further_neighbors(nbd_size=2, Train=True)
    gen = torch.Generator().manual_seed(seed + 7)
    batch = torch.randperm(n, generator=gen)[:batch_size].tolist()
    (adj_idx, adj_val), ents = c.get_batch_adj_data(None, unique_entities_train=batch, start_idx=0, end_idx=len(batch))
    args = types.SimpleNamespace(partial_2hop=partial)
    nhop = c.get_batch_nhop_neighbors_all(args, batch, c.node_neighbors_2hop)
    all_src = list(range(n))
    (fadj_idx, fadj_val), _ = c.get_batch_adj_data(None, unique_entities_train=all_src, start_idx=0, end_idx=n)
    fnhop = c.get_batch_nhop_neighbors_all(args, all_src, c.node_neighbors_2hop)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), triples=triples.numpy(), batch=np.asarray(batch),
                        adj_idx=adj_idx.numpy(), adj_val=adj_val.numpy(), nhop=nhop.reshape(-1, 4),
                        full_adj_idx=fadj_idx.numpy(), full_adj_val=fadj_val.numpy(), full_nhop=fnhop.reshape(-1, 4),
                        partial=np.int32(partial))
    print(name, "ok", adj_idx.shape, nhop.shape, fnhop.shape)


def _reference_function(path, name, env):
    """Compile ONE top-level function of a reference file from its own source text (read here, never copied into the
    repo) into `env`. GAT/main.py cannot be imported: it parses sys.argv and imports matplotlib at module level."""
    import ast
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
    exec(code, env)
    return env[name]


def load_data_case(name, n, t, r, seed):
    """preprocess.load_data (GAT/preprocess.py:48-87) on a temporary triple file, all four flag combinations."""
    import tempfile
    import preprocess as ref_pre
    triples = make_triples(n, t, r, seed)
    d = tempfile.mkdtemp()
    fn = os.path.join(d, "train.txt")
    e2i = {f"e{i}": i for i in range(n)}; r2i = {f"r{i}": i for i in range(r)}
    with open(fn, "w") as f:
        f.write("\n".join(f"e{h} r{k} e{tl}" for h, k, tl in triples.tolist()) + "\n\n")
    out = {"triples": triples.numpy()}
    for directed in (True, False):
        for unw in (False, True):
            td, (rows, cols, data), _ = ref_pre.load_data(fn, e2i, r2i, unw, directed)
            assert td == [tuple(x) for x in triples.tolist()]
            out[f"adj.directed{int(directed)}.unweighted{int(unw)}"] = np.asarray([rows, cols, data], dtype=np.int64)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print(name, "ok")


def sampler_case(name, n, t, r, seed, batch_size, ratio, np_seed):
    """Corpus.get_iteration_triples_batch (create_batch.py:262-351), the reference's own method, np.random seeded."""
    triples = make_triples(n, t, r, seed)
    rows = triples[:, 2].tolist(); cols = triples[:, 0].tolist(); data = triples[:, 1].tolist()
    c = ref_batch.Corpus.__new__(ref_batch.Corpus)
    c.train_adj_matrix = (torch.LongTensor([rows, cols]), torch.LongTensor(data))     # create_batch.py:28-31
    c.graph = c.get_graph(Train=True)
    c.node_neighbors_1hop = c.get_further_neighbors(nbd_size=1, Train=True)
    c.entity2id = {str(i): i for i in range(n)}
    c.relation2id = {str(i): i for i in range(r)}
    c.invalid_valid_ratio = ratio
    train_triples = [tuple(int(v) for v in row) for row in triples.tolist()]
    c.valid_triples_dict = {j: i for i, j in enumerate(train_triples)}                # create_batch.py:82-83
    gen = torch.Generator().manual_seed(seed + 7)
    batch = torch.randperm(n, generator=gen)[:batch_size].tolist()
    np.random.seed(np_seed)
    idx, val = c.get_iteration_triples_batch(batch)
    p = idx.shape[0] // (2 * ratio + 1)
    np.random.seed(np_seed)                                  # replay the seed: the first draws the method made
    init_e = np.random.randint(0, n, p * ratio); init_r = np.random.randint(0, r, p * ratio)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), triples=triples.numpy(), batch=np.asarray(batch),
                        ratio=np.int32(ratio), np_seed=np.int64(np_seed), batch_indices=idx.copy(), batch_values=val.copy(),
                        init_entities=init_e, init_relations=init_r, n_entities=np.int64(n), n_relations=np.int64(r))
    print(name, "ok", idx.shape, "positives", p)


def _reference_class(path, name, env):
    import ast
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == name)
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), env)
    return env[name]


def export_table(seed, rows, width):
    """Rows of floats that exercise every branch of Python's float repr: hot-path-like unit rows, exponent-notation
    boundaries (1e-4 / 1e-5, 1e16), integral values, denormals, +-0, NaN, +-Infinity."""
    g = torch.Generator().manual_seed(seed)
    t = torch.nn.functional.normalize(torch.randn(rows, width, generator=g), dim=1)
    special = torch.tensor([0.0, -0.0, 1.0, -1.0, 100.0, 1e-4, 1e-5, 9.999e-5, 1.5e-5, 1e16, 1.5e16, 9.99e15, 1e15,
                            123456.0, 16777216.0, 3.4028235e38, -3.4028235e38, 1.17549435e-38, 1e-45, 7e-42, 0.1, 0.5,
                            2.5e-7, float("nan"), float("inf"), float("-inf"), 1e22, 1e-10, 4.0e9, 65504.0])
    k = min(special.numel(), t.numel() - 1)
    t.view(-1)[1:1 + k] = special[:k]
    return t


def export_case(name, seed, rows, width):
    """save_embed of GAT/main.py:406-413 with its CustomEncoder (main.py:127-144), both executed from the reference's text."""
    import json
    import datetime
    env = {"json": json, "np": np, "datetime": datetime}
    _reference_class("/root/reference/GAT/main.py", "CustomEncoder", env)
    ref_save = _reference_function("/root/reference/GAT/main.py", "save_embed", env)
    t = export_table(seed, rows, width)
    np.save(os.path.join(HERE, name + ".npy"), t.numpy())
    ref_save(t, os.path.join(HERE, name + ".json"))
    print(name, "ok", os.path.getsize(os.path.join(HERE, name + ".json")), "bytes")


def loss_case(name, n_ent, n_rel, width, n_pos, ratio, margin, seed, hub_share=0.0):
    """batch_gat_loss of GAT/main.py:344-376 (unmodified function text) + loss.backward() + one SGD step."""
    from oracle.loss import make_train_indices
    env = {"torch": torch, "CUDA": False, "int": int,
           "args": types.SimpleNamespace(valid_invalid_ratio_gat=ratio)}
    ref_loss = _reference_function("/root/reference/GAT/main.py", "batch_gat_loss", env)
    gen = torch.Generator().manual_seed(seed)
    tri = make_train_indices(n_ent, n_rel, n_pos, ratio, seed + 1, hub_entity=3, hub_share=hub_share)
    out = {"train_indices": tri, "ratio": np.int32(ratio), "margin": np.float32(margin)}
    ent0 = torch.randn(n_ent, width, generator=gen); rel0 = torch.randn(n_rel, width, generator=gen)
    ent0 = torch.nn.functional.normalize(ent0, p=2, dim=1)        # the hot path's outputs are unit rows (models.py:179)
    tri[0, 0], tri[0, 2] = 6, 5                                   # x of triple 0 has exact zeros: sign(0) = 0 in the backward
    hw, r0 = width // 2, int(tri[0, 1])                           # (dyadic values: the sum is exact in fp32 AND fp64)
    ent0[6, :hw] = torch.round(ent0[6, :hw] * 1024) / 1024
    rel0[r0, :hw] = torch.round(rel0[r0, :hw] * 1024) / 1024
    ent0[5, :hw] = ent0[6, :hw] + rel0[r0, :hw]
    out["entity_embed"], out["relation_embed"] = ent0, rel0
    lr = 1e-3
    for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
        ent = ent0.to(dt).clone().requires_grad_(True); rel = rel0.to(dt).clone().requires_grad_(True)
        loss = ref_loss(torch.nn.MarginRankingLoss(margin=margin), tri, ent, rel)
        loss.backward()
        out[f"{tag}.loss"] = loss.detach(); out[f"{tag}.grad.entity_embed"] = ent.grad; out[f"{tag}.grad.relation_embed"] = rel.grad
        opt = torch.optim.SGD([ent, rel], lr=lr)                   # main.py:445-446
        opt.step()
        out[f"{tag}.sgd.entity_embed"] = ent.detach().clone(); out[f"{tag}.sgd.relation_embed"] = rel.detach().clone()
    out["lr"] = np.float64(lr)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v) for k, v in out.items()})
    print(name, "ok", float(out["f32.loss"]), tuple(tri.shape))


def _load_sep_space():
    """GAT_sep_space/{layers,models}.py under private module names (they shadow the GAT/ module names)."""
    import importlib.util
    mods = {}
    for name in ("layers", "models"):
        spec = importlib.util.spec_from_file_location("_sep_" + name, f"/root/reference/GAT_sep_space/{name}.py")
        mod = importlib.util.module_from_spec(spec)
        saved = sys.modules.get("layers")
        if name == "models":
            sys.modules["layers"] = mods["layers"]
        try:
            spec.loader.exec_module(mod)
        finally:
            if saved is not None:
                sys.modules["layers"] = saved
        mod.CUDA = False
        mods[name] = mod
    return mods["layers"], mods["models"]


def convkb_case(name, n, r, d_half, nheads, batch, n_test, seed, sep_space=False):
    """SpKBGATConvOnly of the reference (GAT/models.py:242-304, or the GAT_sep_space variant with W_ent2rel,
    GAT_sep_space/models.py:311-339): forward on a batch of triples, SoftMarginLoss backward (GAT/main.py:753,818-835),
    batch_test scores of every test triple under every relation (create_batch.py:1367-1393)."""
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    d = d_half * nheads
    ent0 = torch.randn(n, 10); rel0 = torch.randn(r, 10)
    mods = _load_sep_space()[1] if sep_space else ref_models
    conv = mods.SpKBGATConvOnly(ent0, rel0, [d_half, 2 * d_half], [d_half, 2 * d_half], 0.0, 0.0, 0.2, 0.2, [nheads, nheads], 50)
    conv.final_entity_embeddings.data = torch.nn.functional.normalize(torch.randn(n, d, generator=gen), dim=1)
    conv.final_relation_embeddings.data = torch.randn(r, d, generator=gen)
    gat = None
    if sep_space:
        gat = types.SimpleNamespace(W_ent2rel=torch.nn.Parameter(torch.randn(r, d, d, generator=gen) * 0.1),
                                    nonlinearity_ent2rel=torch.tanh)
    tri = torch.stack((torch.randint(0, n, (batch,), generator=gen), torch.randint(0, r, (batch,), generator=gen),
                       torch.randint(0, n, (batch,), generator=gen)), dim=1)
    target = (torch.rand(batch, generator=gen) < 0.3).float() * 2 - 1
    out = {"triples": tri, "target": target}
    for k, v in conv.state_dict().items():
        out["param." + k] = v.detach().clone()
    if sep_space:
        out["W_ent2rel"] = gat.W_ent2rel.detach().clone()
    for tag, m in (("f32", conv), ("f64", copy.deepcopy(conv).double())):
        g = gat
        if sep_space and tag == "f64":
            g = types.SimpleNamespace(W_ent2rel=torch.nn.Parameter(gat.W_ent2rel.detach().double()), nonlinearity_ent2rel=torch.tanh)
        m.zero_grad()
        preds = m(None, None, tri, g) if sep_space else m(None, None, tri)
        loss = torch.nn.SoftMarginLoss()(preds.view(-1), target.to(preds.dtype))
        loss.backward()
        out[tag + ".preds"] = preds.detach()
        out[tag + ".loss"] = loss.detach()
        for k, prm in m.named_parameters():
            if prm.grad is not None:
                out[tag + ".grad." + k] = prm.grad.detach().clone()
        if sep_space:
            out[tag + ".grad.W_ent2rel"] = g.W_ent2rel.grad.detach().clone()
        test = torch.stack((torch.randint(0, n, (n_test,), generator=torch.Generator().manual_seed(seed + 2)),
                            torch.randint(0, r, (n_test,), generator=torch.Generator().manual_seed(seed + 3)),
                            torch.randint(0, n, (n_test,), generator=torch.Generator().manual_seed(seed + 4))), dim=1)
        test[1] = test[0]; test[1, 1] = (test[0, 1] + 1) % r            # an entity pair with two actual relations
        tb = test.unsqueeze(1).repeat(1, r, 1)
        tb[:, :, 1] = torch.arange(r).unsqueeze(0)
        with torch.no_grad():
            sc = m.batch_test(tb.reshape(-1, 3), g) if sep_space else m.batch_test(tb.reshape(-1, 3))
        out["test_triples"] = test
        out[tag + ".scores"] = sc.view(n_test, r)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **{k: np.asarray(v.detach() if isinstance(v, torch.Tensor) else v) for k, v in out.items()})
    print(name, "ok", float(out["f32.loss"]), tuple(out["f32.scores"].shape))


if __name__ == "__main__":
    import warnings
    warnings.filterwarnings("ignore")
    if len(sys.argv) > 1 and sys.argv[1] == "convkb":
        convkb_case("convkb_gat", 300, 11, 100, 2, 257, 40, 80)
        convkb_case("convkb_small", 40, 5, 6, 2, 33, 12, 81)
        convkb_case("convkb_sep", 120, 7, 20, 2, 129, 25, 82, sep_space=True)
        sys.exit(0)
    model_case("model_small_uniform", 60, 300, 7, 12, 8, 2, None, 0, [3, 5, 5, 17, 40, 41], 0.0, 0)
    model_case("model_small_zipf_nhop", 80, 400, 9, 12, 8, 2, 1.1, 150, list(range(80)), 0.0, 1)
    model_case("model_small_dropmask", 70, 350, 5, 16, 12, 2, 1.5, 120, list(range(0, 70, 2)), 0.3, 2)
    model_case("model_refdims", 150, 1200, 11, 50, 100, 2, 1.1, 300, list(range(150)), 0.0, 3)
    model_case("model_batch_test", 60, 300, 7, 12, 8, 2, None, 80, list(range(30)), 0.0, 4, batch_test=True)
    model_case("model_3heads", 50, 260, 6, 10, 4, 3, 2.0, 0, list(range(50)), 0.0, 5)
    layer_case("layer_concat", 40, 200, 0, 10, 8, 6, True, 10)
    layer_case("layer_noconcat_nhop", 40, 200, 60, 12, 16, 12, False, 11)
    spmm_case("spmm_f1", 30, 200, 1, 20)
    spmm_case("spmm_f7", 30, 200, 7, 21)
    edges_case("edges_a", 40, 160, 5, 30, 12)
    edges_case("edges_b", 25, 200, 3, 31, 25)
    edges_case("edges_partial", 40, 160, 5, 32, 15, partial=True)
    loss_case("loss_small", 40, 5, 12, 30, 2, 0.5, 40)
    loss_case("loss_refdims_hub", 300, 11, 200, 700, 2, 5.0, 41, hub_share=0.9)     # entity 3 heads ~630 positives: hub segment
    loss_case("loss_ratio3_oddwidth", 50, 4, 7, 25, 3, 1.0, 42)
    load_data_case("load_data_small", 30, 100, 5, 70)
    sampler_case("sampler_a", 60, 400, 6, 60, 20, 2, 123)
    sampler_case("sampler_dense_r1", 12, 120, 1, 61, 12, 2, 124)      # one relation: relation corruption exhausts (348-350)
    sampler_case("sampler_ratio3", 40, 200, 4, 62, 15, 3, 125)        # odd ratio: untouched +1 copies in the middle
    sampler_case("sampler_ratio1", 40, 200, 4, 63, 15, 1, 126)        # ratio // 2 == 0: relation corruption only
    export_case("export_small", 50, 9, 7)
    export_case("export_w200", 51, 3, 200)
    # the hand-checked toy KG of SURVEY.md 3.4
    toy = np.array([(0, 5, 1), (0, 6, 1), (0, 7, 2), (1, 8, 3), (2, 9, 3), (2, 4, 2), (3, 1, 4), (1, 2, 0)])
    np.savez_compressed(os.path.join(HERE, "edges_toy.npz"), triples=toy)
