"""CPU oracle for the SpKBGAT hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

A functional, op-for-op restatement (PyTorch CPU, fp32 or fp64) of the
reference's KBGAT sparse triple-attention path:

  * seg_sum_coo            <- SpecialSpmmFunctionFinal     /root/reference/GAT/layers.py:51-79
  * attention_layer        <- SpGraphAttentionLayer.forward /root/reference/GAT/layers.py:111-178
  * sp_gat                 <- SpGAT.forward                 /root/reference/GAT/models.py:47-88
  * kbgat_forward          <- SpKBGATModified.forward       /root/reference/GAT/models.py:136-185
  * kbgat_batch_test       <- SpKBGATModified.batch_test    /root/reference/GAT/models.py:188-239

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (recon_b200/) never does.

Pinning: the reference ships no golden vectors, so this
oracle is pinned against outputs of the reference itself, run in the build
container by tests/golden/make_golden.py and committed as tests/golden/*.npz;
tests/test_oracle_golden.py re-checks it on every run.

Dropout is expressed with explicit multiplicative masks (values 0 or 1/(1-p))
so that a supplied mask reproduces the reference's nn.Dropout sites exactly:
  masks["att"][h] : [E]    layers.py:158 (head h of layer 1)
  masks["out"]    : [E]    layers.py:158 (out_att)
  masks["x"]      : [N,H*D] models.py:73  (dropout_layer)
"""
import torch
import torch.nn.functional as F


class _SegSumCOO(torch.autograd.Function):
    """out[i,:] = sum_{e: edge[0,e]==i} edge_w[e,:]  via hybrid COO + sparse.sum, as
    layers.py:56-64 does it; backward gathers grad rows by edge[0] (layers.py:67-79)."""

    @staticmethod
    def forward(ctx, edge, edge_w, n_rows):
        coo = torch.sparse_coo_tensor(edge, edge_w, (n_rows, n_rows, edge_w.shape[1]),
                                      check_invariants=False)
        ctx.save_for_backward(edge[0])
        return torch.sparse.sum(coo, dim=1).to_dense()

    @staticmethod
    def backward(ctx, g):
        (rows,) = ctx.saved_tensors
        return None, g[rows], None


class _NormalizeData(torch.autograd.Function):
    """models.py:160-161 overwrites entity_embeddings.data with its L2-normalised rows and
    then uses the Parameter itself: the value is normalised, the gradient reaches the
    parameter unchanged (no normalise backward)."""

    @staticmethod
    def forward(ctx, x):
        return F.normalize(x, p=2, dim=1)

    @staticmethod
    def backward(ctx, g):
        return g


def seg_sum_coo(edge, edge_w, n_rows):
    return _SegSumCOO.apply(edge, edge_w, n_rows)


def seg_sum_index_add(edge, edge_w, n_rows):
    """Same sum with index_add_ (natively differentiable); used for fp64 runs at sizes
    where the COO path is too slow. Differs from seg_sum_coo only in fp add order."""
    out = torch.zeros(n_rows, edge_w.shape[1], dtype=edge_w.dtype)
    return out.index_add(0, edge[0], edge_w)


def attention_layer(x, edge, edge_embed, edge_nhop, edge_embed_nhop, a, a_2, alpha,
                    concat, drop_mask=None, seg_sum=seg_sum_coo):
    """SpGraphAttentionLayer.forward, layers.py:111-178."""
    n = x.shape[0]
    if edge_nhop is not None and edge_nhop.shape[0] > 0:          # layers.py:124-127
        edge = torch.cat((edge, edge_nhop), dim=1)
        edge_embed = torch.cat((edge_embed, edge_embed_nhop), dim=0)
    edge_h = torch.cat((x[edge[0]], x[edge[1]], edge_embed), dim=1).t()   # 129-130
    edge_m = a.mm(edge_h)                                                  # 137
    powers = -F.leaky_relu(a_2.mm(edge_m).squeeze(0), alpha)               # 143
    edge_e = torch.exp(powers).unsqueeze(1)                                # 146
    assert not torch.isnan(edge_e).any()                                   # 147
    e_rowsum = seg_sum(edge, edge_e, n)                                    # 150-151
    e_rowsum = torch.where(e_rowsum == 0.0, torch.full_like(e_rowsum, 1e-12), e_rowsum)  # 152
    edge_e = edge_e.squeeze(1)
    if drop_mask is not None:                                              # 158
        edge_e = edge_e * drop_mask
    edge_w = (edge_e * edge_m).t()                                         # 161
    h_prime = seg_sum(edge, edge_w, n)                                     # 164-165
    assert not torch.isnan(h_prime).any()                                  # 167
    h_prime = h_prime.div(e_rowsum)                                        # 169
    assert not torch.isnan(h_prime).any()                                  # 172
    return F.elu(h_prime) if concat else h_prime                           # 173-178


def sp_gat(x, rel, edge, edge_type, edge_nhop, edge_type_nhop, heads, W, out_a, out_a2,
           alpha, masks=None, seg_sum=seg_sum_coo):
    """SpGAT.forward, models.py:47-88. heads = [(a, a_2), ...]; returns (x_out, out_relation_1)."""
    masks = masks or {}
    has_nhop = edge_type_nhop is not None and edge_type_nhop.shape[0] > 0
    edge_embed = rel[edge_type]                                            # models.py:156
    edge_embed_nhop = (rel[edge_type_nhop[:, 0]] + rel[edge_type_nhop[:, 1]]) if has_nhop else None  # 56-62
    att = masks.get("att")
    hs = [attention_layer(x, edge, edge_embed, edge_nhop if has_nhop else None, edge_embed_nhop,
                          a, a2, alpha, True, None if att is None else att[i], seg_sum)
          for i, (a, a2) in enumerate(heads)]
    x1 = torch.cat(hs, dim=1)                                              # 71-72
    if masks.get("x") is not None:                                         # 73
        x1 = x1 * masks["x"]
    assert not torch.isnan(W).any() and not torch.isnan(rel).any()         # 75-76
    out_rel = rel.mm(W)                                                    # 77
    edge_embed = out_rel[edge_type]                                        # 79
    edge_embed_nhop = (out_rel[edge_type_nhop[:, 0]] + out_rel[edge_type_nhop[:, 1]]) if has_nhop else None  # 80-84
    x2 = F.elu(attention_layer(x1, edge, edge_embed, edge_nhop if has_nhop else None, edge_embed_nhop,
                               out_a, out_a2, alpha, False, masks.get("out"), seg_sum))  # 86
    return x2, out_rel


def split_nhop(train_indices_nhop):
    """models.py:141-148: rows [s, r1, r2, t] -> edge_list_nhop=[t; s], edge_type_nhop=[r1, r2]."""
    if train_indices_nhop is None or train_indices_nhop.shape[0] == 0:
        return None, None
    nh = train_indices_nhop.long()
    return torch.stack((nh[:, 3], nh[:, 0]), dim=0), nh[:, 1:3].contiguous()


def kbgat_forward(p, batch_entities, adj, train_indices_nhop, alpha, masks=None,
                  seg_sum=seg_sum_coo, entity_override=None, detach_rel=False):
    """SpKBGATModified.forward (models.py:136-185); with entity_override/detach_rel it is
    batch_test (188-239). p maps the reference state_dict names to tensors. Returns
    (out_entity, out_relation, mask, normalised_entity_input)."""
    edge, edge_type = adj
    edge_nhop, edge_type_nhop = split_nhop(train_indices_nhop)
    ent_src = p["entity_embeddings"] if entity_override is None else entity_override
    ent = _NormalizeData.apply(ent_src)                                     # 160-161
    rel = p["relation_embeddings"].detach() if detach_rel else p["relation_embeddings"]
    nheads = sum(1 for k in p if k.startswith("sparse_gat_1.attention_") and k.endswith(".a"))
    heads = [(p[f"sparse_gat_1.attention_{i}.a"], p[f"sparse_gat_1.attention_{i}.a_2"]) for i in range(nheads)]
    x2, out_rel = sp_gat(ent, rel, edge, edge_type, edge_nhop, edge_type_nhop, heads,
                         p["sparse_gat_1.W"], p["sparse_gat_1.out_att.a"], p["sparse_gat_1.out_att.a_2"],
                         alpha, masks, seg_sum)
    mask = torch.zeros(ent.shape[0], dtype=ent.dtype)                       # 167-173
    mask[torch.unique(batch_entities)] = 1.0
    out = ent.mm(p["W_entities"]) + mask.unsqueeze(-1) * x2                 # 175-177
    out = F.normalize(out, p=2, dim=1)                                      # 179
    return out, out_rel, mask, ent.detach()


def init_params(n_ent, n_rel, in_dim, out_dim, nheads, seed=0, dtype=torch.float32):
    """Parameters with the reference's names, shapes and initialisers
    (layers.py:100-105, models.py:37-38,119-134; main.py:259-262 for the embeddings)."""
    g = torch.Generator().manual_seed(seed)

    def xavier_normal(shape, gain=1.414):
        fan_out, fan_in = shape
        return torch.randn(shape, generator=g) * (gain * (2.0 / (fan_in + fan_out)) ** 0.5)

    def xavier_uniform(shape, gain=1.414):
        fan_out, fan_in = shape
        b = gain * (6.0 / (fan_in + fan_out)) ** 0.5
        return (torch.rand(shape, generator=g) * 2 - 1) * b

    hd = out_dim * nheads
    p = {
        "final_entity_embeddings": torch.randn(n_ent, hd, generator=g),
        "final_relation_embeddings": torch.randn(n_rel, hd, generator=g),
        "entity_embeddings": torch.randn(n_ent, in_dim, generator=g),
        "relation_embeddings": torch.randn(n_rel, in_dim, generator=g),
        "W_entities": xavier_uniform((in_dim, hd)),
        "sparse_gat_1.W": xavier_uniform((in_dim, hd)),
    }
    for i in range(nheads):
        p[f"sparse_gat_1.attention_{i}.a"] = xavier_normal((out_dim, 3 * in_dim))
        p[f"sparse_gat_1.attention_{i}.a_2"] = xavier_normal((1, out_dim))
    p["sparse_gat_1.out_att.a"] = xavier_normal((hd, 3 * hd))
    p["sparse_gat_1.out_att.a_2"] = xavier_normal((1, hd))
    return {k: v.to(dtype) for k, v in p.items()}


TRAINABLE = ("entity_embeddings", "relation_embeddings", "W_entities", "sparse_gat_1.W",
             "sparse_gat_1.out_att.a", "sparse_gat_1.out_att.a_2")


def trainable_names(p):
    return [k for k in p if k in TRAINABLE or k.startswith("sparse_gat_1.attention_")]


def fwd_bwd(p, batch_entities, adj, nhop, alpha, g_ent, g_rel, masks=None, seg_sum=seg_sum_coo):
    """One forward + backward with loss = <out_entity, g_ent> + <out_relation, g_rel>
    (SURVEY.md section 8d). Returns (out_entity, out_relation, mask, ent_norm, grads{name: tensor})."""
    names = trainable_names(p)
    q = {k: (v.detach().clone().requires_grad_(True) if k in names else v) for k, v in p.items()}
    out, out_rel, mask, ent = kbgat_forward(q, batch_entities, adj, nhop, alpha, masks, seg_sum)
    loss = (out * g_ent).sum() + (out_rel * g_rel).sum()
    loss.backward()
    return out.detach(), out_rel.detach(), mask, ent, {k: q[k].grad for k in names}
