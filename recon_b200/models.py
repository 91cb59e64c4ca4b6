"""Drop-in for the KBGAT modules of the reference's GAT/models.py on libspkbgat's sm_100a kernels.

  SpGAT            <- GAT/models.py:11-88    (same ctor, parameter names: attention_{i}.a/.a_2, W, out_att.a/.a_2)
  SpKBGATModified  <- GAT/models.py:91-239   (same ctor, forward / batch_test signatures, return tuple,
                                              state_dict keys and the three .data side effects)
Extensions (keyword-only, default None): `dropout_masks` = {"att": [H,E], "out": [E], "x": [N,H*D]} multipliers
in the caller's edge order for bit-reproducible dropout; `adj` may also be a prebuilt recon_b200.KGraph.
"""
import torch
import torch.nn as nn

from . import functional as SF
from .graph import KGraph
from .layers import SpGraphAttentionLayer, ConvKB, check_nanflag, edge_dropout_mask  # noqa: F401


def _fingerprint(t):
    """A few elements of a host index tensor (no device sync for CUDA tensors: identity + version only there)."""
    if t.is_cuda or t.numel() == 0:
        return None
    flat = t.reshape(-1)
    n = flat.numel()
    return tuple(int(flat[i]) for i in sorted({0, n // 3, n // 2, (2 * n) // 3, n - 1}))


def _nhop_rows(edge_list_nhop, edge_type_nhop):
    """([t; s], [r1, r2]) (models.py:145-148) -> rows [s, r1, r2, t]."""
    if edge_type_nhop is None or edge_type_nhop.numel() == 0:
        return None
    el = edge_list_nhop.long()
    et = edge_type_nhop.long().to(el.device)
    return torch.stack((el[1], et[:, 0], et[:, 1], el[0]), dim=1)


class SpGAT(nn.Module):
    def __init__(self, num_nodes, nfeat, nhid, relation_dim, dropout, alpha, nheads):
        super().__init__()
        self.dropout = dropout
        self.alpha = alpha
        self.dropout_layer = nn.Dropout(self.dropout)
        self.attentions = [SpGraphAttentionLayer(num_nodes, nfeat, nhid, relation_dim, dropout=dropout,
                                                 alpha=alpha, concat=True) for _ in range(nheads)]
        for i, attention in enumerate(self.attentions):
            self.add_module('attention_{}'.format(i), attention)
        self.W = nn.Parameter(torch.zeros(size=(relation_dim, nheads * nhid)))
        nn.init.xavier_uniform_(self.W.data, gain=1.414)
        self.out_att = SpGraphAttentionLayer(num_nodes, nhid * nheads, nheads * nhid, nheads * nhid,
                                             dropout=dropout, alpha=alpha, concat=False)

    def forward(self, Corpus_, entity_embeddings, relation_embed, edge_list, edge_type, edge_embed,
                edge_list_nhop, edge_type_nhop, graph=None, dropout_masks=None, nanflag=None):
        """models.py:47-88. `edge_embed` is accepted for signature compatibility and ignored: relation rows are
        gathered inside the fused kernel from `relation_embed` by `edge_type`."""
        x = entity_embeddings
        if not x.is_cuda:
            raise RuntimeError("recon_b200.SpGAT needs CUDA tensors (no CPU fallback)")
        dev = x.device
        with torch.cuda.device(dev):          # launch on the tensors' device, whatever the caller's current device is
            return self._forward(x, relation_embed, edge_list, edge_type, edge_list_nhop, edge_type_nhop, graph,
                                 dropout_masks, nanflag)

    def _forward(self, x, relation_embed, edge_list, edge_type, edge_list_nhop, edge_type_nhop, graph, dropout_masks, nanflag):
        dev = x.device
        if graph is None:
            graph = KGraph(edge_list, edge_type, _nhop_rows(edge_list_nhop, edge_type_nhop), x.shape[0],
                           relation_embed.shape[0], device=dev)
        own_flag = nanflag is None
        if own_flag:
            nanflag = torch.zeros(1, dtype=torch.int32, device=dev)
        nheads = len(self.attentions)
        e = graph.n_edges
        p = self.dropout
        m_att = m_out = m_x = None
        if dropout_masks is not None:
            if dropout_masks.get("att") is not None:
                m_att = graph.to_csr_order(dropout_masks["att"].to(dev, torch.float32).reshape(nheads, e))
            if dropout_masks.get("out") is not None:
                m_out = graph.to_csr_order(dropout_masks["out"].to(dev, torch.float32).reshape(1, e))
            if dropout_masks.get("x") is not None:
                m_x = dropout_masks["x"].to(dev, torch.float32)
        elif self.training and p > 0:
            m_att = edge_dropout_mask(p, nheads, e, dev)
            m_out = edge_dropout_mask(p, 1, e, dev)
            m_x = (torch.rand(x.shape[0], nheads * self.attentions[0].out_features, device=dev) >= p).float().mul_(1.0 / (1.0 - p))

        x = SF.attention_group(x, relation_embed, [att.a for att in self.attentions],
                               [att.a_2 for att in self.attentions], graph, self.alpha, True, m_att, nanflag)   # 71-72
        if m_x is not None:
            x = x * m_x                                                                                       # 73
        out_relation_1 = SF.matmul(relation_embed, self.W)                                                     # 77
        x = SF.attention_group(x, out_relation_1, [self.out_att.a], [self.out_att.a_2], graph, self.alpha,
                               True, m_out, nanflag)                                                           # 86 (F.elu fused)
        if own_flag:
            check_nanflag(nanflag)
        return x, out_relation_1


class SpKBGATModified(nn.Module):
    def __init__(self, initial_entity_emb, initial_relation_emb, entity_out_dim, relation_out_dim,
                 drop_GAT, alpha, nheads_GAT, initial_entity_emb_params=None, *, sep_space=False):
        """`sep_space=True` adds the GAT_sep_space variant's relation-space projection parameter W_ent2rel
        [R, H*D, H*D] (GAT_sep_space/models.py:136-140), used by SpKBGATConvOnly(..., model_gat) and saved / loaded
        with the state dict; loading a checkpoint that carries the key creates it as well."""
        super().__init__()
        self.num_nodes = initial_entity_emb.shape[0]
        self.entity_in_dim = initial_entity_emb.shape[1]
        self.entity_out_dim_1 = entity_out_dim[0]
        self.nheads_GAT_1 = nheads_GAT[0]
        self.entity_out_dim_2 = entity_out_dim[1]
        self.nheads_GAT_2 = nheads_GAT[1]
        self.num_relation = initial_relation_emb.shape[0]
        self.relation_dim = initial_relation_emb.shape[1]
        self.relation_out_dim_1 = relation_out_dim[0]
        self.drop_GAT = drop_GAT
        self.alpha = alpha
        hd = self.entity_out_dim_1 * self.nheads_GAT_1
        self.final_entity_embeddings = nn.Parameter(torch.randn(self.num_nodes, hd))
        self.final_relation_embeddings = nn.Parameter(torch.randn(self.num_relation, hd))
        self.entity_embeddings = nn.Parameter(initial_entity_emb)
        self.relation_embeddings = nn.Parameter(initial_relation_emb)
        self.sparse_gat_1 = SpGAT(self.num_nodes, self.entity_in_dim, self.entity_out_dim_1, self.relation_dim,
                                  self.drop_GAT, self.alpha, self.nheads_GAT_1)
        self.W_entities = nn.Parameter(torch.zeros(size=(self.entity_in_dim, hd)))
        nn.init.xavier_uniform_(self.W_entities.data, gain=1.414)
        self.nonlinearity_ent2rel = torch.tanh
        if sep_space:
            self.W_ent2rel = nn.Parameter(torch.zeros(size=(self.num_relation, hd, hd)))
            nn.init.xavier_uniform_(self.W_ent2rel.data, gain=1.414)
        self._graph_cache = {}
        self.graph_cache = True

    def load_state_dict(self, state_dict, strict=True, **kw):
        """trained_*.pth of either reference variant loads: a GAT_sep_space checkpoint brings W_ent2rel into a model built
        without it, and a GAT checkpoint leaves an existing W_ent2rel at its initial value."""
        has = "W_ent2rel" in self._parameters
        if "W_ent2rel" in state_dict and not has:
            w = state_dict["W_ent2rel"]
            self.W_ent2rel = nn.Parameter(torch.empty_like(w, device=self.W_entities.device))
        elif has and "W_ent2rel" not in state_dict:
            state_dict = dict(state_dict)
            state_dict["W_ent2rel"] = self.W_ent2rel.detach()
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # -- graph handling ------------------------------------------------------------------------
    def _graph_key(self, adj, train_indices_nhop):
        """identity + version + a cheap content fingerprint (first / middle / last elements, read on the host only for CPU
        tensors) + the model's device; writes that bypass the version counter (numpy views, .data) mostly change it.
        `model.graph_cache = False` disables caching altogether."""
        edge_list, edge_type = adj[0], adj[1]
        has2 = train_indices_nhop is not None and train_indices_nhop.numel() > 0
        tensors = (edge_list, edge_type, train_indices_nhop) if has2 else (edge_list, edge_type)
        dev = self.entity_embeddings.device
        return (str(dev),) + tuple((t.data_ptr(), tuple(t.shape), t._version, str(t.device), _fingerprint(t)) for t in tensors)

    def _graph_is_cached(self, adj, train_indices_nhop):
        if isinstance(adj, KGraph):
            return True
        return bool(self.graph_cache) and self._graph_key(adj, train_indices_nhop) in self._graph_cache

    def prepare_graph(self, adj, train_indices_nhop=None):
        """Build (and cache by tensor identity/version) the CSR/CSC/relation layouts of an edge list."""
        if isinstance(adj, KGraph):
            return adj
        edge_list, edge_type = adj[0], adj[1]
        has2 = train_indices_nhop is not None and train_indices_nhop.numel() > 0
        dev = self.entity_embeddings.device
        if dev.type != "cuda":
            raise RuntimeError("recon_b200.SpKBGATModified must live on a CUDA device (no CPU fallback)")
        key = self._graph_key(adj, train_indices_nhop)
        g = self._graph_cache.get(key) if self.graph_cache else None
        if g is None:
            nhop = train_indices_nhop if has2 else None
            with torch.cuda.device(dev):
                g = KGraph(edge_list, edge_type, nhop, self.num_nodes, self.num_relation, device=dev)
            self._graph_cache.clear()                      # keep one graph: batches change every iteration
            if self.graph_cache:
                self._graph_cache[key] = g
                g._keepalive = (edge_list, edge_type, train_indices_nhop)   # data_ptr keys stay valid while cached
        return g

    def _run(self, entity_embeddings, relation_embeddings, batch_entities, graph, dropout_masks, entities_upgraded=None):
        with torch.cuda.device(entity_embeddings.device):
            return self._run_on(entity_embeddings, relation_embeddings, batch_entities, graph, dropout_masks, entities_upgraded)

    def _run_on(self, entity_embeddings, relation_embeddings, batch_entities, graph, dropout_masks, entities_upgraded=None):
        dev = entity_embeddings.device
        nanflag = torch.zeros(1, dtype=torch.int32, device=dev)
        out_entity_1, out_relation_1 = self.sparse_gat_1(
            None, entity_embeddings, relation_embeddings, None, None, None, None, None,
            graph=graph, dropout_masks=dropout_masks, nanflag=nanflag)
        mask = SF.mask_from_index(torch.as_tensor(batch_entities), entity_embeddings.shape[0], dev, flag=nanflag)   # 167-173
        if entities_upgraded is None:
            entities_upgraded = SF.matmul(entity_embeddings, self.W_entities, getattr(graph, "dist", None))  # 175
        out_entity_1 = SF.ResidualNormFn.apply(entities_upgraded, out_entity_1, mask)                 # 176-179
        check_nanflag(nanflag)
        return out_entity_1, out_relation_1, mask

    def forward(self, Corpus_, batch_entities, adj, train_indices_nhop, dropout_masks=None):
        if self.entity_embeddings.device.type != "cuda":
            raise RuntimeError("recon_b200.SpKBGATModified must live on a CUDA device (no CPU fallback)")
        # models.py:160-161 -- the parameter itself is overwritten with its row-normalised value
        # (rebinds .data to a fresh tensor like the reference does, so a tensor the caller shares with the Parameter and
        # anything an earlier forward saved for backward are left untouched)
        pre = None
        with torch.cuda.device(self.entity_embeddings.device):
            self.entity_embeddings.data = SF.rownorm(self.entity_embeddings.data)
            if not self._graph_is_cached(adj, train_indices_nhop):
                # the layouts are about to be built (host tensors: pack + H2D, the GPU idles meanwhile): queue the one
                # product of the forward that does not need the graph, entity_embeddings.mm(W_entities) (models.py:175)
                pre = SF.matmul(self.entity_embeddings, self.W_entities, None)
        graph = self.prepare_graph(adj, train_indices_nhop)
        out_entity_1, out_relation_1, mask = self._run(self.entity_embeddings, self.relation_embeddings,
                                                       batch_entities, graph, dropout_masks, pre)
        self.final_entity_embeddings.data = out_entity_1.data                                         # 181
        self.final_relation_embeddings.data = out_relation_1.data                                     # 183
        return out_entity_1, out_relation_1, mask

    def batch_test(self, Corpus_, batch_entities, adj, train_indices_nhop, entity_embeddings, dropout_masks=None):
        graph = self.prepare_graph(adj, train_indices_nhop)
        relation_embeddings = self.relation_embeddings.detach()                                       # 191
        entity_embeddings = SF.rownorm(entity_embeddings.data.to(self.W_entities.device))             # 216-217
        return self._run(entity_embeddings, relation_embeddings, batch_entities, graph, dropout_masks)
