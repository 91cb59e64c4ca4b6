"""Synthetic power-law knowledge graphs of the shapes BASELINE.json names.

edge[1] (gather index = triple head) is uniform; edge[0] (aggregation row = triple tail,
GAT/preprocess.py:78-80) is Zipf: rank = floor(Pareto(x_m=1, alpha)) mapped through a fixed
random permutation of [0, N) and clipped to N-1; relation types are uniform. Edge order is
generation order (unsorted). CPU generation (torch.Generator) so every rank, the CPU checker and
the CUDA path see identical inputs.
"""
import math
import torch


def zipf_rows(n_nodes, n_edges, alpha, gen, device="cpu"):
    if alpha is None or math.isinf(alpha):
        return torch.randint(0, n_nodes, (n_edges,), generator=gen, dtype=torch.int64, device=device)
    u = torch.rand(n_edges, generator=gen, dtype=torch.float64, device=device).clamp_(min=1e-12)
    rank = torch.floor(u.pow_(-1.0 / alpha)).clamp_(max=float(n_nodes)).to(torch.int64) - 1
    perm = torch.randperm(n_nodes, generator=gen, device=device)
    return perm[rank.clamp_(0, n_nodes - 1)]


def make_kg(n_nodes, n_edges, n_rel, alpha=None, n_nhop=0, seed=0, device="cpu", hub_frac=1.0):
    """Returns (edge int64[2,E1] = [rows(tail); cols(head)], edge_type int64[E1],
    nhop int64[E2,4] = [s, r1, r2, t] rows with t Zipf / s uniform, like 1-hop).
    hub_frac < 1: only that share of the edges gets a Pareto row, the rest uniform rows (a power-law tail over a
    uniform background). device: generate on that device (same seed + same device type -> same graph on every rank)."""
    gen = torch.Generator(device=device).manual_seed(seed)
    kw = dict(generator=gen, dtype=torch.int64, device=device)

    def agg_rows(m):
        rows = zipf_rows(n_nodes, m, alpha, gen, device)
        if hub_frac < 1.0:
            uni = torch.rand(m, generator=gen, device=device) >= hub_frac
            rows = torch.where(uni, torch.randint(0, n_nodes, (m,), **kw), rows)
        return rows

    rows = agg_rows(n_edges)
    cols = torch.randint(0, n_nodes, (n_edges,), **kw)
    etype = torch.randint(0, n_rel, (n_edges,), **kw)
    edge = torch.stack((rows, cols), dim=0)
    if n_nhop > 0:
        t = agg_rows(n_nhop)
        s = torch.randint(0, n_nodes, (n_nhop,), **kw)
        r = torch.randint(0, n_rel, (n_nhop, 2), **kw)
        nhop = torch.stack((s, r[:, 0], r[:, 1], t), dim=1)
    else:
        nhop = torch.zeros((0, 4), dtype=torch.int64, device=device)
    return edge, etype, nhop


def make_triples(n_nodes, n_triples, n_rel, seed=0, multi_edge_frac=0.1, self_loop_frac=0.02):
    """Random (head, rel, tail) triple list with parallel edges and self loops, for the
    edge-construction tests (the cases Corpus.bfs treats specially)."""
    gen = torch.Generator().manual_seed(seed)
    h = torch.randint(0, n_nodes, (n_triples,), generator=gen)
    t = torch.randint(0, n_nodes, (n_triples,), generator=gen)
    r = torch.randint(0, n_rel, (n_triples,), generator=gen)
    n_multi = int(n_triples * multi_edge_frac)
    if n_multi:
        src = torch.randint(0, n_triples, (n_multi,), generator=gen)
        dst = torch.randint(0, n_triples, (n_multi,), generator=gen)
        h[dst], t[dst] = h[src], t[src]
    n_self = int(n_triples * self_loop_frac)
    if n_self:
        idx = torch.randint(0, n_triples, (n_self,), generator=gen)
        t[idx] = h[idx]
    return torch.stack((h, r, t), dim=1)
