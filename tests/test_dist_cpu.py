"""Host logic of the multi-GPU path on CPU: row partition, index remapping and the three collectives,
with world_size = 2 over gloo. The local math uses the CPU oracle as a stand-in for the kernels (tests may)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from recon_b200.dist import RowPartition, DistContext, balanced_row_bounds
from recon_b200.synth import make_kg


def test_balanced_bounds_cover_and_balance():
    edge, _, nhop = make_kg(5000, 80000, 7, alpha=1.5, n_nhop=9000, seed=2)
    agg = torch.cat((edge[0], nhop[:, 3]))
    for world in (1, 2, 3, 8):
        b = balanced_row_bounds(agg, 5000, world)
        assert b[0] == 0 and b[-1] == 5000 and all(b[i] <= b[i + 1] for i in range(world))
        deg = torch.bincount(agg, minlength=5000)
        per = [int(deg[b[g]:b[g + 1]].sum()) for g in range(world)]
        assert sum(per) == agg.numel()
        heavy = int(deg.max())
        assert max(per) <= agg.numel() // world + heavy + 1        # balanced up to one (hub) row


def test_remap_and_local_edges_partition_every_edge_once():
    n = 1000
    edge, et, nhop = make_kg(n, 12000, 5, alpha=None, n_nhop=3000, seed=3)
    part = RowPartition(balanced_row_bounds(torch.cat((edge[0], nhop[:, 3])), n, 4))
    seen1 = torch.zeros(edge.shape[1], dtype=torch.int32)
    seen2 = torch.zeros(nhop.shape[0], dtype=torch.int32)
    for r in range(4):
        e_loc, t_loc, nh_loc, s1, s2 = part.local_edges(r, edge, et, nhop)
        lo, hi = part.rows_of(r)
        seen1[s1] += 1; seen2[s2] += 1
        assert bool((e_loc[0] >= 0).all()) and bool((e_loc[0] < hi - lo).all())
        assert torch.equal(t_loc, et[s1])
        # remapped gather ids decode back to the original entity ids
        owner = e_loc[1] // part.max_rows
        back = torch.tensor(part.bounds)[owner] + e_loc[1] % part.max_rows
        assert torch.equal(back, edge[1, s1])
        assert bool((s1[1:] > s1[:-1]).all())                      # original relative order kept
    assert bool((seen1 == 1).all()) and bool((seen2 == 1).all())


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, e, r, f, w = 300, 4000, 5, 6, 8
        edge, et, nhop = make_kg(n, e, r, alpha=1.5, n_nhop=500, seed=4)
        g = torch.Generator().manual_seed(0)
        X = torch.randn(n, f, generator=g, dtype=torch.float64); W = torch.randn(f, w, generator=g, dtype=torch.float64)
        part = RowPartition(balanced_row_bounds(torch.cat((edge[0], nhop[:, 3])), n, world))
        ctx = DistContext(part, rank)
        lo, hi = part.rows_of(rank)
        e_loc, _, nh_loc, _, _ = part.local_edges(rank, edge, et, nhop)
        rows = torch.cat((e_loc[0], nh_loc[:, 3])); cols = torch.cat((e_loc[1], nh_loc[:, 0]))
        # forward exchange: project own rows, all-gather, aggregate own rows
        P2_all = ctx.all_gather_rows(X[lo:hi] @ W)
        out_loc = torch.zeros(hi - lo, w, dtype=torch.float64).index_add_(0, rows, P2_all[cols])
        # backward exchange: partial gradient over all gathered nodes -> reduce-scatter to owners; weight grad all-reduce
        gout = torch.ones(hi - lo, w, dtype=torch.float64) * (rank + 1)
        dP2_all = torch.zeros_like(P2_all).index_add_(0, cols, gout[rows])
        dP2_loc = torch.empty(hi - lo, w, dtype=torch.float64)
        ctx.reduce_scatter_rows(dP2_all, dP2_loc)
        dW = ctx.all_reduce(X[lo:hi].t() @ dP2_loc)
        # single-process reference
        R = torch.cat((edge[0], nhop[:, 3])); Cc = torch.cat((edge[1], nhop[:, 0]))
        out_ref = torch.zeros(n, w, dtype=torch.float64).index_add_(0, R, (X @ W)[Cc])
        owner_scale = torch.zeros(n, dtype=torch.float64)
        for q in range(world):
            a, b = part.rows_of(q); owner_scale[a:b] = q + 1
        dP2_ref = torch.zeros(n, w, dtype=torch.float64).index_add_(0, Cc, owner_scale[R].unsqueeze(1).expand(-1, w).contiguous())
        ok = (torch.allclose(out_loc, out_ref[lo:hi], atol=1e-10) and torch.allclose(dP2_loc, dP2_ref[lo:hi], atol=1e-10)
              and torch.allclose(dW, X.t() @ dP2_ref, atol=1e-8))
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_world2_gloo_exchange_matches_single_process():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
