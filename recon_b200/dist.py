"""Multi-GPU partitioning of the SpKBGAT path: one process per GPU, torch.distributed (NCCL
over NVLink / NVSwitch on B200; gloo for the CPU tests of this host logic).

1-D row partition: rank g owns a contiguous range of aggregation rows (entities) chosen so that every
rank holds about E/G edges (prefix sum of in-degrees), together with ALL edges whose edge[0] falls in
the range -- the CSR segmented reduce is then purely local. Per layer and direction there is one
exchange step:
    forward : all-gather of the projected gather table P2~ (each rank projects only its own rows)
    backward: reduce-scatter of the partial dP2~ (the transpose of the all-gather),
              all-reduce of dP3~ (per relation) and of the weight gradients (KB..MB).
Gathered node indices are remapped to a padded global numbering  j' = owner(j) * max_rows + (j - lo[owner])
so that the equal-sized all-gather / reduce-scatter buffers are directly indexable by the kernels.
The reference has no distributed path (single process, single device); this is new design.
"""
import torch
import torch.distributed as dist


def balanced_row_bounds(agg_rows, n_rows, world):
    """Row split points [world+1] such that each part holds ~len(agg_rows)/world edges.
    agg_rows: int64 tensor of aggregation rows (edge[0] of every 1-hop and 2-hop edge)."""
    return bounds_from_degrees(torch.bincount(agg_rows, minlength=n_rows), n_rows, world)


def bounds_from_degrees(deg, n_rows, world):
    """Same, from the per-row edge counts (rows whose edges are handled elsewhere carry 0)."""
    cum = torch.cumsum(deg, 0)
    total = int(cum[-1]) if n_rows else 0
    targets = torch.tensor([total * g // world for g in range(1, world)], dtype=cum.dtype, device=cum.device)
    cuts = torch.searchsorted(cum, targets, right=False) + 1 if world > 1 else targets
    bounds = [0] + [int(c) for c in cuts.clamp(max=n_rows)] + [n_rows]
    for i in range(1, len(bounds)):                      # monotone, and no empty middle part swallowing the end
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


class RowPartition:
    """Which rows each rank owns, and the padded global numbering of gathered nodes.
    `hub_ids`: aggregation rows whose in-degree exceeds the per-rank edge budget ("split hub rows across
    GPUs with a final fixed-order add"). They keep their owner for everything row-local, but their EDGES are cut into
    `world` contiguous slices, one per rank; every rank aggregates its slice into a ghost row appended after its own rows
    (local row id n_local + position in hub_ids), the un-normalised partials are summed over the ranks and the owner takes
    the result (recon_b200.functional). Every slot of the padded numbering reserves len(hub_ids) rows for them."""

    def __init__(self, bounds, hub_ids=None):
        self.bounds = list(bounds)
        self.world = len(bounds) - 1
        self.hub_ids = torch.as_tensor(hub_ids if hub_ids is not None else [], dtype=torch.int64).cpu()
        self.n_ghost = int(self.hub_ids.numel())
        self.max_rows = max(1, max(self.bounds[g + 1] - self.bounds[g] for g in range(self.world)) + self.n_ghost)

    def rows_of(self, rank):
        return self.bounds[rank], self.bounds[rank + 1]

    def remap(self, idx):
        """Global entity ids -> padded global ids (owner * max_rows + local offset)."""
        b = torch.tensor(self.bounds, dtype=idx.dtype, device=idx.device)
        owner = torch.searchsorted(b, idx, right=True) - 1
        owner = owner.clamp_(0, self.world - 1)
        return owner * self.max_rows + (idx - b[owner])

    def hub_owner(self):
        """(owner rank, row index inside the owner) of every hub row."""
        b = torch.tensor(self.bounds, dtype=torch.int64)
        owner = (torch.searchsorted(b, self.hub_ids, right=True) - 1).clamp_(0, self.world - 1)
        return owner, self.hub_ids - b[owner]

    def local_edges(self, rank, edge, edge_type, nhop):
        """Edges whose aggregation row is owned by `rank`, in original relative order (1-hop block, 2-hop block),
        with rows made local and gathered nodes remapped; plus this rank's slice of every hub row's edge list, with
        the ghost row ids. Returns (edge[2,E1'], type[E1'], nhop[E2',4], sel1, sel2)."""
        lo, hi = self.rows_of(rank)
        mine = (edge[0] >= lo) & (edge[0] < hi)
        if self.n_ghost:
            hubs = self.hub_ids.to(edge.device)
            in_hub = torch.zeros(int(max(int(edge[0].max()) + 1, int(hubs.max()) + 1)), dtype=torch.bool, device=edge.device)
            in_hub[hubs] = True
            hub_edge = in_hub[edge[0]]
            sel_main = (mine & ~hub_edge).nonzero().flatten()
            he = hub_edge.nonzero().flatten()                               # original order
            order = torch.sort(edge[0, he], stable=True).indices            # grouped by hub row, original order inside
            he = he[order]
            hidx = torch.searchsorted(hubs, edge[0, he])                    # position of the row in hub_ids (ascending)
            cnt = torch.bincount(hidx, minlength=self.n_ghost)
            off = torch.cumsum(cnt, 0) - cnt
            pos = torch.arange(he.numel(), device=edge.device) - off[hidx]
            take = (pos * self.world) // cnt[hidx].clamp_(min=1) == rank    # contiguous 1/world slice of every hub row
            he, hidx = he[take], hidx[take]
            sel1 = torch.cat((sel_main, he))
            rows = torch.cat((edge[0, sel_main] - lo, (hi - lo) + hidx))
        else:
            sel1 = mine.nonzero().flatten()
            rows = edge[0, sel1] - lo
        e_loc = torch.stack((rows, self.remap(edge[1, sel1])), dim=0)
        t_loc = edge_type[sel1]
        if nhop is not None and nhop.numel() > 0:
            sel2 = ((nhop[:, 3] >= lo) & (nhop[:, 3] < hi)).nonzero().flatten()
            nh = nhop[sel2]
            nh_loc = torch.stack((self.remap(nh[:, 0]), nh[:, 1], nh[:, 2], nh[:, 3] - lo), dim=1)
        else:
            sel2 = torch.zeros(0, dtype=torch.int64, device=edge.device)
            nh_loc = torch.zeros((0, 4), dtype=torch.int64, device=edge.device)
        return e_loc, t_loc, nh_loc, sel1, sel2


class _Done:
    """Handle of an exchange that already completed (gloo path)."""

    def wait(self):
        return True


class GhostRows:
    """Exchange helpers for the hub rows that are split across ranks (see RowPartition)."""

    def __init__(self, partition, rank, device, group=None):
        owner, local = partition.hub_owner()
        self.n = partition.n_ghost
        self.group = group
        self.mine = (owner == rank).to(device)                               # [n_ghost] bool: hubs this rank owns
        self.mine_local = local[owner == rank].to(device)                    # their row index among this rank's rows
        self.padded_ids = partition.remap(partition.hub_ids).to(device)      # their row in a gathered table

    def gather_rows(self, local_rows):
        """[n_local, W] -> [n_ghost, W]: the hub rows, from their owners, on every rank."""
        buf = local_rows.new_zeros(self.n, local_rows.shape[1])
        buf[self.mine] = local_rows[self.mine_local]
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
        return buf

    def sum_partials(self, ghost_rows):
        """In-place sum over ranks of the [n_ghost, W] partials every rank computed from its edge slice."""
        dist.all_reduce(ghost_rows, op=dist.ReduceOp.SUM, group=self.group)
        return ghost_rows


class DistContext:
    """Collectives of the partitioned path; attached to a KGraph as `graph.dist`."""

    def __init__(self, partition, rank, group=None, device=None):
        self.part = partition
        self.rank = rank
        self.group = group
        self.world = partition.world
        self.ghost = GhostRows(partition, rank, device, group) if partition.n_ghost else None

    @property
    def n_local(self):
        lo, hi = self.part.rows_of(self.rank)
        return hi - lo

    # ---- zero-copy exchange buffers -------------------------------------------------------------------------------
    # The gathered table of a layer lives in ONE buffer [world * max_rows, W] in padded global numbering. The producer
    # kernel (GEMM epilogue / table builder) writes this rank's rows straight into its slot, the all-gather runs in place,
    # and the reduce-scatter of a partial gradient lands in a [max_rows, W] buffer whose first n_local rows are the result:
    # no pad / slice / contiguous copies around the collectives. Both collectives are started asynchronously (NCCL's own
    # stream) and return a handle whose wait() orders the current stream after them, so independent kernels overlap.

    def gather_buffer(self, width, device, dtype=torch.float32):
        """([world * max_rows, width] buffer, view of this rank's n_local rows inside it). Pad rows of the own slot are zero."""
        mr = self.part.max_rows
        buf = torch.empty(self.world * mr, width, dtype=dtype, device=device)
        lo = self.rank * mr
        if self.n_local < mr:
            buf[lo + self.n_local:lo + mr].zero_()
        return buf, buf[lo:lo + self.n_local]

    def all_gather_start(self, buf):
        """In-place all-gather of every rank's slot of `buf`; returns a handle with wait()."""
        mr = self.part.max_rows
        mine = buf[self.rank * mr:(self.rank + 1) * mr]
        if dist.get_backend(self.group) == "gloo":           # gloo: no aliasing of input and output
            dist.all_gather_into_tensor(buf, mine.clone(), group=self.group)
            return _Done()
        return dist.all_gather_into_tensor(buf, mine, group=self.group, async_op=True)

    def reduce_scatter_start(self, partial_all, out_pad):
        """Sum over ranks of the [world * max_rows, W] partials; this rank's slot lands in out_pad [max_rows, W]."""
        assert partial_all.is_contiguous() and out_pad.is_contiguous() and out_pad.shape[0] == self.part.max_rows
        if dist.get_backend(self.group) == "gloo":           # gloo has no reduce_scatter: all-reduce and slice
            full = partial_all.clone()
            dist.all_reduce(full, op=dist.ReduceOp.SUM, group=self.group)
            out_pad.copy_(full[self.rank * self.part.max_rows:(self.rank + 1) * self.part.max_rows])
            return _Done()
        return dist.reduce_scatter_tensor(out_pad, partial_all, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def all_gather_rows(self, local_rows):
        """[n_local, W] (any row stride) -> [world * max_rows, W] in padded global numbering (copying convenience form)."""
        buf, mine = self.gather_buffer(local_rows.shape[1], local_rows.device, local_rows.dtype)
        mine.copy_(local_rows)
        self.all_gather_start(buf).wait()
        return buf

    def reduce_scatter_rows(self, partial_all, out_local):
        """Sum over ranks of [world * max_rows, W] partials; this rank's rows land in out_local [n_local, W]."""
        pad = partial_all.new_empty(self.part.max_rows, partial_all.shape[1])
        self.reduce_scatter_start(partial_all.contiguous(), pad).wait()
        out_local.copy_(pad[: out_local.shape[0]])
        return out_local

    def all_reduce(self, t):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def all_reduce_start(self, t):
        """Asynchronous all-reduce (sum) of a small replicated gradient; returns a handle with wait()."""
        if dist.get_backend(self.group) == "gloo":
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return _Done()
        return dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group, async_op=True)


class PartitionedKBGAT:
    """SpKBGATModified over a row partition: this rank holds its entity rows (embeddings, outputs, gradients),
    the replicated relation / attention parameters, and the CSR / CSC layouts of its own edges."""

    def __init__(self, n_ent, n_rel, edge, edge_type, nhop, in_dim, out_dim, nheads, alpha, device,
                 entity_emb=None, relation_emb=None, seed=0, group=None, state_dict=None, hub_split=0.05):
        from .models import SpKBGATModified
        from .graph import KGraph
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.device = device
        edge = edge.to(device); edge_type = edge_type.to(device)
        nhop = nhop.to(device) if nhop is not None and nhop.numel() else None
        agg = edge[0] if nhop is None else torch.cat((edge[0], nhop[:, 3]))
        deg = torch.bincount(agg, minlength=n_ent)
        # rows holding more than hub_split x (edges per rank) are split across the ranks (graphs without 2-hop rows)
        hub_ids = None
        if self.world > 1 and hub_split and nhop is None:
            hub_ids = (deg > int(hub_split * agg.numel() / self.world)).nonzero().flatten()
            if hub_ids.numel() == 0 or hub_ids.numel() > 1024:
                hub_ids = None
            else:
                deg[hub_ids] = 0
        self.part = RowPartition(bounds_from_degrees(deg, n_ent, self.world), hub_ids)
        del agg, deg
        lo, hi = self.part.rows_of(self.rank)
        self.lo, self.hi = lo, hi
        e_loc, t_loc, nh_loc, _, _ = self.part.local_edges(self.rank, edge, edge_type, nhop)
        self.n_edges_total = edge.shape[1] + (0 if nhop is None else nhop.shape[0])
        self.n_edges_local = e_loc.shape[1] + nh_loc.shape[0]
        del edge, edge_type, nhop
        g = torch.Generator().manual_seed(seed)
        if relation_emb is None:
            relation_emb = torch.randn(n_rel, in_dim, generator=g)
        if entity_emb is None:                               # synthetic: only this rank's rows are ever materialised
            ent_loc = torch.randn(hi - lo, in_dim, generator=torch.Generator().manual_seed(seed + 1000 + self.rank))
        else:
            ent_loc = entity_emb[lo:hi].clone()
        torch.manual_seed(seed)                              # identical replicated parameters on every rank
        self.model = SpKBGATModified(ent_loc, relation_emb.clone(), [out_dim, 2 * out_dim],
                                     [out_dim, 2 * out_dim], 0.0, alpha, [nheads, nheads], None)
        if state_dict is not None:
            sd = {k: (v[lo:hi] if k in ("entity_embeddings", "final_entity_embeddings") else v) for k, v in state_dict.items()}
            self.model.load_state_dict(sd)
        self.model = self.model.to(device)
        self.n_rel = n_rel
        self.group = group
        self.graph = KGraph(e_loc, t_loc, nh_loc if nh_loc.numel() else None, hi - lo + self.part.n_ghost, n_rel,
                            device=device, n_cols=self.world * self.part.max_rows)
        self.graph.dist = DistContext(self.part, self.rank, group, device)
        self.graph.n_ghost = self.part.n_ghost
        self._local_edges_dev = (e_loc, t_loc, nh_loc)
        self.batch = torch.arange(hi - lo, device=device)
        self.g_ent = None
        self.g_rel = None

    def set_loss_weights(self, g_ent_full, g_rel):
        self.g_ent = g_ent_full[self.lo:self.hi].to(self.device)
        self.g_rel = g_rel.to(self.device)

    def e2e(self, steps):
        """End-to-end step with HOST edge tensors: pinned H2D of this rank's int64 edge list, device CSR / CSC /
        relation rebuild, forward + backward, loss read-back. Returns (seconds per step, h2d bytes per step)."""
        import time
        from .graph import KGraph
        host = tuple(t.cpu() for t in self._local_edges_dev)
        n_loc_edges = host[0].shape[1] + host[2].shape[0]
        h2d = 4 * (4 if host[2].numel() else 3) * n_loc_edges          # int32 staging of the int64 tensors (graph.py)
        res = torch.empty(1, dtype=torch.float32).pin_memory()
        times = []
        for i in range(steps + 1):
            dist.barrier(group=self.group)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e_loc, t_loc, nh_loc = host
            graph = KGraph(e_loc, t_loc, nh_loc if nh_loc.numel() else None, self.hi - self.lo + self.part.n_ghost,
                           self.n_rel, device=self.device, n_cols=self.world * self.part.max_rows)
            graph.dist = self.graph.dist
            graph.n_ghost = self.part.n_ghost
            _, _, loss = self.step(graph)
            res.copy_(loss.detach().reshape(1), non_blocking=True)
            torch.cuda.synchronize()
            if i > 0:
                times.append(time.perf_counter() - t0)
        return sum(times) / len(times), h2d

    def step(self, graph=None):
        """One forward + backward; loss = <out_entity, G_e> + <out_relation, G_r> with the relation term counted once."""
        self.model.zero_grad(set_to_none=True)
        out_e, out_r, _ = self.model(None, self.batch, graph if graph is not None else self.graph, None)
        if self.g_ent is None:
            gen = torch.Generator().manual_seed(1)
            self.g_ent = torch.randn(out_e.shape, generator=gen).to(self.device)
            self.g_rel = torch.randn(out_r.shape, generator=gen).to(self.device)
        from . import functional as SF
        loss = SF.linear_loss_backward((out_e, out_r), (self.g_ent, self.g_rel))
        return out_e, out_r, loss


def parity_check(device, group=None, n=20000, e=200000, r=64, n_nhop=0, seed=21, in_dim=50, out_dim=100, nheads=2,
                 alpha=0.2, zipf=1.1, hub_frac=0.2):
    """Multi-GPU parity against the single-GPU path of this library on one small seeded KG ("results
    identical (<= 1e-6 rel) to the 1-GPU run"): every rank runs the row-partitioned step, rank 0 also runs the whole
    graph on its own GPU; returns {"max_rel": worst rel-L2 over out_entity, out_relation and every gradient, ...} on
    every rank. Used by bench.py (printed in the JSON line at N > 1) and tests/test_dist_gpu.py."""
    from .models import SpKBGATModified
    from .synth import make_kg
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    edge, etype, nhop = make_kg(n, e, r, alpha=zipf, n_nhop=n_nhop, seed=seed, hub_frac=hub_frac)
    gen = torch.Generator().manual_seed(seed + 1)
    ent = torch.randn(n, in_dim, generator=gen)
    rel = torch.randn(r, in_dim, generator=gen)
    torch.manual_seed(seed)
    ref = SpKBGATModified(ent.clone(), rel.clone(), [out_dim, 2 * out_dim], [out_dim, 2 * out_dim], 0.0, alpha,
                          [nheads, nheads], None)
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    g_ent = torch.randn(n, out_dim * nheads, generator=gen)
    g_rel = torch.randn(r, out_dim * nheads, generator=gen)
    pk = PartitionedKBGAT(n, r, edge, etype, nhop, in_dim, out_dim, nheads, alpha, device, group=group, state_dict=sd)
    pk.set_loss_weights(g_ent, g_rel)
    out_e, out_r, _ = pk.step()
    torch.cuda.synchronize()
    lo, hi = pk.lo, pk.hi
    gathered = {}
    for k, t in (("out_entity", out_e.detach()), ("grad.entity_embeddings", pk.model.entity_embeddings.grad)):
        full = torch.zeros(n, t.shape[1], device=device)
        full[lo:hi] = t
        dist.all_reduce(full, group=group)
        gathered[k] = full
    res = torch.zeros(2, dtype=torch.float64, device=device)
    if rank == 0:
        def rl(a, b):
            a = a.detach().double(); b = b.detach().double()
            return float((a - b).norm() / b.norm().clamp_min(1e-30))
        m = ref.to(device)
        m.load_state_dict(sd)
        oe, orl, _ = m(None, torch.arange(n), (edge, etype), nhop)
        ((oe * g_ent.to(device)).sum() + (orl * g_rel.to(device)).sum()).backward()
        errs = {"out_entity": rl(gathered["out_entity"], oe), "out_relation": rl(out_r, orl),
                "grad.entity_embeddings": rl(gathered["grad.entity_embeddings"], m.entity_embeddings.grad)}
        refp = dict(m.named_parameters())
        for k, prm in pk.model.named_parameters():
            if k not in ("entity_embeddings", "final_entity_embeddings", "final_relation_embeddings") and prm.grad is not None:
                errs["grad." + k] = rl(prm.grad, refp[k].grad)
        res[0] = max(errs.values()); res[1] = len(errs)
    dist.broadcast(res, 0, group=group)
    return {"max_rel": float(res[0]), "tensors": int(res[1]), "world": world,
            "vs": f"single-GPU path on the same KG (N={n} E={e + n_nhop} R={r}, Zipf {zipf})"}
