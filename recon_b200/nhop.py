"""K0b: on-device construction of the batch adjacency and the 2-hop path rows, bit-exact (values AND order)
with the reference's Python dict / queue code:

  Corpus.get_graph                    GAT/create_batch.py:708-732  graph[head][tail] = [rel, ...] (insertion = file order)
  Corpus.bfs                          GAT/create_batch.py:788-842  FIFO BFS, first discoverer is the parent
  Corpus.get_batch_adj_data           GAT/create_batch.py:391-436  -> ([trgts; srcs], vals)
  Corpus.get_batch_nhop_neighbors_all GAT/create_batch.py:871-895  -> rows [s, r(s->m)[0], r(m->t)[0], t]

`TripleGraph.batch_edges` runs entirely in libspkbgat (spk_nhop_build, csrc/spk_nhop.cu: prefix sums, level expansion,
stable radix sorts, compaction; buffers come from PyTorch's allocator through a callback). Only the once-per-graph
construction of the distinct-neighbour adjacency below still uses torch for index glue around the library's sorts.

Instead of a per-source BFS the whole batch is expanded at once and the BFS tie-breaks are recovered by
STABLE sorts (libspkbgat's radix sort, graph.sort_pairs):
  * first-occurrence order of the distinct out-neighbours of every head = stable sort of the (head, tail)
    groups by the file index of their first triple;
  * a 2-hop target t of source s is kept iff it is not s, not a 1-hop neighbour of s, and this is the first
    candidate (in (mid, tail) discovery order) reaching t: sort the records (blockers first, then candidates in
    discovery order) stably by (s, t) and keep group heads that are candidates.
"""
import ctypes as C

import torch

from . import _lib
from .graph import sort_pairs, _key_bits


def _stable_order(keys_list, device):
    """Permutation that sorts stably by keys_list[-1] (major) ... keys_list[0] (minor): LSD over the given keys."""
    n = keys_list[0].numel()
    perm = torch.arange(n, dtype=torch.int32, device=device)
    for k in keys_list:
        kk = k.index_select(0, perm.long()).to(torch.int32).contiguous()
        bits = _key_bits(int(k.max().item()) + 1) if n else 1
        _, perm = sort_pairs(kk, perm.contiguous(), bits)
    return perm.long()


def _expand(counts):
    """For segment sizes counts[n] -> (segment id of every element, offset inside its segment)."""
    total = int(counts.sum().item())
    seg = torch.repeat_interleave(torch.arange(counts.numel(), device=counts.device), counts, output_size=total)
    start = torch.cumsum(counts, 0) - counts
    local = torch.arange(total, device=counts.device) - start[seg]
    return seg, local


class TripleGraph:
    """Distinct-neighbour adjacency of a triple list in the reference's insertion order."""

    def __init__(self, triples, n_nodes, device=None):
        if device is None:
            device = triples.device if triples.is_cuda else torch.device("cuda", torch.cuda.current_device())
        tr = triples.to(device=device, dtype=torch.int64)
        self.device, self.n_nodes = device, int(n_nodes)
        h, r, t = tr[:, 0].contiguous(), tr[:, 1].contiguous(), tr[:, 2].contiguous()
        e = h.numel()
        # (A) group the triples by (head, tail), file order kept inside a group
        p = _stable_order([t, h], device)
        hs, ts, self.rs = h[p], t[p], r[p]
        head = torch.ones(e, dtype=torch.bool, device=device)
        if e > 1:
            head[1:] = (hs[1:] != hs[:-1]) | (ts[1:] != ts[:-1])
        gstart = head.nonzero().flatten()
        gend = torch.cat((gstart[1:], torch.tensor([e], device=device)))
        first_idx = p[gstart]                                   # file index of the first triple of the pair
        # (B) distinct tails of every head in first-occurrence order
        q = _stable_order([first_idx, hs[gstart]], device)
        self.uh, self.ut = hs[gstart][q], ts[gstart][q]
        self.ur0 = self.rs[gstart][q]                            # first relation of the pair: graph[h][t][0]
        self.ugs, self.uge = gstart[q], gend[q]                  # its parallel relations: rs[ugs:uge]
        self.uptr = torch.searchsorted(self.uh, torch.arange(self.n_nodes + 1, device=device))
        self._i32 = None

    def batch_edges(self, batch_sources, partial_2hop=False, want_nhop=True):
        """Returns (adj_indices int64[2,E1], adj_values int64[E1], nhop int32[E2,4]) for the given ordered sources
        (one spk_nhop_build call; every kernel of it is libspkbgat's)."""
        lib = _lib.load()
        dev = self.device
        s = torch.as_tensor(batch_sources, dtype=torch.int64, device=dev).contiguous()
        if self._i32 is None:
            self._i32 = tuple(t.to(torch.int32).contiguous() for t in (self.uptr, self.ut, self.ur0, self.ugs, self.uge, self.rs))
        g = _lib.TripleGraphArgs()
        g.uptr, g.ut, g.ur0, g.ugs, g.uge, g.rs = (t.data_ptr() for t in self._i32)
        g.n_nodes, g.n_pairs, g.n_triples = self.n_nodes, self.ut.numel(), self.rs.numel()
        bufs = {}

        def alloc(_ctx, nbytes):
            try:
                t = torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=dev)
            except RuntimeError:                              # out of memory: the C side reports it
                return None
            bufs[t.data_ptr()] = t
            return t.data_ptr()

        res = _lib.NhopResult()
        with torch.cuda.device(dev):
            rc = lib.spk_nhop_build(C.byref(g), s.data_ptr() if s.numel() else None, s.numel(),
                                    (1 if partial_2hop else 0) | (0 if want_nhop else 2), _lib.ALLOC_FN(alloc), None,
                                    C.byref(res), _lib.stream_ptr())
        if rc == 5:
            raise IndexError("batch source id out of range")
        _lib.check(rc, "nhop_build")
        e1, e2 = int(res.e1), int(res.e2)
        adj_idx = bufs[res.adj_idx].view(torch.int64)[:2 * e1].view(2, e1)
        adj_val = bufs[res.adj_val].view(torch.int64)[:e1]
        nhop = bufs[res.nhop].view(torch.int32)[:4 * e2].view(e2, 4) if res.nhop else torch.zeros((0, 4), dtype=torch.int32, device=dev)
        return adj_idx, adj_val, nhop

    def batch_edges_torch(self, batch_sources, partial_2hop=False, want_nhop=True):
        """The same construction written with torch index ops on top of the library's sorts (round-1 form); kept as an
        independent cross-check of spk_nhop_build for the tests, not used by the product."""
        dev = self.device
        s = torch.as_tensor(batch_sources, dtype=torch.int64, device=dev)
        nb = s.numel()
        # (C) level 1: distinct out-neighbours of every source, self loops dropped (already visited)
        cnt1 = self.uptr[s + 1] - self.uptr[s]
        l1_b, loc = _expand(cnt1)
        l1_a = self.uptr[s[l1_b]] + loc
        l1_m = self.ut[l1_a]
        keep1 = l1_m != s[l1_b]
        l1_b, l1_a, l1_m = l1_b[keep1], l1_a[keep1], l1_m[keep1]
        # (D) batch adjacency: every parallel relation of every kept pair
        e_l1, eloc = _expand(self.uge[l1_a] - self.ugs[l1_a])
        adj_idx = torch.stack((l1_m[e_l1], s[l1_b[e_l1]]), dim=0)
        adj_val = self.rs[self.ugs[l1_a[e_l1]] + eloc]
        if not want_nhop:
            return adj_idx, adj_val, torch.zeros((0, 4), dtype=torch.int32, device=dev)
        # (E) level-2 candidates in discovery order: (source, mid in level-1 order, tail in mid's order)
        cnt2 = self.uptr[l1_m + 1] - self.uptr[l1_m]
        c_l1, cloc = _expand(cnt2)
        c_c = self.uptr[l1_m[c_l1]] + cloc
        c_t = self.ut[c_c]
        c_b = l1_b[c_l1]
        n_block = nb + l1_b.numel()
        # (F) blockers (the source itself, its level-1 nodes) then candidates; first record of each (b, t) group wins
        rb = torch.cat((torch.arange(nb, device=dev), l1_b, c_b))
        rt = torch.cat((s, l1_m, c_t))
        o = _stable_order([rt, rb], dev)
        ob, ot = rb[o], rt[o]
        head = torch.ones(o.numel(), dtype=torch.bool, device=dev)
        if o.numel() > 1:
            head[1:] = (ob[1:] != ob[:-1]) | (ot[1:] != ot[:-1])
        acc = o[head & (o >= n_block)] - n_block                 # accepted candidate ids
        if acc.numel():
            keys, _ = sort_pairs(acc.to(torch.int32).contiguous(), torch.zeros_like(acc, dtype=torch.int32),
                                 _key_bits(int(c_t.numel()) + 1))
            acc = keys.long()                                    # back to discovery order
        if partial_2hop and acc.numel():                         # create_batch.py:883-884: first path of each source only
            ab = c_b[acc]
            first = torch.ones(acc.numel(), dtype=torch.bool, device=dev)
            first[1:] = ab[1:] != ab[:-1]
            acc = acc[first]
        l1 = c_l1[acc]
        nhop = torch.stack((s[c_b[acc]], self.ur0[l1_a[l1]], self.ur0[c_c[acc]], c_t[acc]), dim=1).to(torch.int32)
        return adj_idx, adj_val, nhop


def build_batch_edges(triples, n_nodes, batch_sources, partial_2hop=False):
    """Convenience wrapper: triple list [E,3] = (head, rel, tail) in file order -> reference-identical batch tensors."""
    return TripleGraph(triples, n_nodes).batch_edges(batch_sources, partial_2hop)
