"""N3: the per-iteration batch sampler of the reference's training loop, on the device.

  Corpus.get_batch_adj_data            GAT/create_batch.py:391-436   -> nhop.TripleGraph.batch_edges (K0b)
  Corpus.get_batch_nhop_neighbors_all  GAT/create_batch.py:871-895   -> nhop.TripleGraph.batch_edges (K0b)
  Corpus.get_iteration_triples_batch   GAT/create_batch.py:262-351   -> TripleSampler.get_iteration_triples_batch
  valid_triples_dict                   GAT/create_batch.py:82-83     -> sorted int64 keys + binary search

The positives of an iteration are the batch adjacency read as (head = batch entity, relation, tail = neighbour): the
reference walks the same `node_neighbors_1hop[ent][1]` lists in the same order for both (create_batch.py:267-273 and
413-429), so they are bit-identical to the reference's, order included. The corrupted copies follow the reference's row
layout and validity rule; their random numbers can be supplied (`random_entities`, `random_relations`: the arrays the
reference draws at create_batch.py:293-296) or are generated on the device from `seed`.
All arithmetic runs in libspkbgat (`spk_triple_keys`, `spk_corrupt_triples`, the K0 radix sort); no CPU fallback.
"""
import torch

from . import _lib
from .nhop import TripleGraph, _stable_order


class TripleSampler:
    """`train_triples` int [E,3] = (head, rel, tail) in file order: the graph whose 1-hop lists give the positives.
    `valid_triples` (default: the same list) is what the reference puts in valid_triples_dict: train + validation + test."""

    def __init__(self, train_triples, n_entities, n_relations, valid_triples=None, invalid_valid_ratio=2, device=None):
        lib = _lib.load()
        if device is None:
            device = train_triples.device if train_triples.is_cuda else torch.device("cuda", torch.cuda.current_device())
        self.device = device
        self.n_entities, self.n_relations = int(n_entities), int(n_relations)
        self.invalid_valid_ratio = int(invalid_valid_ratio)
        self.graph = TripleGraph(train_triples, n_entities, device=device)
        vt = (train_triples if valid_triples is None else valid_triples).to(device=device, dtype=torch.int64).contiguous()
        m = int(vt.shape[0])
        if m:
            order = _stable_order([vt[:, 2].contiguous(), vt[:, 1].contiguous(), vt[:, 0].contiguous()], device)
            vt = vt.index_select(0, order).contiguous()          # (h, r, t) lexicographic
        self.valid_keys = torch.empty(m, dtype=torch.int64, device=device)
        err = torch.zeros(1, dtype=torch.int32, device=device)
        _lib.check(lib.spk_triple_keys(_lib.ptr(vt) if m else None, m, self.n_entities, self.n_relations,
                                       self.valid_keys.data_ptr(), err.data_ptr(), _lib.stream_ptr()), "triple_keys")
        if int(err.item()) != 0:
            raise IndexError("valid_triples: entity / relation id out of range")

    def positive_triples(self, batch_entities):
        """int64 [P,3] on the device, in the reference's order (create_batch.py:267-273)."""
        if len(batch_entities) == 0:
            return torch.zeros((0, 3), dtype=torch.int64, device=self.device)
        adj_idx, adj_val, _ = self.graph.batch_edges(batch_entities, want_nhop=False)
        return torch.stack((adj_idx[1], adj_val, adj_idx[0]), dim=1).contiguous()

    def get_iteration_triples_batch(self, batch_entities, invalid_valid_ratio=None, random_entities=None,
                                    random_relations=None, seed=0):
        """Drop-in for Corpus.get_iteration_triples_batch: returns (batch_indices int64 [T,3], batch_values float32 [T,1])
        on the device (the reference returns int32 / float32 numpy arrays that main.py:503-505 turns into these)."""
        lib = _lib.load()
        ratio = self.invalid_valid_ratio if invalid_valid_ratio is None else int(invalid_valid_ratio)
        pos = self.positive_triples(batch_entities)
        p = int(pos.shape[0])
        total = p * (2 * ratio + 1)
        out = torch.empty(total, 3, dtype=torch.int64, device=self.device)
        val = torch.empty(total, 1, dtype=torch.float32, device=self.device)
        if p == 0:
            return out, val

        def draws(x, what):
            if x is None:
                return None
            x = torch.as_tensor(x).to(device=self.device, dtype=torch.int64).contiguous()
            if x.numel() != p * ratio:
                raise ValueError(f"{what} must hold P * ratio = {p * ratio} draws (create_batch.py:293-296)")
            return x

        ie, ir = draws(random_entities, "random_entities"), draws(random_relations, "random_relations")
        _lib.check(lib.spk_corrupt_triples(pos.data_ptr(), p, ratio, _lib.ptr(self.valid_keys) if self.valid_keys.numel() else None,
                                           self.valid_keys.numel(), self.n_entities, self.n_relations,
                                           _lib.ptr(ie), _lib.ptr(ir), int(seed) & (2 ** 64 - 1),
                                           out.data_ptr(), val.data_ptr(), _lib.stream_ptr()), "corrupt_triples")
        return out, val
