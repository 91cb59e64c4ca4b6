"""TEST INFRASTRUCTURE ONLY (checker for tests/) -- never imported by recon_b200.

CPU restatement of the reference's per-iteration triple sampler:
  * Corpus.get_iteration_triples_batch      /root/reference/GAT/create_batch.py:262-351
  * valid_triples_dict                      /root/reference/GAT/create_batch.py:82-83
Random numbers come from `rng` (anything with numpy's `randint(low, high[, size])`); with `rng = np.random` after
`np.random.seed(k)` the draws are consumed in the reference's order, so the output is bit-identical to the reference's
(pinned in tests/test_sampler.py against fixtures written by the reference's own Corpus method).

Row layout for P positives and ratio = invalid_valid_ratio (rows P.. are first filled with 2*ratio tiled copies of the
positives, create_batch.py:298-301, then some are overwritten):
  [0, P)                              positives (head = batch entity, rel, tail = 1-hop neighbour), value +1
  P + [0, P*(ratio//2))               head replaced by a random entity, value -1            (create_batch.py:304-313)
  P + [P*(ratio//2), 2*P*(ratio//2))  tail replaced, value -1                               (315-326)
  P + [2*P*(ratio//2), P*ratio)       untouched copies of the positives, value +1 (only when ratio is odd)
  P + [P*ratio, 2*P*ratio)            relation replaced, value -1; left as a +1 copy when no invalid relation was
                                      found in n_relations redraws                          (328-347)
Row P + s is a copy of positive s mod P, which is the pairing batch_gat_loss uses (main.py:351).
"""
import numpy as np

from . import edges as OE


def positive_triples(graph, batch_entities):
    """create_batch.py:267-273: for ent in batch, for each distinct 1-hop neighbour, for each parallel relation."""
    (trgts, srcs), vals = OE.batch_adj(graph, batch_entities)
    return [[s, r, t] for s, r, t in zip(srcs, vals, trgts)]


def get_iteration_triples_batch(graph, batch_entities, valid_triples, n_entities, n_relations, ratio, rng=np.random,
                                trace=None):
    """Returns (batch_indices int32 [T,3], batch_values float32 [T,1]). `valid_triples`: set of (h, r, t).
    trace (optional dict) receives the initial draws: trace['random_entities'], trace['random_relations']."""
    cur = np.array(positive_triples(graph, batch_entities))
    p = cur.shape[0]
    idx = np.zeros((p * (ratio * 2 + 1), 3)).astype(np.int32)       # np.empty in the reference; fully overwritten
    val = np.zeros((p * (ratio * 2 + 1), 1)).astype(np.float32)
    if p:
        idx[:p, :] = cur
    val[:p, :] = 1.0
    last = p
    if ratio > 0:
        random_entities = rng.randint(0, n_entities, last * ratio)                         # 293-294
        random_relations = rng.randint(0, n_relations, last * ratio)                       # 295-296
        if trace is not None:
            trace["random_entities"] = random_entities.copy()
            trace["random_relations"] = random_relations.copy()
        idx[last:last * (ratio * 2 + 1), :] = np.tile(idx[:last, :], (ratio * 2, 1))       # 298-301
        val[last:last * (ratio * 2 + 1), :] = np.tile(val[:last, :], (ratio * 2, 1))
        half = ratio // 2
        for i in range(last):
            for j in range(half):                                                          # 304-313
                c = i * half + j
                while (random_entities[c], idx[last + c, 1], idx[last + c, 2]) in valid_triples:
                    random_entities[c] = rng.randint(0, n_entities)
                idx[last + c, 0] = random_entities[c]
                val[last + c, :] = [-1]
            for j in range(half):                                                          # 315-326
                c = last * half + (i * half + j)
                while (idx[last + c, 0], idx[last + c, 1], random_entities[c]) in valid_triples:
                    random_entities[c] = rng.randint(0, n_entities)
                idx[last + c, 2] = random_entities[c]
                val[last + c, :] = [-1]
            for j in range(ratio):                                                         # 328-347
                c = last * ratio + (i * ratio + j)
                cr = i * ratio + j
                rel_count = 0
                while (idx[last + c, 0], random_relations[cr], idx[last + c, 2]) in valid_triples:
                    random_relations[cr] = rng.randint(0, n_relations)
                    rel_count += 1
                    if rel_count >= n_relations:
                        break
                if rel_count < n_relations:
                    idx[last + c, 1] = random_relations[cr]
                    val[last + c, :] = [-1]
    return idx, val
