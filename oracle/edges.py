"""CPU oracle for edge construction (stage 1) -- TEST INFRASTRUCTURE, NOT PRODUCT.

Pure-Python restatement of how the reference turns a triple list into the tensors the
attention layers consume:

  * triples_to_adj   <- preprocess.load_data            /root/reference/GAT/preprocess.py:48-87
                        Corpus.__init__                  /root/reference/GAT/create_batch.py:28-31
  * build_graph      <- Corpus.get_graph                 create_batch.py:708-732
  * bfs_exact_depth  <- Corpus.bfs                       create_batch.py:788-842
  * batch_adj        <- Corpus.get_batch_adj_data        create_batch.py:391-436
  * batch_nhop       <- Corpus.get_batch_nhop_neighbors_all  create_batch.py:871-895

Pinned against the reference's Corpus run in the build container
(tests/golden/make_golden.py -> tests/golden/edges_*.npz).
"""
from collections import deque


def triples_to_adj(triples, is_unweigted=False, directed=True):
    """triples: iterable of (head, rel, tail) ids in file order.
    Returns rows=[tail...], cols=[head...], data=[rel...] (preprocess.py:60-84); directed=False appends the
    reversed edge (rows=head, cols=tail) before each edge (preprocess.py:66-73), is_unweigted stores 1 as data."""
    rows, cols, data = [], [], []
    for h, r, t in triples:
        d = 1 if is_unweigted else int(r)
        if not directed:
            rows.append(int(h)); cols.append(int(t)); data.append(d)
        rows.append(int(t)); cols.append(int(h)); data.append(d)
    return rows, cols, data


def build_graph(rows, cols, data):
    """graph[source=col][target=row] = [rel, ...]; dicts keep insertion (= file) order and
    multi-edges are kept as a list (create_batch.py:717-729)."""
    graph = {}
    for t, s, r in zip(rows, cols, data):
        graph.setdefault(s, {}).setdefault(t, []).append(r)
    return graph


def bfs_exact_depth(graph, source, depth):
    """FIFO BFS from `source`; first discoverer is the parent; nodes beyond `depth` are neither
    visited nor enqueued (create_batch.py:803-820). Returns, in discovery order, for every node
    at distance == depth: (relation_lists_from_node_back_to_source, nodes_from_node_back) exactly
    like create_batch.py:822-842 (relations[0] = rels(parent->node), relations[-1] = rels(source->first))."""
    dist = {source: 0}
    parent = {source: None}
    q = deque([source])
    while q:
        top = q.popleft()
        for target, rels in graph.get(top, {}).items():
            if target in dist:
                continue
            d = dist[top] + 1
            if d > depth:
                continue
            dist[target] = d
            parent[target] = (top, rels)
            q.append(target)
    out = []
    for node, d in dist.items():          # insertion order == discovery order
        if d != depth:
            continue
        relations, entities, cur = [], [node], node
        while parent[cur] is not None:
            relations.append(parent[cur][1])
            entities.append(parent[cur][0])
            cur = parent[cur][0]
        out.append((tuple(tuple(r) for r in relations), tuple(entities[:-1])))
    return out


def neighbors_table(graph, depth):
    """get_further_neighbors (create_batch.py:844-869): only sources that are graph keys and
    have at least one node at the requested depth get an entry."""
    table = {}
    for source in graph.keys():
        found = bfs_exact_depth(graph, source, depth)
        if found:
            table[source] = found
    return table


def batch_adj(graph, batch_entities):
    """get_batch_adj_data (create_batch.py:413-436): for each batch entity (list order), each
    distinct out-neighbour (first-seen order), each parallel relation (file order) ->
    ([trgts, srcs], vals)."""
    trgts, srcs, vals = [], [], []
    for ent in batch_entities:
        for rel_lists, nodes in bfs_exact_depth(graph, ent, 1) if ent in graph else []:
            for rel in rel_lists[0]:
                trgts.append(nodes[0]); srcs.append(ent); vals.append(rel)
    return [trgts, srcs], vals


def batch_nhop(graph, batch_sources, partial_2hop=False):
    """get_batch_nhop_neighbors_all (create_batch.py:871-895): rows [source, r(source->mid)[0],
    r(mid->target)[0], target] for every target at BFS distance exactly 2."""
    rows = []
    for s in batch_sources:
        if s not in graph:
            continue
        for i, (rel_lists, nodes) in enumerate(bfs_exact_depth(graph, s, 2)):
            if partial_2hop and i >= 1:
                break
            rows.append([s, rel_lists[-1][0], rel_lists[0][0], nodes[0]])
    return rows
