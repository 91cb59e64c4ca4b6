"""CPU oracle for the ConvKB scoring stage -- TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, op for op (PyTorch CPU, fp32 or fp64):
  * convkb_forward          <- ConvKB.forward                  /root/reference/GAT/layers.py:31-48 (the live path is the
                                                                two-layer MLP fc2(LeakyReLU(fc1(x))); the convolution
                                                                is commented out in the reference)
  * conv_only_forward       <- SpKBGATConvOnly.forward/.batch_test  /root/reference/GAT/models.py:294-304
  * ent2rel_project         <- tanh(bmm(e, W_ent2rel[r]))      /root/reference/GAT_sep_space/models.py:316-320
  * conv_only_forward_sep   <- SpKBGATConvOnly.forward/.batch_test  /root/reference/GAT_sep_space/models.py:311-339
  * relation_scores / relation_ranking <- Corpus.get_validation_cnfmat  /root/reference/GAT/create_batch.py:1361-1500
Pinned by tests/golden/convkb_*.npz, which tests/golden/make_golden.py writes by running the reference's own modules.
Only tests/ may import this module.
"""
import torch
import torch.nn.functional as F


def convkb_forward(x, p, prefix="convKB."):
    """x [B, 3*D] -> [B, 1]; p holds fc1.weight [D, 3D], fc1.bias [D], fc2.weight [1, D], fc2.bias [1]."""
    h1 = F.leaky_relu(F.linear(x, p[prefix + "fc1.weight"], p[prefix + "fc1.bias"]))          # layers.py:44 (slope 0.01)
    return F.linear(h1, p[prefix + "fc2.weight"], p[prefix + "fc2.bias"])                       # layers.py:45


def conv_only_forward(p, batch_inputs):
    """GAT/models.py:294-298: conv_input = [E[h] | Rel[r] | E[t]]."""
    ent, rel = p["final_entity_embeddings"], p["final_relation_embeddings"]
    x = torch.cat((ent[batch_inputs[:, 0]], rel[batch_inputs[:, 1]], ent[batch_inputs[:, 2]]), dim=1)
    return convkb_forward(x, p)


def ent2rel_project(rows, w_ent2rel, rel_ids):
    """GAT_sep_space/models.py:316-320: tanh(row . W_ent2rel[r]) per triple."""
    return torch.tanh(torch.bmm(rows.unsqueeze(1), w_ent2rel[rel_ids]).squeeze(1))


def conv_only_forward_sep(p, batch_inputs, w_ent2rel):
    """GAT_sep_space/models.py:311-324."""
    ent, rel = p["final_entity_embeddings"], p["final_relation_embeddings"]
    r = batch_inputs[:, 1]
    src = ent2rel_project(ent[batch_inputs[:, 0]], w_ent2rel, r)
    tail = ent2rel_project(ent[batch_inputs[:, 2]], w_ent2rel, r)
    return convkb_forward(torch.cat((src, rel[r], tail), dim=1), p)


def relation_scores(p, test_triples, num_rels, w_ent2rel=None):
    """create_batch.py:1367-1393: every test triple scored under every relation id -> [T, num_rels]."""
    t = test_triples.shape[0]
    tb = test_triples.unsqueeze(1).repeat(1, num_rels, 1)
    tb[:, :, 1] = torch.arange(num_rels).unsqueeze(0)
    tb = tb.reshape(-1, 3)
    s = conv_only_forward(p, tb) if w_ent2rel is None else conv_only_forward_sep(p, tb, w_ent2rel)
    return s.view(t, num_rels)


def relation_ranking(scores, test_triples):
    """create_batch.py:1413-1495, including its quirks: predictions per entity pair come from the FIRST test row of the
    pair, the top max(#actual relations, 10) ids; the rank lookup uses `sorted_indices[i]` with `i` left over from the
    previous loop, i.e. the LAST test row, for every pair. Returns (sorted_indices, metrics dict)."""
    sorted_scores, sorted_indices = torch.sort(scores, dim=-1, descending=True)
    tb = test_triples.tolist()
    actual, preds = {}, {}
    for e1, r, e2 in tb:
        actual.setdefault((e1, e2), set()).add(r)
    i = -1
    for i in range(sorted_scores.shape[0]):
        k = (tb[i][0], tb[i][2])
        if k not in preds:
            preds[k] = set(sorted_indices[i][:max(len(actual[k]), 10)].tolist())
    hits, ranks, rr = 0, [], []
    last = sorted_indices[i].tolist() if i >= 0 else []
    for k, rels in actual.items():
        hits += len(rels & preds[k])
        for rel in rels:
            if rel in last:
                ranks.append(last.index(rel) + 1)
                rr.append(1.0 / ranks[-1])
    n = max(1, len(ranks))
    return sorted_indices, {"hits_at_10": hits, "average_hits_at_10": hits / n, "average_rank": sum(ranks) / n,
                            "average_recip_rank": sum(rr) / max(1, len(rr))}
