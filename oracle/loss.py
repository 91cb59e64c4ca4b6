"""TEST INFRASTRUCTURE ONLY (checker for tests/, smoke() and bench.py's cpu_baseline leg) -- never imported by recon_b200.

CPU restatement of the training step that follows the hot path, in plain torch ops:
  * batch_gat_loss           /root/reference/GAT/main.py:344-376
  * optimizer step (SGD)     /root/reference/GAT/main.py:445-446, 524
Pinned against the reference's own function (its source text executed unmodified by tests/golden/make_golden.py,
fixtures tests/golden/loss_*.npz) in tests/test_oracle_golden.py.
"""
import torch


def batch_gat_loss(train_indices, entity_embed, relation_embed, ratio, margin):
    """main.py:344-376 with args.valid_invalid_ratio_gat = ratio and gat_loss_func = MarginRankingLoss(margin)."""
    t = train_indices.shape[0]
    n_pos = int(t / (int(ratio) * 2 + 1))                                  # main.py:345-346
    pos, neg = train_indices[:n_pos], train_indices[n_pos:]                # main.py:348-349
    pos = pos.repeat(int(ratio) * 2, 1)                                    # main.py:351

    def l1(tri):                                                           # main.py:353-358 / 360-365
        x = entity_embed[tri[:, 0]] + relation_embed[tri[:, 1]] - entity_embed[tri[:, 2]]
        return x.abs().sum(dim=1)

    pos_norm, neg_norm = l1(pos), l1(neg)
    y = -torch.ones(int(ratio * 2) * n_pos, dtype=pos_norm.dtype)         # main.py:367-370
    # nn.MarginRankingLoss: mean(max(0, -y * (x1 - x2) + margin))          # main.py:372, 451
    loss = torch.clamp_min(-y * (pos_norm - neg_norm) + margin, 0).mean()
    assert not torch.isnan(loss).any()                                     # main.py:374
    return loss


def loss_fwd_bwd(train_indices, entity_embed, relation_embed, ratio, margin):
    """Returns (loss, d loss / d entity_embed, d loss / d relation_embed)."""
    ent = entity_embed.detach().clone().requires_grad_(True)
    rel = relation_embed.detach().clone().requires_grad_(True)
    loss = batch_gat_loss(train_indices, ent, rel, ratio, margin)
    loss.backward()
    return loss.detach(), ent.grad, rel.grad


def sgd_step(params, grads, lr):
    """torch.optim.SGD(lr) without momentum / weight decay: p <- p - lr * g."""
    return [p - lr * g for p, g in zip(params, grads)]


def make_train_indices(n_ent, n_rel, n_pos, ratio, seed, hub_entity=None, hub_share=0.0):
    """Synthetic train_indices with the one property batch_gat_loss relies on (main.py:348-351): P positives first, then
    2*ratio*P corrupted rows where row P + k corrupts positive k mod P (here: the first ratio*P rows get a random head,
    the rest a random tail; the reference's sampler, oracle/sampler.py, also corrupts relations).
    hub_entity / hub_share force one entity into that share of the heads (long incidence segment)."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.stack((torch.randint(0, n_ent, (n_pos,), generator=g), torch.randint(0, n_rel, (n_pos,), generator=g),
                       torch.randint(0, n_ent, (n_pos,), generator=g)), dim=1)
    if hub_entity is not None and hub_share > 0:
        pick = torch.rand(n_pos, generator=g) < hub_share
        pos[pick, 0] = hub_entity
    neg = pos.repeat(2 * ratio, 1)
    half = ratio * n_pos
    neg[:half, 0] = torch.randint(0, n_ent, (half,), generator=g)
    neg[half:, 2] = torch.randint(0, n_ent, (half,), generator=g)
    return torch.cat((pos, neg), dim=0)
