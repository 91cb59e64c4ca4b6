"""Timing of the N1 kernels (batch_gat_loss forward / backward, SGD step) on one B200: CUDA events, inputs resident.
Workload: the hot path's C2 output tables (2M x 200 entity rows, 1k x 200 relation rows), P positives with
valid_invalid_ratio_gat = 2 (main.py:68-69) -> T = 5 P triples. Prints one JSON line.
  python profiles/bench_loss.py [--pos 2000000] [--steps 10]
Algorithmic bytes: forward T * (2 * 4 * width) entity-row gathers (the R x width relation table is
cache-resident) + T * (24 + width + 8) (ids in, packed signs + norm / coef out); backward 3 T incidences * (width + 8)
+ (N + R) * 4 * width dense gradient rows."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pos", type=int, default=2_000_000)
    ap.add_argument("--steps", type=int, default=10)
    args = ap.parse_args()
    from recon_b200.loss import batch_gat_loss, sgd_step
    dev = torch.device("cuda:0")
    n, r, w, ratio = 2_000_000, 1000, 200, 2
    g = torch.Generator(device=dev).manual_seed(0)
    ent = torch.nn.functional.normalize(torch.randn(n, w, device=dev, generator=g), dim=1).requires_grad_(True)
    rel = torch.randn(r, w, device=dev, generator=g).requires_grad_(True)
    p = args.pos
    pos = torch.stack((torch.randint(0, n, (p,), device=dev, generator=g), torch.randint(0, r, (p,), device=dev, generator=g),
                       torch.randint(0, n, (p,), device=dev, generator=g)), 1)
    neg = pos.repeat(2 * ratio, 1)
    neg[: ratio * p, 0] = torch.randint(0, n, (ratio * p,), device=dev, generator=g)
    neg[ratio * p:, 2] = torch.randint(0, n, (ratio * p,), device=dev, generator=g)
    tri = torch.cat((pos, neg), 0)
    t = tri.shape[0]
    f = torch.nn.MarginRankingLoss(margin=5.0)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    tf = tb = ts = 0.0
    for i in range(args.steps + 3):
        ent.grad = rel.grad = None
        ev[0].record()
        loss = batch_gat_loss(f, tri, ent, rel, valid_invalid_ratio_gat=ratio)
        ev[1].record()
        loss.backward()
        ev[2].record()
        sgd_step([ent, rel], 1e-3)
        ev[3].record()
        torch.cuda.synchronize()
        if i >= 3:
            tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2]); ts += ev[2].elapsed_time(ev[3])
    k = args.steps
    fwd_b = t * (2 * 4 * w) + t * (24 + w + 8)
    bwd_b = 3 * t * (w + 8) + (n + r) * 4 * w
    sgd_b = (n + r) * w * 12
    print(json.dumps({"workload": f"loss: N={n} R={r} width={w} P={p} T={t}", "loss": float(loss),
                      "fwd_ms": tf / k, "bwd_ms_incl_incidence_sort": tb / k, "sgd_ms": ts / k,
                      "triples_per_s_fwd_bwd": t / ((tf + tb) / k * 1e-3),
                      "fwd_gbs": fwd_b / (tf / k * 1e-3) / 1e9, "bwd_gbs": bwd_b / (tb / k * 1e-3) / 1e9,
                      "sgd_gbs": sgd_b / (ts / k * 1e-3) / 1e9}))


if __name__ == "__main__":
    main()
