"""Timing of the N3 sampler on one B200 (CUDA events + wall clock, one JSON line): a 2M-entity / 20M-triple KG,
batches of `--batch` entities, valid_invalid_ratio 2.  python profiles/bench_sampler.py [--batch 100000]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=100_000)
    ap.add_argument("--steps", type=int, default=5)
    args = ap.parse_args()
    from recon_b200.sampler import TripleSampler
    dev = torch.device("cuda:0")
    n, r, e = 2_000_000, 1000, 20_000_000
    g = torch.Generator(device=dev).manual_seed(0)
    tri = torch.stack((torch.randint(0, n, (e,), device=dev, generator=g), torch.randint(0, r, (e,), device=dev, generator=g),
                       torch.randint(0, n, (e,), device=dev, generator=g)), 1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    smp = TripleSampler(tri, n, r, invalid_valid_ratio=2, device=dev)
    torch.cuda.synchronize(); t_build = time.perf_counter() - t0
    batch = torch.randperm(n, device=dev, generator=g)[:args.batch]
    ts, tp = [], []
    for i in range(args.steps + 1):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        pos = smp.positive_triples(batch)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        idx, val = smp.get_iteration_triples_batch(batch, seed=i)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        if i:
            tp.append(t1 - t0); ts.append(t2 - t1)
    t = idx.shape[0]
    print(json.dumps({"workload": f"sampler: N={n} R={r} triples={e} batch_entities={args.batch} ratio=2 -> T={t}",
                      "build_s": t_build, "positives_ms": 1e3 * sum(tp) / len(tp), "full_batch_ms": 1e3 * sum(ts) / len(ts),
                      "triples_per_s": t / (sum(ts) / len(ts))}))


if __name__ == "__main__":
    main()
