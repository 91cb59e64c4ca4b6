"""K0b throughput: batch adjacency + 2-hop rows of a 200k-entity KG on the device
(recon_b200.nhop.TripleGraph, radix sorts of libspkbgat + integer glue) against the reference algorithm (Corpus.get_graph +
Corpus.bfs per source, GAT/create_batch.py:708-895, restated in oracle/edges.py and pinned to the reference's Corpus by
tests/golden/edges_*.npz) on a bounded sample of the same sources. Prints one JSON line.

  python profiles/bench_nhop.py [--entities 200000] [--triples 2000000] [--relations 500] [--cpu-sample 2000]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--entities", type=int, default=200_000)
    ap.add_argument("--triples", type=int, default=2_000_000)
    ap.add_argument("--relations", type=int, default=500)
    ap.add_argument("--cpu-sample", type=int, default=2000)
    ap.add_argument("--chunk", type=int, default=50_000)
    args = ap.parse_args()
    from recon_b200.nhop import TripleGraph
    from recon_b200.synth import make_triples
    from oracle import edges as OE                      # checker / CPU baseline only
    dev = torch.device("cuda:0")
    n, t, r = args.entities, args.triples, args.relations
    tr = make_triples(n, t, r, seed=0)
    d_tr = tr.to(dev)

    def sync():
        torch.cuda.synchronize()

    TripleGraph(d_tr[:1000], n, device=dev)              # warm-up (library load, allocator)
    sync(); t0 = time.perf_counter()
    tg = TripleGraph(d_tr, n, device=dev)
    sync(); t_build = time.perf_counter() - t0
    sources = torch.arange(n, device=dev)
    tg.batch_edges(sources[:1000])
    sync(); t0 = time.perf_counter()
    e1 = e2 = 0
    per_chunk = []
    for rep in range(2):                                 # second sweep: allocator warm (what a training loop sees)
        e1 = e2 = 0
        sync(); t0 = time.perf_counter()
        for lo in range(0, n, args.chunk):
            tc = time.perf_counter()
            idx, val, nhop = tg.batch_edges(sources[lo:lo + args.chunk])
            e1 += idx.shape[1]; e2 += nhop.shape[0]
            sync(); per_chunk.append(round(time.perf_counter() - tc, 4))
        sync(); t_gpu = time.perf_counter() - t0
    for rep in range(2):
        sync(); t0 = time.perf_counter()
        for lo in range(0, n, args.chunk):
            tg.batch_edges_torch(sources[lo:lo + args.chunk])
        sync(); t_torch = time.perf_counter() - t0

    # the reference algorithm on the host: dict graph + one BFS per source, on a sample of the sources
    t0 = time.perf_counter()
    graph = OE.build_graph(*OE.triples_to_adj(tr.tolist()))
    t_graph_cpu = time.perf_counter() - t0
    gen = torch.Generator().manual_seed(1)
    sample = torch.randperm(n, generator=gen)[:args.cpu_sample].tolist()
    t0 = time.perf_counter()
    want_idx, want_val = OE.batch_adj(graph, sample)
    want_nhop = OE.batch_nhop(graph, sample)
    t_cpu = time.perf_counter() - t0
    idx, val, nhop = tg.batch_edges(sample)
    ok = idx.cpu().tolist() == want_idx and val.cpu().tolist() == want_val and nhop.cpu().tolist() == want_nhop
    print(json.dumps({
        "what": "K0b batch adjacency + 2-hop rows, all sources of the KG",
        "kg": {"entities": n, "triples": t, "relations": r}, "rows": {"one_hop": e1, "two_hop": e2},
        "gpu": {"triple_graph_build_s": t_build, "all_sources_s": t_gpu, "sources_per_s": n / t_gpu,
                "rows_per_s": (e1 + e2) / t_gpu, "per_chunk_s_two_sweeps": per_chunk,
                "round1_torch_glue_form_s": t_torch},
        "cpu_reference_algorithm": {"get_graph_s": t_graph_cpu, "sample_sources": len(sample), "sample_s": t_cpu,
                                    "sources_per_s": len(sample) / t_cpu, "cores": 1,
                                    "extrapolated_all_sources_s": t_cpu / len(sample) * n},
        "speedup_sources_per_s": (n / t_gpu) / (len(sample) / t_cpu),
        "sample_bit_exact": bool(ok)}))


if __name__ == "__main__":
    main()
