"""Kernel micro-benchmarks at the C2 shapes (one B200): every GEMM shape of the step and the layer-2 edge passes, each
timed alone with CUDA events (burst numbers: compare with MEASURED_PEAKS.json's burst HBM figure). For fast iteration on one
kernel without paying for a whole bench.py run; prints one JSON line per kernel.

  python profiles/bench_kernels.py [--what gemm,edge] [--reps 10] [--rows 2000000] [--edges 20000000]

GEMM bytes = 4 (M K + K N + M N) (operands read once, result written once); edge-pass bytes as bench.py's roofline().
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def timed(fn, reps):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="gemm,edge")
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--edges", type=int, default=20_000_000)
    args = ap.parse_args()
    from recon_b200 import functional as SF, KGraph
    from recon_b200.synth import make_kg
    dev = torch.device("cuda:0")
    peak = 6552.3
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = json.load(open(pk)).get("hbm_gbs", peak)
    m = args.rows
    g = torch.Generator(device=dev).manual_seed(0)

    if "gemm" in args.what:
        # (kind, K or Ka, N or Nb): the products of one C2 step (kernels_ms_per_step of bench.py)
        shapes = [("nn", 200, 416), ("nn", 416, 200), ("nn", 100, 156), ("nn", 156, 100), ("nn", 50, 200), ("nn", 200, 50),
                  ("tn", 200, 416), ("tn", 156, 100), ("tn", 50, 200), ("tn", 50, 4)]
        for kind, k, n in shapes:
            if kind == "nn":
                a = SF.tc_friendly(torch.randn(m, k, device=dev, generator=g))
                b = torch.randn(k, n, device=dev, generator=g)
                out = torch.empty(m, (n + 3) // 4 * 4, device=dev)[:, :n]
                ms = timed(lambda: SF.gemm_nn(a, b, out=out), args.reps)
            else:
                a = SF.tc_friendly(torch.randn(m, k, device=dev, generator=g))
                b = SF.tc_friendly(torch.randn(m, n, device=dev, generator=g))
                out = torch.empty(k, n, device=dev)
                ms = timed(lambda: SF.gemm_tn(a, b, out=out), args.reps)
            byt = 4.0 * (m * k + k * n + m * n)
            print(json.dumps({"kernel": f"gemm_{kind}:{m}x{k}x{n}", "ms": round(ms, 4), "gbs": round(byt / ms / 1e6, 1),
                              "frac_hbm": round(byt / ms / 1e6 / peak, 3),
                              "tflops_fp32_equiv": round(2.0 * m * k * n / ms / 1e9, 1),
                              "ms_at_hbm_peak": round(byt / peak / 1e6, 4)}))
            del a, b, out

    if "edge" in args.what:
        n, e, r = m, args.edges, 1000
        edge, etype, nhop = make_kg(n, e, r, 1.1, 0, 0, device=dev, hub_frac=0.2)
        graph = KGraph(edge, etype, None, n, r, device=dev)
        geom = SF.Geometry(1, 200)                      # layer 2: one head of 200 (Wd = 208)
        wd, dt = geom.Wd, 200
        P = torch.randn(n, 2 * wd, device=dev, generator=g) * 0.1
        P3 = torch.randn(r, wd, device=dev, generator=g) * 0.1
        nan = torch.zeros(1, dtype=torch.int32, device=dev)
        P1, P2 = P[:, :wd], P[:, wd:]
        res = {}

        def fwd():
            res["o"] = SF.edge_attn_forward(graph, P1, P2, P3, geom, 0.2, True, None, nan)
        ms = timed(fwd, args.reps)
        byt = e * (4 * dt + 8) + n * (8 * dt)
        print(json.dumps({"kernel": "edge_attn_fwd (layer 2)", "ms": round(ms, 4), "gbs": round(byt / ms / 1e6, 1),
                          "frac_hbm": round(byt / ms / 1e6 / peak, 3)}))
        out, den, sw = res["o"]
        dout = torch.randn(n, dt, device=dev, generator=g)
        dP = torch.empty(n, 2 * wd, device=dev)
        dP3 = torch.empty(r, wd, device=dev)
        for mode in ("split", "fused", "rows"):
            SF.BWD_MODE = mode
            ms = timed(lambda: SF.edge_attn_backward(graph, P1, P2, P3, geom, 0.2, True, None, out, dout, den, sw,
                                                     dP[:, :wd], dP[:, wd:], dP3), args.reps)
            byt = e * (8 * dt + 16 + 16) + n * (24 * dt)
            print(json.dumps({"kernel": f"edge_attn_backward[{mode}] (layer 2)", "ms": round(ms, 4),
                              "gbs": round(byt / ms / 1e6, 1), "frac_hbm": round(byt / ms / 1e6 / peak, 3)}))
        SF.BWD_MODE = "split"


if __name__ == "__main__":
    main()
