"""Timeline of one C2 step from torch.profiler (no nsys in the image): GPU busy time vs span, the launches that do not
come from libspkbgat (ATen / cuBLAS glue) with the op that issued them, and the largest idle gaps between kernels.
Usage: python profiles/trace_gaps.py [out.json]   (run on the GPU box; numbers under the profiler are NOT bench values)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/trace_gaps.json"
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    bench.add_workload("c2")
    n, r, edge, etype, nhop = bench.make_inputs("c2")
    job = bench.SingleGPU(n, r, edge, etype, nhop, dev)
    for _ in range(3):
        job.step()
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2):
            job.step()
        torch.cuda.synchronize()
    tmp = "/tmp/trace.json"
    prof.export_chrome_trace(tmp)
    ev = json.load(open(tmp))["traceEvents"]
    ks = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    ks.sort(key=lambda e: e["ts"])
    # correlate kernels with the CPU op that launched them (External id -> cpu_op name)
    ops = {}
    for e in ev:
        if e.get("cat") == "cpu_op" and "args" in e and "External id" in e["args"]:
            ops[e["args"]["External id"]] = e["name"]
    half = len(ks) // 2
    ks = ks[half:]                                   # second profiled step
    span = ks[-1]["ts"] + ks[-1]["dur"] - ks[0]["ts"]
    busy = sum(e["dur"] for e in ks)
    glue = {}
    for e in ks:
        nm = e["name"]
        if "spk::" in nm:
            continue
        key = (nm[:70], ops.get(e.get("args", {}).get("External id"), "?"))
        g = glue.setdefault(key, [0, 0.0])
        g[0] += 1
        g[1] += e["dur"]
    gaps = []
    for a, b in zip(ks[:-1], ks[1:]):
        gap = b["ts"] - (a["ts"] + a["dur"])
        if gap > 15:
            gaps.append((gap, a["name"][:50], b["name"][:50]))
    gaps.sort(reverse=True)
    res = {"span_ms": span / 1e3, "busy_ms": busy / 1e3, "idle_ms": (span - busy) / 1e3, "kernels": len(ks),
           "glue_ms": sum(v[1] for v in glue.values()) / 1e3,
           "glue": [{"kernel": k[0], "op": k[1], "n": v[0], "ms": v[1] / 1e3} for k, v in sorted(glue.items(), key=lambda kv: -kv[1][1])[:40]],
           "gaps_over_15us": len(gaps), "gap_ms_over_15us": sum(g[0] for g in gaps) / 1e3,
           "top_gaps": [{"us": g[0], "after": g[1], "before": g[2]} for g in gaps[:40]]}
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: res[k] for k in ("span_ms", "busy_ms", "idle_ms", "kernels", "glue_ms", "gaps_over_15us", "gap_ms_over_15us")}))


if __name__ == "__main__":
    main()
