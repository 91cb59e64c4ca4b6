"""Per-launch breakdown of KGraph construction at the C2 shape (library timing hooks + torch profiler for the glue)."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from recon_b200 import KGraph
from recon_b200.synth import make_kg
dev = torch.device("cuda:0")
n, e, r = 2_000_000, 20_000_000, 1000
edge, etype, nhop = make_kg(n, e, r, alpha=1.1, seed=0, device=dev, hub_frac=0.2)
KGraph(edge, etype, None, n, r, device=dev); torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3):
        KGraph(edge, etype, None, n, r, device=dev)
    torch.cuda.synchronize()
rows = [(ev.key, ev.device_time_total / 3e3, ev.count // 3) for ev in prof.key_averages() if ev.device_time_total > 0]
rows.sort(key=lambda x: -x[1])
tot = 0
for k, ms, c in rows[:30]:
    print(f"{ms:8.3f} ms x{c:3d}  {k[:100]}")
    tot += ms
print("total device ms per build:", round(tot, 3))
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(3):
    KGraph(edge, etype, None, n, r, device=dev)
ev1.record(); torch.cuda.synchronize()
print("wall (events) ms per build:", ev0.elapsed_time(ev1) / 3)
