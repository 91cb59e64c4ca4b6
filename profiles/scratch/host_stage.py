import os, sys, time, torch
sys.path.insert(0, os.getcwd())
from recon_b200 import _lib
lib = _lib.load()
n = 20_000_000
src = torch.randint(0, 2_000_000, (n,), dtype=torch.int64).pin_memory()
dst = torch.empty(n, dtype=torch.int32).pin_memory()
dev = torch.empty(n, dtype=torch.int32, device="cuda")
dev64 = torch.empty(n, dtype=torch.int64, device="cuda")
for nt in (0, 16, 8, 4, 1):
    ts = []
    for _ in range(4):
        t = time.perf_counter(); lib.spk_pack_index_host(src.data_ptr(), n, 1, 0, 2_000_000, dst.data_ptr(), nt); ts.append(time.perf_counter() - t)
    print(f"pack 20M int64->int32, threads={nt}: {min(ts)*1e3:.2f} ms")
for t_, name in ((dst, "int32 80MB"), (src, "int64 160MB")):
    d = dev if t_ is dst else dev64
    torch.cuda.synchronize()
    ts = []
    for _ in range(4):
        t = time.perf_counter(); d.copy_(t_, non_blocking=True); torch.cuda.synchronize(); ts.append(time.perf_counter() - t)
    print(f"H2D {name}: {min(ts)*1e3:.2f} ms -> {t_.numel()*t_.element_size()/min(ts)/1e9:.1f} GB/s")
# numpy astype for comparison
import numpy as np
a = src.numpy()
t = time.perf_counter(); b = a.astype(np.int32); print("numpy astype 1 thread:", (time.perf_counter()-t)*1e3, "ms")
