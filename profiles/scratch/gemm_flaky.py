"""Debug: repeat the C += A B tensor-core GEMM on the flaky shape and describe the wrong elements."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from recon_b200.functional import gemm_nn
dev = torch.device("cuda:0")
m, k, n = 20000, 52, 416
g = torch.Generator().manual_seed(m + k + n)
a = torch.randn(m, k, generator=g); b = torch.randn(k, n, generator=g); c0 = torch.randn(m, n, generator=g)
ad, bd = a.to(dev), b.to(dev)
ref_ab = (a.double() @ b.double()).to(dev)
c0d = c0.to(dev)
torch.cuda.synchronize()
bad = 0
for it in range(int(sys.argv[1]) if len(sys.argv) > 1 else 40):
    acc = it % 2 == 1
    out = c0d.clone() if acc else torch.full((m, n), 7.0, device=dev)
    if it % 4 >= 2:
        torch.cuda.synchronize()
    gemm_nn(ad, bd, out=out, accumulate=acc)
    torch.cuda.synchronize()
    ref = ref_ab + (c0d.double() if acc else 0)
    err = (out.double() - ref).abs()
    wrong = err > 1e-3
    nw = int(wrong.sum())
    if nw:
        bad += 1
        rows = wrong.any(1).nonzero().flatten(); cols = wrong.any(0).nonzero().flatten()
        w = wrong.nonzero()[:1][0]
        r, c = int(w[0]), int(w[1])
        print(f"it {it} acc={acc} sync={it % 4 >= 2}: {nw} wrong; rows {int(rows[0])}..{int(rows[-1])} ({rows.numel()}), cols {int(cols[0])}..{int(cols[-1])} ({cols.numel()});"
              f" sample [{r},{c}] got {float(out[r, c]):.5f} ab {float(ref_ab[r, c]):.5f} c0 {float(c0d[r, c]):.5f}")
print("bad iterations:", bad)
