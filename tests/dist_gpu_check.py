"""Multi-GPU parity worker (launched by tests/test_dist_gpu.py under torch.distributed.run with >= 2 GPUs, or by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Every rank runs the row-partitioned SpKBGAT step; rank 0 also runs the single-GPU model on the whole graph and checks
that the gathered outputs and all gradients agree (recon_b200.dist.parity_check; SURVEY.md 8e asks <= 1e-6 .. 1e-4).
DIST_CHECK_CASES selects the cases: "plain" (split-dot backward), "nhop" (2-hop edges), "hub" (one row holding 60 % of
the edges, which exceeds every rank's edge budget)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recon_b200.dist import parity_check                    # noqa: E402

TOL = 1e-5


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank = dist.get_rank()
    cases = os.environ.get("DIST_CHECK_CASES", "plain,nhop,hub").split(",")
    ok = True
    for case in cases:
        kw = dict(n=6000, e=90000, r=37, seed=21)
        if case == "nhop":
            kw["n_nhop"] = 20000
        elif case == "hub":
            kw.update(zipf=0.6, seed=23, hub_frac=1.0)  # every row Pareto(0.6): the top row holds a third of the edges and
                                                       # some ranks own no rows at all
        res = parity_check(dev, **kw)
        if rank == 0:
            print("DIST PARITY", case, res, flush=True)
        ok = ok and res["max_rel"] < TOL and res["tensors"] >= 10
    if rank == 0:
        print("DIST PARITY", "OK" if ok else "FAILED", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
