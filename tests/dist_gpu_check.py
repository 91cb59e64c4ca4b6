"""Multi-GPU parity check (run under torchrun with >= 2 GPUs; not collected by pytest):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dist_gpu_check.py
Every rank runs the row-partitioned SpKBGAT step; rank 0 also runs the single-GPU model on the whole graph and
checks that the gathered outputs and all gradients agree (<= 1e-5 rel-L2; SURVEY.md 8e asks <= 1e-6 .. 1e-4)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from recon_b200 import SpKBGATModified                       # noqa: E402
from recon_b200.dist import PartitionedKBGAT                 # noqa: E402
from recon_b200.synth import make_kg                         # noqa: E402
from oracle import ref_torch as O                            # noqa: E402  (parameter initialiser only)


def rel(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    n, e, r, f, d, h = 6000, 90000, 37, 50, 100, 2
    n_nhop = int(os.environ.get("DIST_CHECK_NHOP", "20000"))       # 0: no 2-hop edges -> the split-dot backward runs
    edge, etype, nhop = make_kg(n, e, r, alpha=1.1, n_nhop=n_nhop, seed=21)
    p = O.init_params(n, r, f, d, h, seed=21)
    gen = torch.Generator().manual_seed(22)
    g_ent, g_rel = torch.randn(n, d * h, generator=gen), torch.randn(r, d * h, generator=gen)

    pk = PartitionedKBGAT(n, r, edge, etype, nhop, f, d, h, 0.2, dev, state_dict=p)
    pk.set_loss_weights(g_ent, g_rel)
    out_e, out_r, _ = pk.step()
    torch.cuda.synchronize()
    # gather the partitioned pieces on rank 0
    lo, hi = pk.lo, pk.hi
    pieces = {"out_entity": out_e.detach(), "grad.entity_embeddings": pk.model.entity_embeddings.grad}
    gathered = {}
    for k, t in pieces.items():
        full = torch.zeros(n, t.shape[1], device=dev)
        full[lo:hi] = t
        dist.all_reduce(full)
        gathered[k] = full
    ok = True
    if rank == 0:
        m = SpKBGATModified(p["entity_embeddings"].clone(), p["relation_embeddings"].clone(), [d, 2 * d], [d, 2 * d],
                            0.0, 0.2, [h, h], None)
        m.load_state_dict(p)
        m = m.to(dev)
        oe, orl, _ = m(None, torch.arange(n), (edge, etype), nhop)
        ((oe * g_ent.to(dev)).sum() + (orl * g_rel.to(dev)).sum()).backward()
        errs = {"out_entity": rel(gathered["out_entity"], oe), "out_relation": rel(out_r, orl),
                "grad.entity_embeddings": rel(gathered["grad.entity_embeddings"], m.entity_embeddings.grad)}
        ref = dict(m.named_parameters())
        for k, prm in pk.model.named_parameters():
            if k not in ("entity_embeddings", "final_entity_embeddings", "final_relation_embeddings") and prm.grad is not None:
                errs["grad." + k] = rel(prm.grad, ref[k].grad)
        print("world", world, "bounds", pk.part.bounds, {k: f"{v:.2e}" for k, v in errs.items()})
        ok = max(errs.values()) < 1e-5 and len(errs) >= 10
        print("DIST PARITY", "OK" if ok else "FAILED")
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
