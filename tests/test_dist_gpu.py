"""Multi-GPU parity as a collected test (-m gpu): self-launches torch.distributed.run over every visible GPU (2, 4, 8)
and checks the row-partitioned path against the single-GPU path (tests/dist_gpu_check.py). Skips on a 1-GPU box; the
same check is printed by `bench.py --gpus N` in its JSON line (`parity`), so the driver's scaling run carries it too."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_partitioned_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    for attempt in range(2):
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
               os.path.join(ROOT, "tests", "dist_gpu_check.py")]
        p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
        # a launch that never reached the first comparison (rendezvous / port / NCCL start-up trouble right after the
        # previous world's job, seen once on an 8-GPU box) is retried once; a parity result is never retried
        if p.returncode == 0 or "DIST PARITY" in p.stdout:
            break
    assert p.returncode == 0 and "DIST PARITY OK" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])


def test_ops_run_on_the_tensors_device_not_the_current_one():
    """a model on cuda:1 while cuda:0 is the current device (kernel attribute caches are per device, every
    entry point switches to the tensors' device)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from recon_b200 import SpKBGATModified
    from recon_b200.synth import make_kg
    from helpers import rel_l2
    n, r = 300, 7
    edge, etype, _ = make_kg(n, 3000, r, alpha=1.1, seed=3)
    torch.manual_seed(0)
    ent, rel = torch.randn(n, 50), torch.randn(r, 50)
    outs = []
    for d in (0, 1):
        torch.cuda.set_device(0)
        torch.manual_seed(1)
        m = SpKBGATModified(ent.clone(), rel.clone(), [100, 200], [100, 200], 0.0, 0.2, [2, 2], None).to(f"cuda:{d}")
        o, _, _ = m(None, torch.arange(n), (edge, etype), None)
        o.sum().backward()
        outs.append((o.detach().cpu(), m.entity_embeddings.grad.cpu()))
    assert rel_l2(outs[1][0], outs[0][0]) < 1e-6 and rel_l2(outs[1][1], outs[0][1]) < 1e-5
