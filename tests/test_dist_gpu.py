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
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "dist_gpu_check.py")]
    p = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "DIST PARITY OK" in p.stdout, (p.stdout[-3000:], p.stderr[-3000:])
