"""Multi-GPU parity (needs >= 2 GPUs on the box; skipped otherwise): launches
scripts/dist_check.py under torchrun, one rank per GPU, NCCL."""

import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2])
def test_distributed_assembly_and_cg(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "scripts", "dist_check.py"), "6", "10"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "DIST CHECK OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
