"""Real multi-GPU run of the brick partition (NCCL halo exchange).  Skipped on boxes with < 2 GPUs;
the same logic is covered on CPU by tests/test_bricks_gloo.py."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_bricks_over_nccl(gpu):
    n = gpu.getNumDevicesCUDA()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "run_bricks_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert "BRICKS_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
