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


@pytest.mark.gpu
@pytest.mark.xfail(reason="opt-in path written after the round's GPU budget was spent: verified under the emulator "
                          "(tests/test_bricks_p2p_threads.py), not yet run on hardware", strict=False)
def test_bricks_direct_push_over_peer_memory(gpu):
    """SPIM_BRICK_P2P=1: the halo exchange as one fused copy + signal kernel over CUDA-IPC peer memory (mvd_p2p_*), with
    the statistics-free iterations captured into a CUDA graph; the run must really have adopted the push path."""
    n = gpu.getNumDevicesCUDA()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, SPIM_BRICK_P2P="1", SPIM_TEST_EXTRA_ITERS="2", SPIM_P2P_TIMEOUT_S="10")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "run_bricks_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert "BRICKS_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "EXCHANGE_PATH p2p" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
