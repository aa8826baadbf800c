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
    # the NCCL fallback path (pack / unpack kernels around one batch of send / recv): the direct push is switched off
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, SPIM_BRICK_P2P="0", SPIM_TEST_EXTRA_ITERS="2"))
    print(r.stdout[-600:])
    assert "BRICKS_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "EXCHANGE_PATH pack" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_bricks_direct_push_over_peer_memory(gpu):
    """The default exchange: one fused copy + signal kernel over CUDA-IPC peer memory (mvd_p2p_*), with the
    statistics-free iterations captured into a CUDA graph; the run must really have adopted the push path."""
    n = gpu.getNumDevicesCUDA()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ, SPIM_BRICK_P2P="1", SPIM_TEST_EXTRA_ITERS="2", SPIM_P2P_TIMEOUT_S="10")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29534", os.path.join(ROOT, "tests", "run_bricks_nccl.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    print(r.stdout[-600:])
    assert "BRICKS_NCCL_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    assert "EXCHANGE_PATH p2p" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_bricks_in_one_process_one_thread_per_device(gpu, monkeypatch):
    """The reference's multi-device mode (one host thread per device, MVDeconFFT.java:447-469) with persistent bricks: the
    ranks are threads of THIS process (spim_registration_b200/inprocess.py), peers are reached through raw device pointers
    with peer access enabled, the halo exchange is the fused push + wait kernels alone."""
    import threading
    import numpy as np
    import torch
    from oracle import mvdecon_oracle as O
    from spim_registration_b200 import bricks, synthetic
    from spim_registration_b200.inprocess import ThreadGroup
    n = gpu.getNumDevicesCUDA()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    monkeypatch.setenv("SPIM_BRICK_P2P", "1")
    monkeypatch.setenv("SPIM_BRICK_GRAPH", "0")
    brick, V, ks, iters = (24, 28, 32), 3, 7, 2
    typ, gen = O.EFFICIENT_BAYESIAN, 2
    grid = bricks.grid_for(world)
    gshape = tuple(brick[d] * grid[d] for d in range(3))
    _, imgs, ws, psfs = synthetic.make_dataset(gshape, V, ks, kind="beads", seed=3)
    group = ThreadGroup(world)
    out, errors = {}, []

    def rank_main(r):
        try:
            torch.cuda.set_device(r)
            c = bricks.rank_coords(r, grid)
            sl = tuple(slice(c[d] * brick[d], (c[d] + 1) * brick[d]) for d in range(3))
            run = bricks.BrickRunner(brick, V, typ, generation=gen, lam=0.006, device=r, rank=r, world=world, grid=grid,
                                     dist=group.rank(r))
            for v in range(V):
                run.session.set_view(v, np.ascontiguousarray(imgs[v][sl]), np.ascontiguousarray(ws[v][sl]), psfs[v])
            run.init()
            used = run.use_p2p
            run.run(iters, stats=True)
            run.finish()
            out[r] = (sl, run.get_psi(), used)
            run.close()
        except BaseException as e:      # noqa: BLE001
            errors.append((r, repr(e)))
            group.abort()

    threads = [threading.Thread(target=rank_main, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=500)
    assert not errors, errors
    assert len(out) == world and all(u for _, _, u in out.values())
    psi = np.zeros(gshape, np.float32)
    for sl, p, _ in out.values():
        psi[sl] = p
    ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=typ, num_iterations=iters, lam=0.006, gen=gen))
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= 1e-3 and l2 <= 1e-4, (per, l2)


def _fullsize(args, world, timeout):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tests", "run_bricks_fullsize.py")] + args
    if world == 1:
        cmd = [sys.executable, os.path.join(ROOT, "tests", "run_bricks_fullsize.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    print(r.stdout[-1500:])
    assert "FULLSIZE_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_fullsize_configs1_one_full_iteration_against_the_oracle(gpu):
    """BASELINE configs[1] at full size on one GPU (7 views, 512x512x256, 31^3 PSFs, Efficient-Bayesian): one full iteration,
    then sub-bricks at the volume centre and at the volume corner are recomputed by the oracle from crops of the inputs
    (dependency cone of 7 view-steps included)."""
    _fullsize(["--config", "custom", "--brick", "256", "512", "512", "--views", "7", "--iter-type", "2"], 1, 500)


@pytest.mark.gpu
@pytest.mark.timeout(900)
def test_fullsize_configs2_bricks_on_all_gpus(gpu):
    """BASELINE configs[2] geometry: bricks of 512x512x256 on every GPU of the box (8 -> the 1024x1024x512 volume), one full
    iteration, sub-bricks straddling brick faces / the corner of all bricks / the volume corner against the oracle."""
    n = gpu.getNumDevicesCUDA()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _fullsize(["--config", "c3"], 8 if n >= 8 else (4 if n >= 4 else 2), 800)


@pytest.mark.gpu
@pytest.mark.timeout(1200)
def test_configs4_on_8_gpus(gpu):
    """BASELINE configs[4]: 8 views, 2048x2048x1024 in bricks of 1024x1024x512 on 8 GPUs (89 GB of HBM each), views handed over
    cell by cell with mvd_upload_region; the first view-steps checked against the oracle, then 10 iterations."""
    if gpu.getNumDevicesCUDA() < 8:
        pytest.skip("needs 8 GPUs")
    _fullsize(["--config", "c5"], 8, 1100)
