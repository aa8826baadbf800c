"""Barrier placement of the FFT-convolution kernels under ThreadSanitizer, without a GPU.  With SPIM_EMU_THREADS=T the kernel
emulator runs every block of every kernel as T real threads that split the work items like the
threads of a CUDA block and meet at real barriers (csrc/hd.h, csrc/runtime.h); tests/cpp/kernel_tsan_driver.cpp drives all
extension rules (also on 16-byte aligned rows: the TMA-fed x-forward kernel with its fix-up, de-duplication and split phases),
narrow column tiles, 8-line x-forward tiles, the full / literal forward sweeps and a deconvolution with the fused update
epilogue through the C ABI.  No data race may be reported and the threaded results must equal the single-thread ones
bit for bit; the negative control -- the same program with the barrier after a radix stage removed -- must be caught."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "spim_registration_b200", "csrc", "spim_b200.cu"), os.path.join(ROOT, "tests", "cpp", "kernel_tsan_driver.cpp")]


def _build(out, extra):
    return subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-DSPIM_HOST_EMU"] + extra +
                          ["-x", "c++"] + SRC + ["-o", out], capture_output=True, text=True)


def _run(exe):
    return subprocess.run([exe], capture_output=True, text=True, timeout=900,
                          env=dict(os.environ, TSAN_OPTIONS="exitcode=0 report_signal_unsafe=0"))


def test_kernel_barriers_are_race_free_under_tsan(tmp_path):
    good, bad = os.path.join(str(tmp_path), "k_tsan"), os.path.join(str(tmp_path), "k_tsan_no_barrier")
    with ThreadPoolExecutor(2) as ex:
        rg, rb = ex.map(lambda a: _build(*a), [(good, []), (bad, ["-DSPIM_EMU_NO_STAGE_BARRIER"])])
    if rg.returncode != 0 and "tsan" in (rg.stderr or "").lower():
        pytest.skip("ThreadSanitizer runtime not available: " + rg.stderr.strip()[-200:])
    assert rg.returncode == 0 and rb.returncode == 0, (rg.stderr or "")[-2000:] + (rb.stderr or "")[-2000:]
    r = _run(good)
    if "unexpected memory mapping" in r.stderr or "FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + r.stderr.strip()[-200:])
    assert "KERNEL_DRIVER_OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    n = _run(bad)
    assert "WARNING: ThreadSanitizer: data race" in n.stderr, "the negative control (no barrier after a radix stage) went undetected"


def test_threaded_blocks_give_the_same_bits(emu_lib, monkeypatch):
    """the same mode without the sanitizer, through the Python layer: 3 and 5 threads per block"""
    import numpy as np
    import parity_cases as P
    from oracle import mvdecon_oracle as O
    from spim_registration_b200 import synthetic
    shape = (14, 18, 22)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 5, kind="beads")
    a, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 2)
    for t in ("3", "5"):
        monkeypatch.setenv("SPIM_EMU_THREADS", t)
        b, *_ = P.run_session(emu_lib, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 2)
        assert np.array_equal(a, b)
        P.conv_case(emu_lib, (9, 7, 11), (3, 5, 3), 2)
