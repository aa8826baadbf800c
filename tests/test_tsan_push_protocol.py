"""The direct halo push (mvd_p2p_*) under ThreadSanitizer.  tests/cpp/p2p_tsan_driver.cpp runs the ranks as std::threads over
the C ABI of the kernel emulator (no Python in the loop: the interpreter lock's hand-overs would order the threads behind
ThreadSanitizer's back) with every access of the emulated kernels to their own and their neighbours' buffers instrumented.
The protocol's happens-before edges -- release store of the epoch flag after the pushed data, acquire load before the halo
is read, alternation of the two buffers before a halo is overwritten -- must leave no data race; the negative control, the
same program with relaxed flag accesses, must be reported.  This is the CPU stand-in for compute-sanitizer's racecheck,
which does not look across GPUs."""
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = [os.path.join(ROOT, "spim_registration_b200", "csrc", "spim_b200.cu"), os.path.join(ROOT, "tests", "cpp", "p2p_tsan_driver.cpp")]


def _build(out, extra):
    return subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-DSPIM_HOST_EMU"] + extra +
                          ["-x", "c++"] + SRC + ["-o", out], capture_output=True, text=True)


def _run(exe, grid):
    return subprocess.run([exe] + [str(g) for g in grid] + ["3"], capture_output=True, text=True, timeout=600,
                          env=dict(os.environ, TSAN_OPTIONS="exitcode=0 report_signal_unsafe=0"))


def test_push_protocol_is_race_free_under_tsan(tmp_path):
    good, bad = os.path.join(str(tmp_path), "p2p_tsan"), os.path.join(str(tmp_path), "p2p_tsan_relaxed")
    with ThreadPoolExecutor(2) as ex:
        rg, rb = ex.map(lambda a: _build(*a), [(good, []), (bad, ["-DSPIM_EMU_RELAXED_FLAGS"])])
    if rg.returncode != 0 and "tsan" in (rg.stderr or "").lower():
        pytest.skip("ThreadSanitizer runtime not available: " + rg.stderr.strip()[-200:])
    assert rg.returncode == 0 and rb.returncode == 0, (rg.stderr or "")[-2000:] + (rb.stderr or "")[-2000:]
    probe = _run(good, (1, 1, 2))
    if "unexpected memory mapping" in probe.stderr or "FATAL: ThreadSanitizer" in probe.stderr:
        pytest.skip("ThreadSanitizer cannot run in this environment: " + probe.stderr.strip()[-200:])
    for grid in ((1, 1, 2), (1, 2, 2), (2, 2, 2)):
        r = _run(good, grid)
        assert "P2P_DRIVER_OK" in r.stdout, r.stdout[-1000:] + r.stderr[-2000:]
        assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    r = _run(bad, (2, 2, 2))
    assert "WARNING: ThreadSanitizer: data race" in r.stderr, "the negative control (relaxed epoch flags) went undetected"
