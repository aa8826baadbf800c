"""The header-only C++ host layer (include/spim_mvdecon.hpp: LRFFT / LRInput / BayesMVDeconvolution,
MVDeconFFT / MVDeconInput / MVDeconvolution) compiled with g++ and run against the kernel emulator
(CPU) -- and against the CUDA library on a GPU box -- then compared with the oracle."""
import os
import subprocess

import numpy as np
import pytest

from oracle import mvdecon_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_and_run(lib_path, tmp_path, gen):
    exe = os.path.join(str(tmp_path), "host_mirror_test")
    libdir, libname = os.path.split(lib_path)
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "host_mirror_test.cpp"), "-o", exe,
           f"-L{libdir}", f"-l:{libname}", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = os.path.join(str(tmp_path), f"out{gen}.bin")
    r = subprocess.run([exe, out, str(gen)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "HOST_MIRROR_OK" in r.stdout, r.stdout + r.stderr
    raw = np.fromfile(out, dtype=np.uint8)
    X, Y, Z, K, V, g = np.frombuffer(raw[:24], dtype=np.int32)
    p = 24
    imgs, ws, psfs = [], [], []
    n, nk = X * Y * Z, K ** 3

    def take(count, shape):
        nonlocal p
        a = np.frombuffer(raw[p:p + 4 * count], dtype=np.float32).reshape(shape)
        p += 4 * count
        return a
    for _ in range(V):
        imgs.append(take(n, (Z, Y, X))); ws.append(take(n, (Z, Y, X))); psfs.append(take(nk, (K, K, K)))
    psi = take(n, (Z, Y, X)); k2 = take(nk, (K, K, K)); c1 = take(n, (Z, Y, X))
    avg = float(np.frombuffer(raw[p:p + 8], dtype=np.float64)[0])
    return imgs, ws, psfs, psi, k2, c1, avg


def check(lib_path, tmp_path, gen):
    imgs, ws, psfs, psi, k2, c1, avg = build_and_run(lib_path, tmp_path, gen)
    if gen == 2:
        ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=O.EFFICIENT_BAYESIAN, num_iterations=3, lam=0.006, gen=2))
    else:
        ref = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=O.OPTIMIZATION_I, num_iterations=3, lam=0.006, gen=1, osem_index=2))
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= 1e-3 and l2 <= 1e-4, (per, l2)
    assert np.isclose(avg, ref.avg, rtol=1e-6)
    assert np.abs(k2 - ref.kernel2[1]).max() <= 5e-5 * ref.kernel2[1].max()
    want = O.convolve(psi, ref.kernel1[1], O.EXT_MIRROR_SINGLE, dtype=np.float64)
    assert np.abs(c1 - want).max() / np.abs(want).max() < 2e-5


@pytest.mark.parametrize("gen", [1, 2])
def test_cpp_host_mirror_on_emulator(tmp_path, gen):
    import __graft_entry__ as g
    check(g.build_emulator(), tmp_path, gen)


@pytest.mark.gpu
@pytest.mark.parametrize("gen", [1, 2])
def test_cpp_host_mirror_on_gpu(gpu, tmp_path, gen):
    from spim_registration_b200 import native
    check(native.default_library_path(), tmp_path, gen)
