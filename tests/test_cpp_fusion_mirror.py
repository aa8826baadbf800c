"""The header-only C++ host layer of the fusion pre-step (include/spim_fusion.hpp: AffineTransform3D, ExtractPSF,
ProcessForDeconvolution) compiled with g++ and run against the kernel emulator (CPU) -- and against the CUDA library on
a GPU box (test_zz_gpu_fusion.py) -- then recomputed with the oracle: transformed image / weights / PSFs bit-exact, psi at the parity bar."""
import os
import subprocess

import numpy as np
import pytest

from oracle import fusion_oracle as F
from oracle import mvdecon_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Reader:
    def __init__(self, path):
        self.raw = np.fromfile(path, dtype=np.uint8)
        self.p = 0

    def take(self, dtype, count):
        n = np.dtype(dtype).itemsize * count
        a = np.frombuffer(self.raw[self.p:self.p + n], dtype=dtype)
        self.p += n
        return a

    def image(self):
        x, y, z = self.take(np.int32, 3)
        return self.take(np.float32, x * y * z).reshape(z, y, x)


def check(lib_path, tmp_path):
    exe = os.path.join(str(tmp_path), "fusion_mirror_test")
    libdir, libname = os.path.split(lib_path)
    cmd = ["g++", "-std=c++17", "-O1", os.path.join(ROOT, "tests", "cpp", "fusion_mirror_test.cpp"), "-o", exe,
           f"-L{libdir}", f"-l:{libname}", f"-Wl,-rpath,{libdir}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = os.path.join(str(tmp_path), "fusion.bin")
    r = subprocess.run([exe, out], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "FUSION_MIRROR_OK" in r.stdout, r.stdout + r.stderr
    rd = Reader(out)
    V, NB, BX, BY = rd.take(np.int32, 4)
    BZ = int(rd.take(np.int32, 1)[0])
    stacks, models, beads = [], [], []
    for _ in range(V):
        stacks.append(rd.image())
        models.append(rd.take(np.float64, 12))
        beads.append(rd.take(np.float64, 3 * NB).reshape(NB, 3))
    img0, w1_raw, psf1, orig0 = rd.image(), rd.image(), rd.image(), rd.image()
    mn, avg, osem = rd.take(np.float64, 3)
    w1_final, psi = rd.image(), rd.image()

    dims, off, border, rng_ = (BZ, BY, BX), (-1, 0, 1), (1, 1, 0), (4, 4, 2)
    srcs = [F.loader_normalize(s) for s in stacks]
    invs = [F.invert_affine(m) for m in models]
    imgs, ws = zip(*[F.transform_input_and_weights(srcs[v], invs[v], dims, off, border, rng_) for v in range(V)])
    np.testing.assert_array_equal(img0, imgs[0])
    np.testing.assert_array_equal(w1_raw, ws[1])
    psfs, origs = zip(*[F.extract_next_img(srcs[v], models[v], beads[v], (5, 5, 3)) for v in range(V)])
    np.testing.assert_array_equal(psf1, psfs[1])
    np.testing.assert_array_equal(orig0, origs[0])
    sumw, omn, oavg = F.weight_normalizer_virtual(ws, 4)
    assert mn == max(1, omn) and avg == max(1.0, oavg) and osem == avg
    wv = [F.normalizing_access(w, sumw, osem) for w in ws]
    np.testing.assert_array_equal(w1_final, wv[1])
    ref = O.deconvolve(list(imgs), wv, list(psfs), O.DeconParams(iteration_type=O.INDEPENDENT, num_iterations=2, lam=0.006, gen=2))
    per, l2 = O.parity_errors(psi, ref.psi)
    assert per <= 1e-3 and l2 <= 1e-4, (per, l2)


def test_cpp_fusion_mirror_on_emulator(tmp_path):
    import __graft_entry__ as g
    check(g.build_emulator(), tmp_path)
