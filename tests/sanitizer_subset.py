"""A small parity subset for `compute-sanitizer --tool memcheck|racecheck|synccheck` on hardware (profiles/r2_sanitizer.sh):
every radix 2..10 once along each axis, all five out-of-bounds rules, both fused epilogues in both generations, the TMA-fed x
kernels and their plain fallbacks, brick-mode halo fill / pack / unpack on one GPU.  Small volumes: the sanitizer slows
kernels down by two to three orders of magnitude.   python tests/sanitizer_subset.py [quick]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np

import parity_cases as P
from oracle import mvdecon_oracle as O
from spim_registration_b200 import build, native, synthetic
from spim_registration_b200.deconvolution import Session


def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    if os.environ.get("SPIM_SUBSET_EMU") == "1":       # dry run of this script under the kernel emulator (no GPU here)
        import __graft_entry__ as g
        lib = native.load_library(g.build_emulator())
    else:
        lib = native.load_library(build.build_cuda_library())
    assert lib.getNumDevicesCUDA() > 0, "no CUDA device"
    n_cases = 0
    # radices 2..10 (and 11 / 13) along each axis: exact periodic mode keeps P = n
    for n in ((16, 20, 36, 42, 54, 70) if quick else (16, 18, 20, 24, 28, 36, 40, 42, 54, 60, 66, 70, 78, 80, 100)):
        for shape in ((n, 4, 8), (4, n, 8), (4, 4, n)):
            P.legacy_case(lib, shape, (3, 3, 3), seed=n)
            n_cases += 1
    for ext in range(5):
        P.conv_case(lib, (12, 20, 18), (3, 7, 5), ext)      # unaligned rows: plain x kernels
        P.conv_case(lib, (10, 12, 16), (5, 5, 5), ext)      # 16-byte aligned rows: TMA-fed x kernels
        n_cases += 2
    for gen in (1, 2):
        for typ in ((2,) if quick else (0, 1, 2, 3)):
            P.decon_case(lib, (14, 18, 24), 3, 5, typ, gen, 2)
            n_cases += 1
    P.decon_case(lib, (9, 11, 13), 2, 3, O.INDEPENDENT, 2, 2, lam=0.0, use_weights=False)
    P.exact_tikhonov_case(lib)
    P.cells_case(lib)
    n_cases += 3
    # brick mode on one GPU: halo fill by the out-of-bounds rule, pack / unpack round trip
    shape, V = (12, 16, 20), 2
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, 5)
    with Session(shape, V, 2, generation=2, haloed=True, lib=lib) as s:
        for v in range(V):
            s.set_view(v, imgs[v], ws[v], psfs[v])
        s.init()
        part = s.init_partials()
        s.set_avg(part[0] / part[1], 1.0)
        s.set_halo_mask(7, 7)
        for v in range(V):
            s.fill_halo(0, 7, 7)
            s.view_phase(v, 0)
            s.fill_halo(1, 7, 7)
            s.view_phase(v, 1)
        s.finish()
        brick = s.get_psi()
    plain, *_ = P.run_session(lib, imgs, ws, psfs, 2, 2, 1)
    assert np.abs(brick - plain).max() <= 1e-5 * np.abs(plain).max()
    n_cases += 1
    print(f"SANITIZER_SUBSET_OK {n_cases} cases")


if __name__ == "__main__":
    main()
