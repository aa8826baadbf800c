"""Generates tests/golden/*.npz.

The reference repository contains no golden vectors, known-answer tests or fixtures for the
deconvolution path, and it cannot be executed here (Java, no JVM).  These fixtures
are therefore outputs of the oracle itself (oracle/mvdecon_oracle.py) in its fp64 'truth' mode on
small seeded inputs: they pin the oracle against accidental change and give the CUDA path a fixed
target, but they do NOT pin it against the Java implementation ("parity unpinned").

Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import mvdecon_oracle as O          # noqa: E402
from spim_registration_b200 import synthetic    # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    shape, V, ks = (16, 18, 20), 3, 5
    _, imgs, ws, psfs = synthetic.make_dataset(shape, V, ks, kind="beads", seed=11)
    out = {"shape": np.array(shape), "num_views": np.array(V)}
    for v in range(V):
        out[f"img{v}"] = imgs[v]
        out[f"w{v}"] = ws[v]
        out[f"psf{v}"] = psfs[v]
    for gen in (1, 2):
        for typ in range(4):
            p = O.DeconParams(iteration_type=typ, num_iterations=2, lam=0.006, gen=gen, dtype=np.float64)
            r = O.deconvolve(imgs, ws, psfs, p)
            out[f"psi_g{gen}_t{typ}"] = r.psi.astype(np.float32)
            out[f"avg_g{gen}_t{typ}"] = np.array(r.avg)
            if gen == 2:
                for v in range(V):
                    out[f"k2_t{typ}_v{v}"] = r.kernel2[v]
    np.savez_compressed(os.path.join(HERE, "decon_small.npz"), **out)

    rng = np.random.default_rng(5)
    img = rng.random((10, 12, 14)).astype(np.float32)
    k = rng.random((3, 5, 3)).astype(np.float32)
    conv = {"img": img, "kernel": k}
    for ext in range(5):
        conv[f"ext{ext}"] = O.convolve(img, k, ext, value=1.0, dtype=np.float64).astype(np.float32)
    conv["circular"] = O.circular_convolve(img, k, dtype=np.float64).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "conv_small.npz"), **conv)
    print("wrote", os.listdir(HERE))


if __name__ == "__main__":
    main()
