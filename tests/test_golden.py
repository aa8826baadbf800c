"""Oracle (fp32 mode) against the committed golden fixtures (oracle fp64 outputs; see
tests/golden/make_golden.py for why these are self-generated: the reference pins nothing)."""
import os

import numpy as np
import pytest

from oracle import mvdecon_oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_decon():
    d = np.load(os.path.join(G, "decon_small.npz"))
    V = int(d["num_views"])
    imgs = [d[f"img{v}"] for v in range(V)]
    ws = [d[f"w{v}"] for v in range(V)]
    psfs = [d[f"psf{v}"] for v in range(V)]
    return d, imgs, ws, psfs


@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("typ", [0, 1, 2, 3])
def test_oracle_fp32_matches_golden(gen, typ):
    d, imgs, ws, psfs = load_decon()
    r = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=typ, num_iterations=2, lam=0.006, gen=gen))
    per, l2 = O.parity_errors(r.psi, d[f"psi_g{gen}_t{typ}"])
    assert per <= 1e-3 and l2 <= 1e-4          # BASELINE.md section 5 tolerances
    assert np.isclose(r.avg, float(d[f"avg_g{gen}_t{typ}"]), rtol=1e-6)


def test_oracle_conv_matches_golden():
    d = np.load(os.path.join(G, "conv_small.npz"))
    for ext in range(5):
        out = O.convolve(d["img"], d["kernel"], ext, value=1.0)
        assert np.abs(out - d[f"ext{ext}"]).max() / np.abs(d[f"ext{ext}"]).max() < 1e-5
    out = O.circular_convolve(d["img"], d["kernel"])
    assert np.abs(out - d["circular"]).max() / np.abs(d["circular"]).max() < 1e-5
