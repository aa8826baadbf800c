"""Pins the oracle against the REFERENCE's own Java CPU implementation -- when somebody has produced the vectors.

This build environment has no JVM, so `oracle/mvdecon_oracle.py` is "parity unpinned" (DESIGN.md section 2).
`tests/golden/reference/GenerateGolden.java` closes that gap for anyone with a Fiji installation: it runs
`MVDeconvolution` (gen-2), `BayesMVDeconvolution` (gen-1) for all four PSFTYPEs and one FFT convolution per out-of-bounds
rule on the inputs exported by `tests/golden/reference/export_inputs.py`, and writes raw float32 files to
`tests/golden/reference/out/`.  With those files present these tests compare the oracle (fp32 mode) with them at the
BASELINE tolerances (per-voxel 1e-3, relative L2 1e-4; a single convolution 2e-5 of the maximum); without them they skip
and say how to make them.  In particular `conv_imglib1_default` decides the one extension rule the oracle *assumes*
(ImgLib1's default OutOfBoundsStrategyMirrorFactory in gen-1, D2/Block.java:90)."""
import os

import numpy as np
import pytest

from oracle import mvdecon_oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, "golden", "reference", "out")
HOWTO = ("reference golden vectors not present: run tests/golden/reference/export_inputs.py, then "
         "tests/golden/reference/GenerateGolden.java with a Fiji class path (see the header of that file)")


def _need(name):
    p = os.path.join(REF, name)
    if not os.path.exists(p):
        pytest.skip(HOWTO)
    return p


def _load_inputs():
    d = np.load(os.path.join(HERE, "golden", "decon_small.npz"))
    V = int(d["num_views"])
    return d, [d[f"img{v}"] for v in range(V)], [d[f"w{v}"] for v in range(V)], [d[f"psf{v}"] for v in range(V)]


def test_exporter_round_trip(tmp_path, monkeypatch):
    """the raw files GenerateGolden.java reads are exactly the committed fixture (x fastest, little endian)"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("export_inputs", os.path.join(HERE, "golden", "reference", "export_inputs.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    monkeypatch.setattr(m, "HERE", str(tmp_path))
    m.main()
    d, imgs, ws, psfs = _load_inputs()
    meta = open(tmp_path / "inputs" / "meta.txt").read().split()
    nx, ny, nz, V = (int(x) for x in meta[:4])
    assert (nz, ny, nx) == imgs[0].shape and V == len(imgs)
    got = np.fromfile(tmp_path / "inputs" / "img1.raw", "<f4").reshape(nz, ny, nx)
    assert np.array_equal(got, imgs[1])


@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("typ", [0, 1, 2, 3])
def test_oracle_matches_reference_deconvolution(gen, typ):
    p = _need(f"psi_g{gen}_t{typ}.raw")
    d, imgs, ws, psfs = _load_inputs()
    ref = np.fromfile(p, "<f4").reshape(imgs[0].shape)
    r = O.deconvolve(imgs, ws, psfs, O.DeconParams(iteration_type=typ, num_iterations=2, lam=0.006, gen=gen))
    per, l2 = O.parity_errors(r.psi, ref)
    assert per <= 1e-3 and l2 <= 1e-4, (gen, typ, per, l2)


@pytest.mark.parametrize("typ", [0, 1, 2, 3])
def test_oracle_kernel2_matches_reference(typ):
    d, imgs, ws, psfs = _load_inputs()
    _need(f"k2_g2_t{typ}_v0.raw")
    k1, k2 = O.init_kernels(psfs, typ)
    for v in range(len(psfs)):
        ref = np.fromfile(_need(f"k2_g2_t{typ}_v{v}.raw"), "<f4").reshape(psfs[v].shape)
        assert np.abs(k2[v] - ref).max() <= 5e-5 * np.abs(ref).max(), (typ, v)


@pytest.mark.parametrize("name,ext,value", [("conv_imglib2_default", O.EXT_MIRROR_SINGLE, 0.0),
                                            ("conv_imglib2_value1", O.EXT_CONSTANT, 1.0),
                                            ("conv_imglib1_default", O.EXT_MIRROR_SINGLE, 0.0)])
def test_oracle_convolution_matches_reference_extension_rule(name, ext, value):
    p = _need(name + ".raw")
    c = np.load(os.path.join(HERE, "golden", "conv_small.npz"))
    ref = np.fromfile(p, "<f4").reshape(c["img"].shape)
    out = O.convolve(c["img"], c["kernel"], ext, value=value)
    assert np.abs(out - ref).max() <= 2e-5 * np.abs(ref).max(), name
