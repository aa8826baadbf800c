"""GPU parity tests of the narrow column tiles (8 x-frequencies per tile row, ColPassNarrow): selected automatically for FFT
lengths above ~880 (the 1080-long axes of a 1024 x 1024 x 512 volume with a 31^3 PSF on ONE GPU, where a 16-column tile
would leave a single block per SM), forced here with SPIM_COL_NARROW=1 on the ordinary cases as well.  The arithmetic per
column is the same as in the 16-column tiles, so the bar is the same as in test_gpu_parity.py.

The file sorts after every other GPU test on purpose: the variant was written after the round's GPU budget was spent
(verified under the kernel emulator, tests/test_emulator.py), so under `pytest -x` nothing here can hide a result above."""
import pytest

import parity_cases as P
from oracle import mvdecon_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [18, 30, 48, 56, 70, 90, 126, 144, 288, 560])
def test_narrow_tiles_radix_paths(gpu, monkeypatch, n):
    monkeypatch.setenv("SPIM_COL_NARROW", "1")
    c0 = gpu.mvd_debug_counter(0)
    for shape in ((n, 4, 8), (4, n, 8)):
        P.legacy_case(gpu, shape, (3, 3, 3), seed=n)
    assert gpu.mvd_debug_counter(0) > c0


@pytest.mark.parametrize("ext", [0, 1, 2, 3, 4])
def test_narrow_tiles_conv_all_extensions(gpu, monkeypatch, ext):
    monkeypatch.setenv("SPIM_COL_NARROW", "1")
    P.conv_case(gpu, (9, 7, 11), (3, 5, 3), ext)
    P.conv_case(gpu, (40, 50, 70), (7, 9, 5), ext)


def test_narrow_tiles_deconvolution(gpu, monkeypatch):
    monkeypatch.setenv("SPIM_COL_NARROW", "1")
    P.decon_case(gpu, (40, 48, 56), 3, 7, O.EFFICIENT_BAYESIAN, 2, 3)
    P.decon_case(gpu, (33, 41, 50), 2, 5, O.OPTIMIZATION_I, 1, 2)
    P.golden_case(gpu, 2, 2)
    P.golden_conv_case(gpu)


def test_long_axes_select_narrow_tiles_automatically(gpu):
    c0 = gpu.mvd_debug_counter(0)
    P.conv_case(gpu, (6, 1050, 40), (3, 31, 3), 2)        # Py = 1080
    c1 = gpu.mvd_debug_counter(0)
    assert c1 > c0
    P.conv_case(gpu, (1040, 6, 40), (15, 3, 3), 1)        # Pz = 1056 / 1080
    assert gpu.mvd_debug_counter(0) > c1


def test_configs2_size_on_one_gpu_delta_identity_and_shift(gpu):
    """BASELINE configs[2] volume on ONE GPU (1024 x 1024 x 512, 31^3 PSF -> FFT size 1080 x 1080 x 560, 2 GiB per real
    volume, 2.6 GB per spectrum): 64-bit index math and the automatically selected narrow tiles at full size, through
    properties the oracle is not needed for -- a delta kernel is the identity, a shifted delta a mirrored shift."""
    import numpy as np
    from spim_registration_b200 import native
    shape = (512, 1024, 1024)
    rng = np.random.default_rng(5)
    a = rng.random(shape, dtype=np.float32)
    delta = np.zeros((31, 31, 31), np.float32)
    delta[15, 15, 15] = 1.0
    c0 = gpu.mvd_debug_counter(0)
    out = native.convolve(a, delta, O.EXT_MIRROR_SINGLE, lib=gpu)
    assert gpu.mvd_debug_counter(0) > c0                  # the 1080-long y passes ran on narrow tiles
    assert np.abs(out - a).max() < 5e-6                   # fp32 round-off of a 1080 x 1080 x 560 transform pair on data in [0, 1)
    sh = np.zeros((31, 31, 31), np.float32)
    sh[20, 15, 15] = 1.0                                  # kernel index 20 along z = shift by +5 planes
    out = native.convolve(a, sh, O.EXT_MIRROR_SINGLE, lib=gpu)
    assert np.abs(out[5:] - a[:-5]).max() < 5e-6
    assert np.abs(out[:5] - a[5:0:-1]).max() < 5e-6       # mirror-single at the low z face


def test_column_staging_modes_are_bit_identical(gpu, monkeypatch):
    """SPIM_COLP (0 first stage from global memory, 2 one-shot cp.async staging, 3 persistent TMA pipeline) only changes how
    a column tile reaches shared memory."""
    import numpy as np
    from spim_registration_b200 import synthetic
    shape = (40, 48, 56)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 7, kind="beads")
    a, *_ = P.run_session(gpu, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 3)
    for colp in ("0", "2", "3"):
        monkeypatch.setenv("SPIM_COLP", colp)
        b, *_ = P.run_session(gpu, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, 2, 3)
        assert np.array_equal(a, b), colp


def test_bench_plan_x_lines(gpu):
    """x FFT length 560 (N2 = 280 = 8 * 7 * 5, the bench plan) and 1080 (N2 = 540 = 10 * 9 * 6, the 1024-wide volume on one
    GPU) through the TMA-fed x-forward kernel and both x-inverse block sizes."""
    P.decon_case(gpu, (16, 20, 524), 2, 31, O.EFFICIENT_BAYESIAN, 2, 2)
    P.decon_case(gpu, (12, 18, 1040), 2, 31, O.OPTIMIZATION_II, 2, 2)
    for ext in range(5):
        P.conv_case(gpu, (40, 50, 70), (7, 9, 5), ext)


def test_lean_column_pass(gpu):
    """Small tiles of plans without radices 9 / 10 (the 288-point z axis of the bench volume) run from the instantiation
    compiled for radices <= 8 -- 80 registers, six resident blocks."""
    for n in (48, 64, 96, 128, 288):
        for shp in ((n, 4, 8), (4, n, 8)):
            P.legacy_case(gpu, shp, (3, 3, 3), seed=n)


def test_deduplicated_forward_sweeps_are_bit_identical(gpu, monkeypatch):
    """Mirror extension: halo lines / planes are transformed once and stored twice (SPIM_DEDUP=0: every padded line)."""
    import numpy as np
    from spim_registration_b200 import synthetic
    for shape, k in (((40, 48, 56), 7), ((36, 70, 524), 31)):
        _, imgs, ws, psfs = synthetic.make_dataset(shape, 2, k, kind="beads")
        for gen in (1, 2):
            monkeypatch.setenv("SPIM_DEDUP", "1")
            c1 = gpu.mvd_debug_counter(1)
            a, *_ = P.run_session(gpu, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, gen, 2)
            assert gpu.mvd_debug_counter(1) > c1
            monkeypatch.setenv("SPIM_DEDUP", "0")
            b, *_ = P.run_session(gpu, imgs, ws, psfs, O.EFFICIENT_BAYESIAN, gen, 2)
            assert np.array_equal(a, b), (shape, gen)


def test_constant_extension_by_shift(gpu, monkeypatch):
    """gen-2 conv2: zero extension of (quotient - 1) plus sum(K2) instead of the literal extension by 1 (SPIM_CONST_SHIFT=0)."""
    from spim_registration_b200 import synthetic
    shape = (40, 48, 56)
    _, imgs, ws, psfs = synthetic.make_dataset(shape, 3, 7, kind="beads")
    for typ in (O.EFFICIENT_BAYESIAN, O.INDEPENDENT, O.OPTIMIZATION_I, O.OPTIMIZATION_II):
        monkeypatch.setenv("SPIM_CONST_SHIFT", "1")
        a, *_ = P.run_session(gpu, imgs, ws, psfs, typ, 2, 3)
        monkeypatch.setenv("SPIM_CONST_SHIFT", "0")
        b, *_ = P.run_session(gpu, imgs, ws, psfs, typ, 2, 3)
        per, l2 = O.parity_errors(a, b)
        assert per <= 2e-5 and l2 <= 2e-6, (typ, per, l2)
    monkeypatch.setenv("SPIM_CONST_SHIFT", "0")
    P.decon_case(gpu, shape, 3, 7, O.EFFICIENT_BAYESIAN, 2, 3)
