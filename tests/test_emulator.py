"""Index-math tests of the kernel bodies under the CPU emulator (tests/emu, built from the very
same sources with -DSPIM_HOST_EMU).  These are host-logic tests: they validate tiling, digit
reversal, extension-on-load, the R2C/C2R split steps and the fused epilogues before any GPU time is
spent.  The emulator is test infrastructure only and is never loaded by the package."""
import numpy as np
import pytest

import parity_cases as P
from oracle import mvdecon_oracle as O


def test_emulator_identifies_itself(emu_lib):
    assert b"EMULATOR" in emu_lib.mvd_version()


@pytest.mark.parametrize("ext", [0, 1, 2, 3, 4])
@pytest.mark.parametrize("shape,kshape", [((8, 8, 8), (3, 3, 3)), ((9, 7, 11), (3, 5, 3)), ((6, 6, 6), (4, 2, 6)),
                                          ((5, 30, 33), (1, 7, 9)), ((1, 16, 18), (1, 5, 5))])
def test_conv_all_extensions(emu_lib, shape, kshape, ext):
    P.conv_case(emu_lib, shape, kshape, ext)


def test_conv_halo_wider_than_image(emu_lib):
    # multiple reflections / wraps: kernel larger than the image along an axis
    for ext in (2, 3, 4):
        P.conv_case(emu_lib, (3, 4, 6), (7, 9, 5), ext)


@pytest.mark.parametrize("n", [16, 18, 20, 22, 24, 26, 28, 30, 36, 40, 42, 44, 48, 50, 52, 54, 56, 60, 64, 66, 70, 72, 78, 80,
                               84, 88, 90, 96, 98, 100, 104, 108, 110, 112, 120, 126, 128, 130, 132, 140, 144])
def test_every_radix_path_along_each_axis(emu_lib, n):
    # exact periodic mode keeps P = n, so the stage plan of n (and n/2 on x) is what runs
    rng = np.random.default_rng(n)
    k = rng.random((3, 3, 3), dtype=np.float32)
    for shape in ((n, 4, 8), (4, n, 8), (4, 4, n)):
        P.legacy_case(emu_lib, shape, (3, 3, 3), seed=n)


@pytest.mark.parametrize("shape,kshape", [((16, 16, 16), (5, 5, 5)), ((12, 20, 18), (3, 7, 5)), ((10, 9, 7), (3, 3, 3)),
                                          ((8, 8, 8), (8, 8, 8))])
def test_legacy_entry_is_circular_convolution(emu_lib, shape, kshape):
    P.legacy_case(emu_lib, shape, kshape)


def test_golden_conv(emu_lib):
    P.golden_conv_case(emu_lib)


@pytest.mark.parametrize("gen", [1, 2])
@pytest.mark.parametrize("typ", [0, 1, 2, 3])
def test_golden_deconvolution(emu_lib, gen, typ):
    P.golden_case(emu_lib, gen, typ)


@pytest.mark.parametrize("gen,typ", [(2, O.EFFICIENT_BAYESIAN), (1, O.OPTIMIZATION_I)])
def test_deconvolution_vs_oracle(emu_lib, gen, typ):
    P.decon_case(emu_lib, (14, 18, 22), 3, 5, typ, gen, 3)


def test_deconvolution_odd_dims_no_weights_no_tikhonov(emu_lib):
    P.decon_case(emu_lib, (9, 11, 13), 2, 3, O.INDEPENDENT, 2, 2, lam=0.0, use_weights=False)


def test_gen1_osem_from_overlap(emu_lib):
    P.decon_case(emu_lib, (10, 12, 14), 3, 3, O.OPTIMIZATION_II, 1, 2, osem_index=2)
    P.decon_case(emu_lib, (10, 12, 14), 3, 3, O.OPTIMIZATION_II, 1, 2, osem=2.0, osem_index=0)


def test_exact_tikhonov_switch(emu_lib):
    P.exact_tikhonov_case(emu_lib)


def test_error_paths(emu_lib):
    import ctypes
    from spim_registration_b200 import native
    from spim_registration_b200.deconvolution import Session
    with pytest.raises(native.NativeError):
        Session((4, 4, 4), 0, 2, lib=emu_lib)
    with pytest.raises(native.NativeError):
        Session((4, 4, 4), 1, 7, lib=emu_lib)
    with Session((4, 4, 4), 2, 2, lib=emu_lib) as s:
        s.set_view(0, np.ones((4, 4, 4), np.float32), None, np.ones((3, 3, 3), np.float32))
        with pytest.raises(native.NativeError, match="view 1 not set"):
            s.init()
        with pytest.raises(native.NativeError, match="not initialised"):
            s.run(1)
        with pytest.raises(ValueError):
            s.set_view(1, np.ones((4, 4, 5), np.float32), None, np.ones((3, 3, 3), np.float32))
